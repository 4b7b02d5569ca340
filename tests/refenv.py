"""Where the UNMODIFIED reference lives (test / benchmark infrastructure, never imported by the product).

Two places, in this order:
  * ``baseline/_ref``  -- the pip ``--target`` install made by ``baseline/install_reference.sh`` (git-ignored, travels to
    the GPU box with the snapshot): the only copy that exists on a GPU box;
  * ``/root/reference`` -- the source tree mounted in the build container.
Both need the two import stubs of ``tests/golden/_stubs`` (pyvista, skimage: hard imports of the reference that are not
installed in this image)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
STUBS = os.path.join(HERE, "golden", "_stubs")
INSTALLED = os.path.join(ROOT, "baseline", "_ref")
MOUNT = "/root/reference"


def reference_paths(prefer_installed=True):
    """sys.path entries that make ``LoopStructural`` / ``loop_interpolation`` / ``loop_common`` importable, or []."""
    have_inst = os.path.isdir(os.path.join(INSTALLED, "loop_interpolation"))
    have_mount = os.path.isdir(os.path.join(MOUNT, "packages"))
    if have_inst and (prefer_installed or not have_mount):
        return [STUBS, INSTALLED]
    if have_mount:
        return [STUBS, f"{MOUNT}/packages/loop_common/src", f"{MOUNT}/packages/loop_interpolation/src", MOUNT]
    return []


def available():
    return bool(reference_paths())


def activate(quiet=True):
    """Put the reference on sys.path (idempotent).  Returns where it came from: 'installed', 'mount' or None."""
    paths = reference_paths()
    if not paths:
        return None
    sys.dont_write_bytecode = True
    for p in reversed(paths):
        if p not in sys.path:
            sys.path.insert(0, p)
    if quiet:
        import logging

        logging.disable(logging.CRITICAL)
    return "installed" if INSTALLED in paths else "mount"
