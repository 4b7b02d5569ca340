#!/usr/bin/env bash
# Installs the UNMODIFIED reference (LoopStructural 1.8.0 + its two workspace packages loop_common / loop_interpolation,
# all pure Python) into baseline/_ref, which is git-ignored but travels to the GPU box with the snapshot.  Used by
# `bench.py --impl reference` (the reference arm) and by the device tests that run the plug-in inside the reference's own
# call chain.  Nothing of the product imports it.  Run in the build container (needs /root/reference):
#     bash baseline/install_reference.sh
# The reference's source tree is read-only, and setuptools writes build/ and *.egg-info into it: install from a copy.
# --no-deps: numpy / scipy / pandas / scikit-learn are in the image; pyvista and scikit-image are not (the reference
# hard-imports both) -- tests/golden/_stubs holds two import stubs for them, put on sys.path by loopstructural_b200.refenv.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="${1:-/root/reference}"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
cp -r "$SRC" "$TMP/ref"
chmod -R u+w "$TMP/ref"
rm -rf "$HERE/_ref"
for pkg in "$TMP/ref/packages/loop_common" "$TMP/ref/packages/loop_interpolation" "$TMP/ref"; do
  python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$HERE/_ref" "$pkg"
done
find "$HERE/_ref" -name __pycache__ -type d -prune -exec rm -rf {} +
echo "installed: $(ls "$HERE/_ref" | tr '\n' ' ')"
