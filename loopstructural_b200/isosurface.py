"""Isosurface of a lattice volume on the device (csrc/iso.cuh, lsb200_isosurface): the call LoopIsosurfacer.fit makes to
skimage.measure.marching_cubes (LoopStructural/utils/_surface.py:139-147), same arguments, same four return values.

The triangulation is marching tetrahedra (six Kuhn tetrahedra per cell), not scikit-image's Lewiner cubes: the level set
of the same piecewise-linear interpolant of the same samples, one vertex per crossed lattice edge, watertight; more
triangles.  scikit-image is not installed in this image, so no fixture of its output exists (parity unpinned against it);
oracle/iso_oracle.py restates the construction in numpy and the tests check both against analytic surfaces."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nat


def marching_tetrahedra(volume, level=None, *, spacing=(1.0, 1.0, 1.0), mask=None, device: int = 0, **_ignored):
    """(verts (V,3), faces (F,3) int32, normals (V,3), values (V,)) like skimage.measure.marching_cubes.
    ``mask``: boolean array, False = sample not to be used (as NaN samples are)."""
    vol = np.array(volume, dtype=np.float64, order="C", copy=mask is not None)
    if vol.ndim != 3:
        raise ValueError("Input volume should be a 3D numpy array.")
    if min(vol.shape) < 2:
        raise ValueError("Input array must be at least 2x2x2.")
    if level is None:
        level = 0.5 * (np.nanmin(vol) + np.nanmax(vol))
    level = float(level)
    if mask is not None:
        vol[~np.asarray(mask, dtype=bool)] = np.nan
    if not (np.nanmin(vol) <= level <= np.nanmax(vol)):
        raise ValueError("Surface level must be within volume data range.")
    lib = nat.load()
    dims = np.ascontiguousarray(vol.shape, dtype=np.int64)
    sp = nat.as_f64(spacing, (3,))
    iso = C.c_void_p()
    nv, nt = C.c_int64(0), C.c_int64(0)
    nat.check(lib.lsb200_isosurface(int(device), nat.iptr(dims), nat.dptr(vol), level, nat.dptr(sp), C.byref(iso), C.byref(nv), C.byref(nt)))
    try:
        verts = np.empty((nv.value, 3), dtype=np.float64)
        normals = np.empty((nv.value, 3), dtype=np.float64)
        faces = np.empty((nt.value, 3), dtype=np.int32)
        nat.check(lib.lsb200_isosurface_fetch(iso, nat.dptr(verts), faces.ctypes.data_as(C.POINTER(C.c_int32)), nat.dptr(normals)))
    finally:
        lib.lsb200_isosurface_free(iso)
    if nt.value == 0:
        raise RuntimeError("No surface found at the given iso value.")
    return verts, faces, normals, np.full(nv.value, level)
