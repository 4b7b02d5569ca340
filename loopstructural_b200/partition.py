"""z-slab decomposition of the structured lattice for multi-GPU solves.

The flat node index gi = i + nx*j + nx*ny*k (LC/supports/_3d_base_structured.py:346-350) makes
z-slabs contiguous.  Rank g owns the planes [k0_g, k1_g); the stencil operator needs the two planes
on either side (halo 2), exchanged once per matvec.  Pure host logic: testable on CPU with gloo.
"""
from __future__ import annotations

import numpy as np

HALO = 2


def slab_bounds(nz: int, nranks: int) -> list[tuple[int, int]]:
    """Contiguous, near-equal plane ranges; every rank owns at least HALO planes so that a halo is
    always served by the direct neighbour."""
    nz, nranks = int(nz), int(nranks)
    if nranks < 1:
        raise ValueError("nranks must be >= 1")
    if nz < HALO * nranks:
        raise ValueError(f"{nz} planes cannot be split into {nranks} slabs of at least {HALO} planes")
    base, extra = divmod(nz, nranks)
    out, k = [], 0
    for r in range(nranks):
        n = base + (1 if r < extra else 0)
        out.append((k, k + n))
        k += n
    return out


def halo_planes(bounds: list[tuple[int, int]], rank: int, nz: int):
    """(lower halo planes, upper halo planes) of ``rank`` as global plane indices inside the domain."""
    k0, k1 = bounds[rank]
    lo = [k for k in range(k0 - HALO, k0) if k >= 0]
    hi = [k for k in range(k1, k1 + HALO) if k < nz]
    return lo, hi


def owner_of_plane(bounds: list[tuple[int, int]], k: int) -> int:
    for r, (k0, k1) in enumerate(bounds):
        if k0 <= k < k1:
            return r
    raise ValueError(f"plane {k} outside the domain")


def rows_for_slab(cols: np.ndarray, nxny: int, k0: int, k1: int) -> np.ndarray:
    """Boolean mask of the rows a slab must apply: those with at least one column on an owned plane.
    Such a row must have every column within the halo planes (checked by the device ingest)."""
    kz = np.asarray(cols, dtype=np.int64) // int(nxny)
    return np.any((kz >= k0) & (kz < k1), axis=1)


def scatter_slabs(vec: np.ndarray, nxny: int, bounds) -> list[np.ndarray]:
    return [np.ascontiguousarray(vec[k0 * nxny:k1 * nxny]) for k0, k1 in bounds]


def gather_slabs(parts: list[np.ndarray]) -> np.ndarray:
    return np.concatenate(parts)
