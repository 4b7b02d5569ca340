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


def allgather_slabs(local, bounds, nxny: int, group=None):
    """All-gather of the z slabs of a node vector into the full vector on every rank (SURVEY.md section 8e,
    "Evaluation": replicate c after a slab solve, then split the query points).  ``local`` is a 1-D torch tensor
    holding this rank's owned planes (CPU under gloo, CUDA under nccl); slabs may differ by one plane, so every rank
    sends a buffer padded to the largest slab."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if len(bounds) != world:
        raise ValueError(f"{len(bounds)} slabs for {world} ranks")
    rank = dist.get_rank(group)
    sizes = [(k1 - k0) * int(nxny) for k0, k1 in bounds]
    if local.numel() != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.numel()} values, its slab has {sizes[rank]}")
    pad = max(sizes)
    send = torch.zeros(pad, dtype=local.dtype, device=local.device)
    send[: sizes[rank]] = local
    recv = [torch.empty(pad, dtype=local.dtype, device=local.device) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    return torch.cat([recv[r][: sizes[r]] for r in range(world)])


def split_points(n: int, nranks: int) -> list[tuple[int, int]]:
    """Even split of n query points over the ranks: [(first, last), ...]."""
    base, extra = divmod(int(n), int(nranks))
    out, a = [], 0
    for r in range(nranks):
        m = base + (1 if r < extra else 0)
        out.append((a, a + m))
        a += m
    return out


def replicate_solution(slab_engine, bounds, device_index: int, group=None):
    """After a z-slab solve: a new Engine on this rank's GPU that owns every plane and holds the whole solution
    (device-to-device all-gather, no host round trip); evaluation then runs on each rank's share of the points."""
    import torch

    from .engine import Engine

    nxny = int(slab_engine.nsteps[0] * slab_engine.nsteps[1])
    dev = torch.device("cuda", device_index)
    local = torch.empty(slab_engine.local_dof, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    slab_engine.get_c_device(local.data_ptr(), st)
    full_c = allgather_slabs(local, bounds, nxny, group)
    full = Engine(slab_engine.nsteps, slab_engine.origin, slab_engine.step, kind=slab_engine.kind, device=device_index)
    torch.cuda.current_stream(dev).synchronize()
    full.set_c_device(full_c.data_ptr(), st)
    return full
