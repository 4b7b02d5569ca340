"""Host logic of the z-slab path on CPU: slab bounds, row distribution, and a world_size-2 gloo run in which
each rank applies the oracle's operator to its slab + exchanged halo planes and reproduces the global result
(the partition-independence the device path relies on: row masks use GLOBAL indices)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from loopstructural_b200 import partition
from oracle import fdi_oracle as fo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_bounds_cover_the_domain():
    for nz, n in [(7, 2), (256, 8), (1024, 8), (9, 4), (16, 3)]:
        b = partition.slab_bounds(nz, n)
        assert b[0][0] == 0 and b[-1][1] == nz and all(b[i][1] == b[i + 1][0] for i in range(n - 1))
        assert min(k1 - k0 for k0, k1 in b) >= partition.HALO and max(k1 - k0 for k0, k1 in b) - min(k1 - k0 for k0, k1 in b) <= 1
    with pytest.raises(ValueError):
        partition.slab_bounds(3, 2)
    assert partition.owner_of_plane(partition.slab_bounds(10, 3), 9) == 2
    lo, hi = partition.halo_planes(partition.slab_bounds(10, 3), 0, 10)
    assert lo == [] and hi == [4, 5]


def test_rows_for_slab_selects_rows_touching_owned_planes():
    cols = np.array([[0, 1, 16, 17], [32, 33, 48, 49], [64, 65, 80, 81]])  # 4x4 planes: planes (0,1),(2,3),(4,5)
    assert partition.rows_for_slab(cols, 16, 0, 2).tolist() == [True, False, False]
    assert partition.rows_for_slab(cols, 16, 3, 5).tolist() == [False, True, True]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nodes = np.array([9, 7, 13])
    grid = fo.Grid(nodes, np.zeros(3), np.ones(3))
    w = dict(dxx=1.3, dyy=0.8, dzz=1.1, dxy=2.0, dyz=0.5, dxz=1.5, dx=0.7, dy=1.9, dz=1.2)
    rng = np.random.default_rng(5)
    p = rng.standard_normal(grid.n_nodes)
    nxny = int(nodes[0] * nodes[1])
    bounds = partition.slab_bounds(int(nodes[2]), world)
    k0, k1 = bounds[rank]
    mine = torch.from_numpy(partition.scatter_slabs(p, nxny, bounds)[rank].copy())
    H = partition.HALO
    # halo exchange: 2 planes down, 2 planes up (what halo_exchange() does with ncclSend/ncclRecv)
    lo = torch.zeros(H * nxny, dtype=torch.float64)
    hi = torch.zeros(H * nxny, dtype=torch.float64)
    reqs = []
    if rank > 0:
        reqs += [dist.isend(mine[: H * nxny].clone(), rank - 1), dist.irecv(lo, rank - 1)]
    if rank < world - 1:
        reqs += [dist.isend(mine[-H * nxny:].clone(), rank + 1), dist.irecv(hi, rank + 1)]
    for r in reqs:
        r.wait()
    ext = np.concatenate([lo.numpy() if rank > 0 else np.zeros(0), mine.numpy(), hi.numpy() if rank < world - 1 else np.zeros(0)])
    # the global operator restricted to the owned planes, from owned + halo values only
    full = np.zeros(grid.n_nodes)
    a = (k0 - (H if rank > 0 else 0)) * nxny
    full[a:a + ext.size] = ext
    q = fo.regularisation_normal_matvec(grid, w, full)[k0 * nxny:k1 * nxny]
    pq = torch.tensor([float(np.dot(mine.numpy(), q))], dtype=torch.float64)
    dist.all_reduce(pq)  # the p.q all-reduce of the CG loop
    parts = [None] * world
    dist.all_gather_object(parts, q)
    if rank == 0:
        ref = fo.regularisation_normal_matvec(grid, w, p)
        got = partition.gather_slabs(parts)
        out.put((float(np.abs(got - ref).max()), float(pq.item()), float(np.dot(p, ref))))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_operator_is_partition_independent_gloo(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29700 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for pr in procs:
        pr.start()
    err, pq, pq_ref = out.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert err <= 1e-12
    assert abs(pq - pq_ref) <= 1e-10 * abs(pq_ref)


def _gather_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nx, ny, nz = 5, 4, 11  # 11 planes over 3 ranks: slabs of 4, 4, 3 planes
    nxny = nx * ny
    c = np.random.default_rng(9).standard_normal(nxny * nz)
    bounds = partition.slab_bounds(nz, world)
    k0, k1 = bounds[rank]
    full = partition.allgather_slabs(torch.from_numpy(c[k0 * nxny:k1 * nxny].copy()), bounds, nxny)
    # every rank evaluates its share of the query points on the replicated field (the oracle stands in for the kernel)
    grid = fo.Grid(np.array([nx, ny, nz]), np.zeros(3), np.ones(3))
    pts = np.random.default_rng(3).uniform(-0.5, 10.5, (1001, 3)) * np.array([0.4, 0.3, 1.0])
    a, b = partition.split_points(len(pts), world)[rank]
    mine = fo.evaluate_value(grid, pts[a:b], full.numpy())
    parts = [None] * world
    dist.all_gather_object(parts, (a, b, mine))
    if rank == 0:
        got = np.concatenate([m for _a, _b, m in sorted(parts, key=lambda t: t[0])])
        ref = fo.evaluate_value(grid, pts, c)
        out.put((bool(np.array_equal(full.numpy(), c)), bool(np.array_equal(got, ref, equal_nan=True)), [p[:2] for p in parts]))
    dist.destroy_process_group()


def test_allgather_of_unequal_slabs_and_point_split_gloo():
    """Replicating the solution of a slab solve on every rank (padded all-gather of unequal slabs) and splitting the
    query points: the path partition.replicate_solution drives with CUDA tensors over NCCL."""
    assert partition.split_points(10, 3) == [(0, 4), (4, 7), (7, 10)] and partition.split_points(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    world = 3
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, 29731, out)) for r in range(world)]
    for pr in procs:
        pr.start()
    same_c, same_eval, spans = out.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert same_c and same_eval and sorted(spans) == [(0, 334), (334, 668), (668, 1001)]
