"""The P1 / TetMesh kernels (k_p1_normal, k_p1_diag, k_p1_point_rows, unmodified from loopstructural_b200/csrc/p1.cuh)
run on the CPU by tests/hostsim on the full mesh and on every z slab of 2 / 3 / 5-way partitions (halo planes filled
as the NCCL exchange would), against the oracle's explicit face-jump rows and point rows.  This is what checks the
slab indexing of the P1 path without a multi-GPU box."""
import ctypes as C
import glob
import os
import shutil
import subprocess

import numpy as np
import pytest

from loopstructural_b200 import p1_tables, partition
from oracle import fdi_oracle as fo
from oracle import p1_oracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostsim", "p1_emu.cpp")
LIB = os.path.join(HERE, "hostsim", "libp1_emu.so")
CSRC = os.path.join(os.path.dirname(HERE), "loopstructural_b200", "csrc")
dp = C.POINTER(C.c_double)
llp = C.POINTER(C.c_longlong)


@pytest.fixture(scope="module")
def emu():
    gxx = shutil.which("g++")
    if gxx is None or not os.path.isdir("/usr/local/cuda/include"):
        pytest.skip("g++ / CUDA headers not available")
    deps = [SRC, os.path.join(HERE, "hostsim", "cuda_emu.h")] + glob.glob(os.path.join(CSRC, "*.cuh"))
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.run([gxx, "-O1", "-std=c++20", "-pthread", "-mfma", "-w", "-DLSB_EMU", "-I/usr/local/cuda/include", "-fPIC",
                        "-fvisibility=hidden", "-Wl,-Bsymbolic", "-shared", "-o", LIB, SRC], check=True, capture_output=True)
    return C.CDLL(LIB)


def _bounds(nz, world):
    return partition.slab_bounds(nz, world) if world > 1 else [(0, nz)]


@pytest.mark.parametrize("tiles", [0, 1, 2, 3, (11, 0), (11, 3), (12, 0), (12, 5), (13, 4), (20, 0), (20, 3)])
@pytest.mark.parametrize("cells,step,world", [((5, 4, 6), (1.0, 1.0, 1.0), 1), ((7, 8, 9), (1.0, 0.7, 1.3), 2),
                                               ((8, 7, 11), (0.9, 1.1, 1.0), 3), ((9, 6, 10), (1.0, 1.0, 2.0), 5),
                                               ((20, 13, 15), (1.0, 1.2, 0.8), 2), ((12, 6, 17), (1.0, 1.0, 1.0), 3),
                                               ((70, 9, 8), (1.0, 1.1, 0.9), 1), ((11, 12, 13), (1.3, 1.0, 0.7), 4)])
def test_face_jump_operator_on_every_slab_matches_the_explicit_rows(emu, cells, step, world, tiles):
    """tiles > 0: the k_p1_tiles plan with tile configuration tiles - 1 (interior nodes from de-interleaved shared-memory
    tiles with the eight class tables as constant-bank operands, boundary nodes in the generic form by the first CTAs of
    the same launch); meshes with an axis of <= 6 cells have no interior classes and fall back to k_p1_normal."""
    tiles, cz = tiles if isinstance(tiles, tuple) else (tiles, 0)
    grid = fo.Grid.from_cells(np.array(cells), np.zeros(3), np.array(step))
    fam = po.face_rows(grid, 1.0)
    R = fam.matrix(grid.n_nodes)
    w2, ridge2 = 0.49, 1e-3
    N = (R.T @ R).tocsr() * w2
    rng = np.random.default_rng(sum(cells))
    p = rng.standard_normal(grid.n_nodes)
    ref = N @ p + ridge2 * p
    dref = N.diagonal() + ridge2
    m, table = p1_tables.regularisation_table(np.array(cells), np.array(step))
    m = np.ascontiguousarray(m, dtype=np.int64)
    table = np.ascontiguousarray(table, dtype=np.float64)
    nx, ny, nz = (int(v) for v in grid.nsteps)
    plane = nx * ny
    stp = np.ascontiguousarray(step, dtype=np.float64)
    total = 0.0
    for k0, k1 in _bounds(nz, world):
        dims = (C.c_int * 5)(nx, ny, nz, k0, k1)
        q, d = np.empty(plane * (k1 - k0)), np.empty(plane * (k1 - k0))
        pq = C.c_double(0.0)
        rc = emu.emu_p1_normal(dims, stp.ctypes.data_as(dp), m.ctypes.data_as(llp), table.ctypes.data_as(dp), C.c_double(w2),
                               C.c_double(ridge2), tiles, cz, p.ctypes.data_as(dp), q.ctypes.data_as(dp), d.ctypes.data_as(dp), C.byref(pq))
        assert rc >= 0, f"slab {k0, k1}: rc {rc} (-2 = written outside the owned planes)"
        assert rc == (1 if tiles and min(cells) > 6 else 0), (rc, cells, (k0, k1))
        sl = slice(k0 * plane, k1 * plane)
        assert not np.isnan(q).any()
        assert np.abs(q - ref[sl]).max() <= 1e-12 * np.abs(ref).max(), f"slab {k0, k1}"
        assert np.abs(d - dref[sl]).max() <= 1e-12 * np.abs(dref).max(), f"slab {k0, k1}"
        assert abs(pq.value - float(p[sl] @ ref[sl])) <= 1e-10 * abs(float(p @ ref))
        total += pq.value
    assert abs(total - float(p @ ref)) <= 1e-10 * abs(float(p @ ref))


@pytest.mark.parametrize("world", [1, 2, 4])
def test_point_rows_on_every_slab_match_the_oracle(emu, world):
    """k_p1_point_rows on slabs: every rank keeps the rows whose cell touches its planes, with the node ids of the
    global mesh; the union over ranks is the oracle's family (a cell that straddles two slabs belongs to both)."""
    from scipy import sparse

    cells = (6, 5, 8)
    grid = fo.Grid.from_cells(np.array(cells), np.array([0.5, -1.0, 2.0]), np.array([1.0, 0.8, 1.25]))
    rng = np.random.default_rng(4)
    lo, hi = grid.origin, grid.origin + grid.step * np.array(cells)
    n = 400
    xyz = lo + rng.uniform(-0.05, 1.05, (n, 3)) * (hi - lo)
    value = np.column_stack([xyz, rng.normal(size=n), np.ones(n)])
    vec = np.column_stack([xyz, rng.normal(size=(n, 3)), np.zeros(n)])
    nx, ny, nz = (int(v) for v in grid.nsteps)
    org, stp = np.ascontiguousarray(grid.origin), np.ascontiguousarray(grid.step)
    fold = vec.copy()
    fold[:, 6] = 0.7
    cases = ((0, value, 5, po.value_rows(grid, value, 1.0), 1), (2, vec, 7, po.normal_rows(grid, vec, 1.0), 3),
             (3, vec, 7, po.orthogonal_rows(grid, vec[:, :3], vec[:, 3:6], 1.0), 1),
             (4, fold, 7, po.element_vector_rows(grid, fold[:, :3], fold[:, 3:6], 1.0, 0.7), 1))
    for kind, data, ncols, fam, mult in cases:
        data = np.ascontiguousarray(data)
        M_ref = fam.matrix(grid.n_nodes)
        _c, _ids, _v, _t, ins = po.locate(grid, data[:, :3])
        cz = np.clip(np.floor((data[:, 2] - grid.origin[2]) / grid.step[2]).astype(int), 0, cells[2] - 1)
        seen = np.zeros(n * mult, dtype=bool)
        for k0, k1 in _bounds(nz, world):
            dims = (C.c_int * 5)(nx, ny, nz, k0, k1)
            nr = n * mult
            vals, cols = np.empty((nr, 4)), np.empty((nr, 4), dtype=np.int64)
            b, w2 = np.empty(nr), np.empty(nr)
            kept = emu.emu_p1_point_rows(dims, org.ctypes.data_as(dp), stp.ctypes.data_as(dp), kind, C.c_longlong(n),
                                         data.ctypes.data_as(dp), ncols, C.c_double(1.0), vals.ctypes.data_as(dp),
                                         cols.ctypes.data_as(llp), b.ctypes.data_as(dp), w2.ctypes.data_as(dp))
            touches = ins & (cz + 1 >= k0) & (cz < k1)
            assert kept == int(touches.sum()) * mult
            live = np.repeat(touches, mult)  # rows of point r are r*mult .. r*mult + mult - 1
            assert np.array_equal(cols[:, 0] >= 0, live)
            # compare the live rows with the oracle's rows of the same points
            ref_rows = np.flatnonzero(np.repeat(ins, mult))           # oracle families hold inside points only
            pos = {int(r): i for i, r in enumerate(ref_rows)}
            sel = np.flatnonzero(live)
            M = sparse.coo_matrix((vals[sel].ravel(), (np.repeat(np.arange(sel.size), 4), cols[sel].ravel())),
                                  shape=(sel.size, grid.n_nodes)).tocsr()
            ref_sel = np.array([pos[int(r)] for r in sel], dtype=int)
            assert abs(M - M_ref[ref_sel]).max() <= 1e-12
            np.testing.assert_allclose(b[sel], fam.b[ref_sel], rtol=0, atol=1e-12)
            np.testing.assert_allclose(w2[sel], fam.w[ref_sel] ** 2, rtol=1e-12, atol=0)
            seen |= live
        assert np.array_equal(seen, np.repeat(ins, mult))


@pytest.mark.parametrize("world", [1, 3])
def test_face_vector_rows_on_every_slab_match_the_oracle(emu, world):
    """k_p1_face_vector_vals + k_rows_ingest (lsb200_add_face_vector_rows): direction-weighted face-jump rows from the
    node ids of the two tetrahedra of a shared face; a rank keeps the rows with a node in its planes."""
    from scipy import sparse

    cells = (4, 3, 7)
    grid = fo.Grid.from_cells(np.array(cells), np.array([0.5, -1.0, 2.0]), np.array([1.0, 0.8, 1.25]))
    plain = po.face_rows(grid, 1.0)          # every interior face once: its index array lists the two tetrahedra
    na, nb = np.ascontiguousarray(plain.idx[:, :4]), np.ascontiguousarray(plain.idx[:, 4:])
    n = na.shape[0]
    rng = np.random.default_rng(9)
    vec = np.ascontiguousarray(rng.normal(size=(n, 3)))
    fam = po.face_vector_rows(grid, na, nb, vec, 0.3)
    M_ref = fam.matrix(grid.n_nodes)
    nx, ny, nz = (int(v) for v in grid.nsteps)
    org, stp = np.ascontiguousarray(grid.origin), np.ascontiguousarray(grid.step)
    plane = nx * ny
    kz = np.hstack([na, nb]) // plane
    seen = np.zeros(n, dtype=bool)
    for k0, k1 in _bounds(nz, world):
        dims = (C.c_int * 5)(nx, ny, nz, k0, k1)
        vals, cols = np.empty((n, 8)), np.empty((n, 8), dtype=np.int64)
        b, w2 = np.empty(n), np.empty(n)
        kept = emu.emu_p1_face_vector_rows(dims, org.ctypes.data_as(dp), stp.ctypes.data_as(dp), C.c_longlong(n),
                                           na.ctypes.data_as(llp), nb.ctypes.data_as(llp), vec.ctypes.data_as(dp), C.c_double(0.3),
                                           vals.ctypes.data_as(dp), cols.ctypes.data_as(llp), b.ctypes.data_as(dp), w2.ctypes.data_as(dp))
        touches = np.any((kz >= k0) & (kz < k1), axis=1)
        assert kept == int(touches.sum())
        assert np.array_equal(cols[:, 0] >= 0, touches)
        sel = np.flatnonzero(touches)
        M = sparse.coo_matrix((vals[sel].ravel(), (np.repeat(np.arange(sel.size), 8), cols[sel].ravel())),
                              shape=(sel.size, grid.n_nodes)).tocsr()
        assert abs(M - M_ref[sel]).max() <= 1e-12
        assert not np.any(b[sel])
        np.testing.assert_allclose(w2[sel], 0.09, rtol=1e-12)
        seen |= touches
    assert seen.all()


@pytest.mark.parametrize("world", [1, 2, 3])
def test_directional_stencil_rows_on_every_slab_match_the_oracle(emu, world):
    """k_fdi_directional_rows (lsb200_add_directional_rows; odd nx: the storage has an even row pitch): per slab the live
    rows are the oracle's rows whose stencil touches an owned plane, with the same weights."""
    from scipy import sparse

    cells = (6, 5, 8)
    grid = fo.Grid.from_cells(np.array(cells), np.array([0.5, -1.0, 2.0]), np.array([1.0, 0.8, 1.25]))
    nx, ny, nz = (int(v) for v in grid.nsteps)
    rng = np.random.default_rng(12)
    vec = rng.normal(size=(grid.n_nodes, 3))
    vec[rng.random(grid.n_nodes) < 0.2, 0] = 0.0
    vec[nx * ny * 3: nx * ny * 4] = 0.0
    vec[7] = np.nan
    vec = np.ascontiguousarray(vec)
    fams = fo.directional_families(grid, vec, 0.7)
    ref = sparse.vstack([f.matrix(grid.n_nodes).multiply(f.w[:, None]) for f in fams]).tocsr()
    p = rng.standard_normal(grid.n_nodes)
    q_ref = ref.T @ (ref @ p)
    org, stp = np.ascontiguousarray(grid.origin), np.ascontiguousarray(grid.step)
    q = np.zeros(grid.n_nodes)
    for k0, k1 in _bounds(nz, world):
        ka, kb = max(1, k0 - 1), min(nz - 1, k1 + 1)
        nr = 6 * (nx - 2) * (ny - 2) * (kb - ka)
        dims = (C.c_int * 5)(nx, ny, nz, k0, k1)
        vals, cols, w2 = np.empty((nr, 4)), np.empty((nr, 4), dtype=np.int64), np.empty(nr)
        kept = emu.emu_fdi_directional_rows(dims, org.ctypes.data_as(dp), stp.ctypes.data_as(dp), vec.ctypes.data_as(dp),
                                            C.c_double(0.7), vals.ctypes.data_as(dp), cols.ctypes.data_as(llp), w2.ctypes.data_as(dp))
        live = w2 > 0
        assert kept == int(live.sum())
        cc = np.where(cols < 0, 0, cols)
        vv = np.where(cols < 0, 0.0, vals)
        M = sparse.coo_matrix((vv[live].ravel(), (np.repeat(np.arange(int(live.sum())), 4), cc[live].ravel())),
                              shape=(int(live.sum()), grid.n_nodes)).tocsr()
        part = M.T @ (w2[live] * (M @ p))
        own = slice(k0 * nx * ny, k1 * nx * ny)
        q[own] = part[own]  # a rank applies its rows to the planes it owns
        if world == 1:
            assert kept == ref.shape[0]
    assert np.abs(q - q_ref).max() <= 1e-12 * np.abs(q_ref).max()
