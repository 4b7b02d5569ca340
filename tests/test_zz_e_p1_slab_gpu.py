"""P1 / TetMesh z-slab handles on ONE device (lsb200_create_tetmesh_slab): without neighbours the halo planes stay
zero, so a slab applies the global face-jump operator to p extended by zero; point rows keep the cells that touch the
slab.  The NCCL solve on slabs is tests/slab_worker.py; the same kernels on the CPU emulator: tests/test_p1_emu.py.
(Written after the round's GPU budget was spent: not yet run on a device, hence last.)"""
import numpy as np
import pytest

from loopstructural_b200 import partition
from oracle import fdi_oracle as fo
from oracle import p1_oracle as po

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cells,step,world", [((8, 7, 11), (0.9, 1.1, 1.0), 3), ((17, 16, 24), (1.0, 1.0, 1.0), 4)])
def test_face_jump_operator_and_rows_on_slab_handles(cells, step, world):
    from loopstructural_b200 import Engine, build
    from loopstructural_b200 import engine as en

    build.build()
    grid = fo.Grid.from_cells(np.array(cells), np.zeros(3), np.array(step))
    R = po.face_rows(grid, 1.0).matrix(grid.n_nodes)
    cgw, ridge = 0.7, 0.05
    N = (R.T @ R).tocsr() * cgw ** 2
    nx, ny, nz = (int(v) for v in grid.nsteps)
    plane = nx * ny
    rng = np.random.default_rng(1)
    lo, hi = grid.origin, grid.origin + grid.step * np.array(cells)
    xyz = lo + rng.uniform(-0.05, 1.05, (300, 3)) * (hi - lo)
    value = np.column_stack([xyz, rng.normal(size=300), np.ones(300)])
    fam = po.value_rows(grid, value, 1.0)
    M = fam.matrix(grid.n_nodes)
    _c, _ids, _v, _t, ins = po.locate(grid, xyz)
    cz = np.clip(np.floor((xyz[:, 2] - grid.origin[2]) / grid.step[2]).astype(int), 0, cells[2] - 1)
    for k0, k1 in partition.slab_bounds(nz, world):
        e = Engine(grid.nsteps, grid.origin, grid.step, kind="tetmesh", slab=(k0, k1))
        e.set_p1_regularisation(cgw)
        e.set_options(ridge_factor=ridge, apply_scaling=False)
        p_loc = rng.standard_normal(plane * (k1 - k0))
        p_ext = np.zeros(grid.n_nodes)
        p_ext[k0 * plane:k1 * plane] = p_loc
        sl = slice(k0 * plane, k1 * plane)
        ref = (N @ p_ext + ridge ** 2 * p_ext)[sl]
        q = e.apply_normal(p_loc, parts=en.PART_REGULARISATION | en.PART_RIDGE)
        assert np.abs(q - ref).max() <= 1e-12 * max(np.abs(ref).max(), 1.0), (k0, k1)
        # value rows: the rank keeps the rows whose cell touches its planes and scatters into its own nodes only
        kept = e.add_point_rows(en.ROWS_VALUE, value, 1.0)
        touches = ins & (cz + 1 >= k0) & (cz < k1)
        assert kept == int(touches.sum())
        rows = np.flatnonzero(touches[ins])  # oracle families hold the inside points only
        Ms = M[rows]
        qd = e.apply_normal(p_loc, parts=en.PART_ROWS)
        ref_rows = (Ms.T @ (fam.w[rows] ** 2 * (Ms @ p_ext)))[sl]
        assert np.abs(qd - ref_rows).max() <= 1e-12 * max(np.abs(ref_rows).max(), 1.0), (k0, k1)
        e.close()


def test_p1_device_tests_under_the_generic_plan():
    """LSB200_P1_TILES=0 (k_p1_normal for every node, the kernel the tiled plan falls back to on meshes without an
    interior; read once per process): the P1 device tests again in a process of their own."""
    import os
    import subprocess
    import sys

    if os.environ.get("LSB200_P1_TILES") == "0":
        pytest.skip("already inside the generic-plan run")
    here = os.path.dirname(os.path.abspath(__file__))
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_p1_gpu.py"), os.path.join(here, "test_zz_e_p1_slab_gpu.py"),
                          "-q", "-m", "gpu", "-x"], capture_output=True, text=True, timeout=1500,
                         env=dict(os.environ, LSB200_P1_TILES="0"), cwd=os.path.dirname(here))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
