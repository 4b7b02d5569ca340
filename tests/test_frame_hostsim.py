"""The per-point code of the fault-frame kernels (loopstructural_b200/csrc/frame.cuh) run on the CPU by
tests/hostsim (same LSB_HD functions the kernels call per thread, compiled for the host) against the golden
fixtures of the reference and the oracle.  Complements tests/test_zz_fault_gpu.py, which needs a device."""
import ctypes as C
import glob
import os
import shutil
import subprocess

import numpy as np
import pytest

from golden_helpers import GOLDEN, fault_case
from oracle import fault_oracle as fa

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostsim", "frame_hostsim.cu")
LIB = os.path.join(HERE, "hostsim", "libframe_hostsim.so")
CSRC = os.path.join(os.path.dirname(HERE), "loopstructural_b200", "csrc")
CASES = sorted(os.path.basename(p)[6:-4] for p in glob.glob(os.path.join(GOLDEN, "fault_*.npz")))
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def sim():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    deps = [SRC] + glob.glob(os.path.join(CSRC, "*.cuh"))
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.run([nvcc, "-O2", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xlinker", "-Bsymbolic", "-shared", "-gencode",
                        "arch=compute_100a,code=sm_100a", "-o", LIB, SRC], check=True, capture_output=True)
    return C.CDLL(LIB)


def _frame_args(z):
    ns = np.ascontiguousarray(np.concatenate([z[f"f{i}_ncells"] + 1 for i in range(3)]), dtype=np.int64)
    org = np.ascontiguousarray(np.concatenate([z[f"f{i}_origin"] for i in range(3)]), dtype=np.float64)
    stp = np.ascontiguousarray(np.concatenate([z[f"f{i}_step"] for i in range(3)]), dtype=np.float64)
    cs = [np.ascontiguousarray(z[f"f{i}_c"], dtype=np.float64) for i in range(3)]
    from loopstructural_b200.fault import transform_rows

    tr = np.ascontiguousarray(np.concatenate([transform_rows(z[f"f{i}_w2l"]) for i in range(3)]))
    keep = (ns, org, stp, cs, tr)
    return keep, [ns.ctypes.data_as(C.POINTER(C.c_int64)), org.ctypes.data_as(dp), stp.ctypes.data_as(dp),
                  *[c.ctypes.data_as(dp) for c in cs], tr.ctypes.data_as(dp)]


def native_profile(z):
    from loopstructural_b200 import fault as fl

    if not int(z["has_function"]):
        return fl.fault_profile(None, 0.0)
    tab, comp = z["profile_table"], z["profile_composite"]

    def one(row):
        return fl.profile1d(int(row[0]), row[1:5], row[5:7], row[7], row[8])

    return fl.fault_profile([(one(tab[a, 0]), one(tab[a, 1])) if comp[a] else one(tab[a, 0]) for a in range(3)],
                            float(z["profile_scale"]))


@pytest.mark.parametrize("tag", CASES)
def test_frame_gather_code_matches_the_reference(sim, tag):
    z = np.load(os.path.join(GOLDEN, f"fault_{tag}.npz"))
    keep, args = _frame_args(z)
    pts = np.ascontiguousarray(z["points"])
    n = pts.shape[0]
    vals, grad = np.empty((n, 3)), np.empty((n, 3))
    assert sim.hostsim_frame_eval(*args, 1, C.c_int64(n), pts.ctypes.data_as(dp), vals.ctypes.data_as(dp), grad.ctypes.data_as(dp)) == 0
    np.testing.assert_array_equal(np.isnan(vals), np.isnan(z["frame_values"]))
    np.testing.assert_allclose(vals, z["frame_values"], rtol=0, atol=1e-11, equal_nan=True)
    np.testing.assert_allclose(grad, z["frame_gradient1"], rtol=0, atol=1e-13)


@pytest.mark.parametrize("tag", CASES)
def test_displace_code_matches_the_reference(sim, tag):
    z = np.load(os.path.join(GOLDEN, f"fault_{tag}.npz"))
    keep, args = _frame_args(z)
    prof = native_profile(z)
    pts = np.ascontiguousarray(z["points"])
    n = pts.shape[0]
    moved = np.empty((n, 3))
    rc = sim.hostsim_fault_displace(*args, C.byref(prof), C.c_double(float(z["displacement"])), C.c_int(int(z["steps"])),
                                    C.c_int64(n), pts.ctypes.data_as(dp), moved.ctypes.data_as(dp))
    assert rc == 0
    np.testing.assert_array_equal(np.isnan(moved), np.isnan(z["moved"]))
    # points the reference leaves alone are left alone bit for bit
    still = np.all((z["moved"] == pts) | np.isnan(pts), axis=1)
    np.testing.assert_array_equal(moved[still], pts[still])
    np.testing.assert_allclose(moved, z["moved"], rtol=0, atol=1e-9, equal_nan=True)  # coordinates are O(1e4)


@pytest.mark.parametrize("tag", CASES)
def test_displacement_code_matches_the_reference(sim, tag):
    z = np.load(os.path.join(GOLDEN, f"fault_{tag}.npz"))
    keep, args = _frame_args(z)
    prof = native_profile(z)
    pts = np.ascontiguousarray(z["points"])
    n = pts.shape[0]
    d = np.empty(n)
    disp = float(z["displacement"])
    rc = sim.hostsim_fault_displacement(*args, C.byref(prof), C.c_double(disp), C.c_double(float(z["fault_offset"])),
                                        C.c_int64(n), pts.ctypes.data_as(dp), d.ctypes.data_as(dp))
    assert rc == 0
    np.testing.assert_allclose(d, z["displacement_at"], rtol=1e-12, atol=1e-12 * abs(disp))


def test_rotated_frames_and_distinct_supports_match_the_oracle(sim):
    """Three different supports, each with its own rotated world -> local transform (the fixtures of the reference
    all use translation only): the per-point code against the oracle."""
    rng = np.random.default_rng(3)
    fields, ns, org, stp, cs, trs = [], [], [], [], [], []
    from loopstructural_b200.fault import transform_rows
    from oracle import fdi_oracle as fo

    for f, cells in enumerate(((9, 7, 8), (6, 11, 5), (8, 8, 8))):
        ang = 0.3 + 0.4 * f
        rot = np.array([[np.cos(ang), -np.sin(ang), 0.0], [np.sin(ang), np.cos(ang), 0.0], [0.0, 0.0, 1.0]])
        if f == 2:
            rot = rot[[2, 0, 1]]
        w2l = np.eye(4)
        w2l[:3, :3] = rot
        w2l[:3, 3] = -rot @ np.array([0.5, -0.25, 0.1 * f])
        grid = fo.Grid.from_cells(np.array(cells), np.array([-1.0, -1.2, -0.9]), 2.4 / np.array(cells))
        xyz = np.stack(np.meshgrid(*[grid.origin[a] + grid.step[a] * np.arange(grid.nsteps[a]) for a in range(3)], indexing="ij"), -1)
        xyz = xyz.transpose(2, 1, 0, 3).reshape(-1, 3)  # flat index i + nx j + nx ny k
        c = (xyz @ rng.normal(size=3)) + 0.3 * np.sin(2.0 * xyz[:, f]) + 0.1 * rng.normal(size=xyz.shape[0])
        fields.append(fa.FrameField(grid, c, w2l))
        ns.append(grid.nsteps); org.append(grid.origin); stp.append(grid.step); cs.append(np.ascontiguousarray(c)); trs.append(transform_rows(w2l))
    ns = np.ascontiguousarray(np.concatenate(ns), dtype=np.int64)
    org, stp, tr = (np.ascontiguousarray(np.concatenate(a), dtype=np.float64) for a in (org, stp, trs))
    args = [ns.ctypes.data_as(C.POINTER(C.c_int64)), org.ctypes.data_as(dp), stp.ctypes.data_as(dp),
            *[c.ctypes.data_as(dp) for c in cs], tr.ctypes.data_as(dp)]
    pts = np.ascontiguousarray(rng.uniform(-1.6, 1.6, (5000, 3)))
    n = pts.shape[0]
    from loopstructural_b200 import fault as fl

    cub = fl.profile1d(fl.PROFILE_CUBIC, (0.3, -0.2, 0.5, 0.1), (-1.5, 1.2), -0.4, 0.9)
    prof = fl.fault_profile([(cub, fl.profile1d(fl.PROFILE_SMOOTH_PEAK)), fl.profile1d(fl.PROFILE_ONES), cub], 0.7)
    oprof = fa.FaultProfile((fa.AxisProfile(fa.Profile1D(fa.CUBIC, (0.3, -0.2, 0.5, 0.1), (-1.5, 1.2), -0.4, 0.9), fa.Profile1D(fa.SMOOTH_PEAK)),
                             fa.AxisProfile(fa.Profile1D(fa.ONES)),
                             fa.AxisProfile(fa.Profile1D(fa.CUBIC, (0.3, -0.2, 0.5, 0.1), (-1.5, 1.2), -0.4, 0.9))), 0.7)
    vals, grad = np.empty((n, 3)), np.empty((n, 3))
    for gof in (0, 2):
        assert sim.hostsim_frame_eval(*args, gof, C.c_int64(n), pts.ctypes.data_as(dp), vals.ctypes.data_as(dp), grad.ctypes.data_as(dp)) == 0
        ov, og = fa.frame_eval(fields, pts, gradient_of=gof)
        np.testing.assert_allclose(vals, ov, rtol=0, atol=1e-13, equal_nan=True)
        np.testing.assert_allclose(grad, og, rtol=0, atol=1e-12)
    moved = np.empty((n, 3))
    assert sim.hostsim_fault_displace(*args, C.byref(prof), C.c_double(0.35), C.c_int(7), C.c_int64(n), pts.ctypes.data_as(dp),
                                      moved.ctypes.data_as(dp)) == 0
    ref = fa.apply_to_points(fields, oprof, 0.35, 7, pts)
    assert np.sum(np.any(ref != pts, axis=1)) > 500
    # a point that crosses a cell face within rounding of the reference's path may take the neighbour's gradient for
    # one step: allow a handful of such points, hold all others to 1e-12
    err = np.abs(moved - ref).max(axis=1)
    assert np.sum(err > 1e-12) <= 3 and err.max() < 0.35 / 7
    d = np.empty(n)
    assert sim.hostsim_fault_displacement(*args, C.byref(prof), C.c_double(0.35), C.c_double(0.2), C.c_int64(n),
                                          pts.ctypes.data_as(dp), d.ctypes.data_as(dp)) == 0
    np.testing.assert_allclose(d, fa.evaluate_displacement(fields, oprof, 0.35, 0.2, pts), rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("tag", ["world", "zero_origin"])
def test_field_orthogonal_row_code_matches_the_reference(sim, tag):
    """Body of k_fdi_field_orthogonal_rows against the '<feature>_orthogonal' families of the unmodified reference."""
    from scipy import sparse

    from loopstructural_b200.fault import transform_rows
    from test_oracle_golden_orthogonal import orth_case

    z = np.load(os.path.join(GOLDEN, f"orth_{tag}.npz"))
    for key in z["families"]:
        grid, other, step, w, B, (m_ref, b_ref, w_ref) = orth_case(z, str(key))
        ns = np.ascontiguousarray(np.concatenate([grid.nsteps, other.grid.nsteps, [2, 2, 2]]), dtype=np.int64)
        org = np.ascontiguousarray(np.concatenate([grid.origin, other.grid.origin, np.zeros(3)]))
        stp = np.ascontiguousarray(np.concatenate([grid.step, other.grid.step, np.ones(3)]))
        tr = np.ascontiguousarray(np.concatenate([transform_rows(None), transform_rows(other.world_to_local), transform_rows(None)]))
        c_other = np.ascontiguousarray(other.c)
        n_rows = m_ref.shape[0]
        vals, cols = np.empty((n_rows, 8)), np.empty((n_rows, 8), dtype=np.int64)
        b, w2 = np.empty(n_rows), np.empty(n_rows)
        kept = sim.hostsim_field_orthogonal_rows(ns.ctypes.data_as(C.POINTER(C.c_int64)), org.ctypes.data_as(dp), stp.ctypes.data_as(dp),
                                                 c_other.ctypes.data_as(dp), tr.ctypes.data_as(dp), C.c_int64(step), C.c_double(w),
                                                 C.c_double(B), C.c_int64(n_rows), vals.ctypes.data_as(dp),
                                                 cols.ctypes.data_as(C.POINTER(C.c_int64)), b.ctypes.data_as(dp), w2.ctypes.data_as(dp))
        assert kept == n_rows and cols.min() >= 0
        m = sparse.coo_matrix((vals.ravel(), (np.repeat(np.arange(n_rows), 8), cols.ravel())), shape=m_ref.shape).tocsr()
        assert abs(m - m_ref).max() <= 1e-13
        np.testing.assert_allclose(b, b_ref, rtol=0, atol=1e-14)
        np.testing.assert_array_equal(w2, w_ref**2)


def _lattice_call(sim, fld, start, stop, n, order, gradient):
    from loopstructural_b200.fault import transform_rows

    ns = np.ascontiguousarray(np.concatenate([fld.grid.nsteps, [2, 2, 2], [2, 2, 2]]), dtype=np.int64)
    org = np.ascontiguousarray(np.concatenate([fld.grid.origin, np.zeros(6)]))
    stp = np.ascontiguousarray(np.concatenate([fld.grid.step, np.ones(6)]))
    tr = np.ascontiguousarray(np.concatenate([transform_rows(fld.world_to_local), transform_rows(None), transform_rows(None)]))
    c = np.ascontiguousarray(fld.c)
    start, stop = np.ascontiguousarray(start, dtype=np.float64), np.ascontiguousarray(stop, dtype=np.float64)
    n = np.ascontiguousarray(n, dtype=np.int64)
    total = int(np.prod(n))
    out = np.empty((total, 3) if gradient else total)
    ip = C.POINTER(C.c_int64)
    rc = sim.hostsim_eval_lattice(ns.ctypes.data_as(ip), org.ctypes.data_as(dp), stp.ctypes.data_as(dp), c.ctypes.data_as(dp),
                                  tr.ctypes.data_as(dp), start.ctypes.data_as(dp), stop.ctypes.data_as(dp), n.ctypes.data_as(ip),
                                  C.c_int(order == "F"), C.c_int(int(gradient)), out.ctypes.data_as(dp))
    assert rc == 0
    return out


def test_lattice_code_matches_the_reference_and_the_oracle(sim):
    """Body of k_eval_lattice: BaseFeature.scalar_field of the unmodified reference (lattice_claudius.npz), then
    lattices that reach outside the support, single-point axes and both flattening orders against the oracle."""
    from oracle import fdi_oracle as fo

    z = np.load(os.path.join(GOLDEN, "lattice_claudius.npz"))
    fld = fa.FrameField(fo.Grid.from_cells(z["ncells"], z["origin"], z["step"]), z["c"], z["w2l"])
    lo, hi, n = z["box_origin"], z["box_maximum"], z["box_nsteps"]
    rng_ = float(np.nanmax(z["node_values_f"]) - np.nanmin(z["node_values_f"]))
    np.testing.assert_allclose(_lattice_call(sim, fld, lo, hi, n, "F", False), z["node_values_f"], rtol=0, atol=1e-13 * rng_)
    np.testing.assert_allclose(_lattice_call(sim, fld, lo, hi, n, "C", False), z["node_values_c"], rtol=0, atol=1e-13 * rng_)
    np.testing.assert_allclose(_lattice_call(sim, fld, lo, hi, n, "F", True), z["node_gradient_f"], rtol=0, atol=1e-12)
    ext = hi - lo
    for start, stop, nn in ((lo - 0.7 * ext, hi + 0.7 * ext, (13, 9, 21)), (lo, hi, (1, 7, 5)), (hi, lo, (6, 1, 4)),
                            (lo + 0.1 * ext, lo + 0.1 * ext, (3, 3, 3))):
        for order in ("F", "C"):
            v = _lattice_call(sim, fld, start, stop, nn, order, False)
            ov = fa.evaluate_lattice(fld, start, stop, nn, order)
            np.testing.assert_array_equal(np.isnan(v), np.isnan(ov))
            np.testing.assert_allclose(v, ov, rtol=0, atol=1e-13 * rng_, equal_nan=True)
            g = _lattice_call(sim, fld, start, stop, nn, order, True)
            np.testing.assert_allclose(g, fa.evaluate_lattice(fld, start, stop, nn, order, gradient=True), rtol=0, atol=1e-12)
    assert np.isnan(_lattice_call(sim, fld, lo - 0.7 * ext, hi + 0.7 * ext, (13, 9, 21), "F", False)).sum() > 500


P1_CASES = sorted(os.path.basename(p)[8:-4] for p in glob.glob(os.path.join(GOLDEN, "faultp1_*.npz")))


@pytest.mark.parametrize("tag", P1_CASES)
def test_tet_mesh_frame_code_matches_the_reference(sim, tag):
    """The tet-mesh branch of the frame kernels (P1 barycentrics / tet gradients, NaN outside for value and gradient)
    against the reference's FaultSegment on a frame solved with interpolatortype='PLI'."""
    z = np.load(os.path.join(GOLDEN, f"faultp1_{tag}.npz"))
    keep, args = _frame_args(z)
    prof = native_profile(z)
    pts = np.ascontiguousarray(z["points"])
    n = pts.shape[0]
    sim.hostsim_set_kinds(1, 1, 1)
    try:
        vals, grad = np.empty((n, 3)), np.empty((n, 3))
        assert sim.hostsim_frame_eval(*args, 1, C.c_int64(n), pts.ctypes.data_as(dp), vals.ctypes.data_as(dp), grad.ctypes.data_as(dp)) == 0
        np.testing.assert_array_equal(np.isnan(vals), np.isnan(z["frame_values"]))
        np.testing.assert_allclose(vals, z["frame_values"], rtol=0, atol=1e-11, equal_nan=True)
        np.testing.assert_array_equal(np.isnan(grad), np.isnan(z["frame_gradient1"]))
        np.testing.assert_allclose(grad, z["frame_gradient1"], rtol=0, atol=1e-12, equal_nan=True)
        moved = np.empty((n, 3))
        assert sim.hostsim_fault_displace(*args, C.byref(prof), C.c_double(float(z["displacement"])), C.c_int(int(z["steps"])),
                                          C.c_int64(n), pts.ctypes.data_as(dp), moved.ctypes.data_as(dp)) == 0
        np.testing.assert_array_equal(np.isnan(moved), np.isnan(z["moved"]))
        np.testing.assert_allclose(moved, z["moved"], rtol=0, atol=1e-9, equal_nan=True)
        d = np.empty(n)
        disp = float(z["displacement"])
        assert sim.hostsim_fault_displacement(*args, C.byref(prof), C.c_double(disp), C.c_double(float(z["fault_offset"])),
                                              C.c_int64(n), pts.ctypes.data_as(dp), d.ctypes.data_as(dp)) == 0
        np.testing.assert_allclose(d, z["displacement_at"], rtol=1e-12, atol=1e-12 * abs(disp))
    finally:
        sim.hostsim_set_kinds(0, 0, 0)
