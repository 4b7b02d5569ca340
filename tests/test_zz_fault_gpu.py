"""Fault-frame evaluation on the CUDA path (csrc/frame.cuh through lsb200_frame_eval / lsb200_fault_displace /
lsb200_fault_displacement) against the golden fixtures of the unmodified reference (tests/golden/fault_*.npz:
FaultSegment.apply_to_points / evaluate_displacement of a faulted GeologicalModel) and against the oracle.

NOTE (round 1): these kernels were written after the round's GPU budget was spent.  Their per-point code is checked
on the CPU by tests/test_frame_hostsim.py (same LSB_HD functions, host instantiation); the launches below have not
run on a device yet, hence the file sorts last."""
import glob
import os

import numpy as np
import pytest

from golden_helpers import GOLDEN, fault_case
from oracle import fault_oracle as fa
from oracle import fdi_oracle as fo

pytestmark = pytest.mark.gpu
CASES = sorted(os.path.basename(p)[6:-4] for p in glob.glob(os.path.join(GOLDEN, "fault_*.npz")))


def _frame(z):
    from loopstructural_b200 import Engine, build
    from loopstructural_b200.fault import FaultFrame

    build.build()
    engines = []
    for i in range(3):
        e = Engine(z[f"f{i}_ncells"] + 1, z[f"f{i}_origin"], z[f"f{i}_step"])
        e.set_c(z[f"f{i}_c"])
        engines.append(e)
    return FaultFrame(engines, [z[f"f{i}_w2l"] for i in range(3)]), engines


def _profile(z):
    from test_frame_hostsim import native_profile

    return native_profile(z)


@pytest.mark.parametrize("tag", CASES)
def test_frame_gathers_match_the_reference(tag):
    z = np.load(os.path.join(GOLDEN, f"fault_{tag}.npz"))
    frame, engines = _frame(z)
    vals, grad = frame.evaluate(z["points"], gradient_of=1)
    np.testing.assert_array_equal(np.isnan(vals), np.isnan(z["frame_values"]))
    np.testing.assert_allclose(vals, z["frame_values"], rtol=0, atol=1e-11, equal_nan=True)
    np.testing.assert_allclose(grad, z["frame_gradient1"], rtol=0, atol=1e-12)  # FMA-contracted sums on the device
    vals2, none = frame.evaluate(z["points"], gradient_of=None)
    assert none is None
    np.testing.assert_array_equal(vals2, vals)
    # the fused pass equals the three separate gathers of the existing evaluation entry points
    w2l = z["f0_w2l"]
    local = z["points"] @ w2l[:3, :3].T + w2l[:3, 3]
    np.testing.assert_allclose(vals[:, 0], engines[0].evaluate_value(local), rtol=0, atol=1e-12, equal_nan=True)
    for e in engines:
        e.close()


@pytest.mark.parametrize("tag", CASES)
def test_apply_to_points_matches_the_reference(tag):
    z = np.load(os.path.join(GOLDEN, f"fault_{tag}.npz"))
    frame, engines = _frame(z)
    pts = z["points"]
    moved = frame.apply_to_points(pts, _profile(z), float(z["displacement"]), int(z["steps"]))
    np.testing.assert_array_equal(np.isnan(moved), np.isnan(z["moved"]))
    still = np.all((z["moved"] == pts) | np.isnan(pts), axis=1)
    np.testing.assert_array_equal(moved[still], pts[still])
    np.testing.assert_allclose(moved, z["moved"], rtol=0, atol=1e-9, equal_nan=True)  # coordinates are O(1e4)
    for e in engines:
        e.close()


@pytest.mark.parametrize("tag", CASES)
def test_evaluate_displacement_matches_the_reference(tag):
    z = np.load(os.path.join(GOLDEN, f"fault_{tag}.npz"))
    frame, engines = _frame(z)
    disp = float(z["displacement"])
    d = frame.evaluate_displacement(z["points"], _profile(z), disp, float(z["fault_offset"]))
    np.testing.assert_allclose(d, z["displacement_at"], rtol=1e-12, atol=1e-12 * abs(disp))
    for e in engines:
        e.close()


def test_large_cloud_device_pointers_and_properties():
    """2e6 points on three 64^3 fields through the device-pointer entry point: zero displacement and zero steps are
    the identity, the moved distance never exceeds |displacement| * max|profile|, and a sample equals the oracle."""
    import torch

    from loopstructural_b200 import Engine, build
    from loopstructural_b200 import fault as fl

    build.build()
    rng = np.random.default_rng(11)
    n_nodes = 64
    grid = fo.Grid.from_cells(np.array([n_nodes - 1] * 3), np.zeros(3), np.full(3, 1.0 / (n_nodes - 1)))
    ax = np.arange(n_nodes) / (n_nodes - 1)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
    flat = lambda a: np.ascontiguousarray(a.transpose(2, 1, 0).ravel())  # noqa: E731  (i fastest)
    cs = [flat(2.0 * (X - 0.5) + 0.2 * np.sin(3 * Y)), flat(Y + 0.3 * Z * Z + 0.1 * np.cos(4 * X)), flat(2.0 * (Z - 0.5) + 0.1 * X * Y)]
    engines = []
    for c in cs:
        e = Engine(grid.nsteps, grid.origin, grid.step)
        e.set_c(c)
        engines.append(e)
    frame = fl.FaultFrame(engines)
    hw = fl.profile1d(fl.PROFILE_CUBIC, (2.0, -3.0, 0.0, 1.0), (0.0, 1.0))
    fw = fl.profile1d(fl.PROFILE_CUBIC, (2.0, 3.0, 0.0, -1.0), (-1.0, 0.0))
    prof = fl.fault_profile([(hw, fw), fl.profile1d(fl.PROFILE_ONES), fl.profile1d(fl.PROFILE_SMOOTH_PEAK)], 0.5)
    n = 2_000_000
    pts = rng.uniform(-0.1, 1.1, (n, 3))
    x = torch.from_numpy(pts).cuda()
    out = torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    frame.apply_to_points_device(x.data_ptr(), out.data_ptr(), n, prof, 0.05, 10, stream=st)
    moved = out.cpu().numpy()
    frame.apply_to_points_device(x.data_ptr(), out.data_ptr(), n, prof, 0.0, 10, stream=st)
    assert torch.equal(out, x)
    frame.apply_to_points_device(x.data_ptr(), out.data_ptr(), n, prof, 0.05, 0, stream=st)
    assert torch.equal(out, x)
    shift = np.linalg.norm(moved - pts, axis=1)
    assert shift.max() <= 0.05 * 0.5 * (1 + 1e-12) and np.sum(shift > 0) > n // 4
    # the host-buffer entry point gives the same bits
    np.testing.assert_array_equal(frame.apply_to_points(pts[:100000], prof, 0.05, 10), moved[:100000])
    fields = [fa.FrameField(grid, c, None) for c in cs]
    oprof = fa.FaultProfile((fa.AxisProfile(fa.Profile1D(fa.CUBIC, (2.0, -3.0, 0.0, 1.0), (0.0, 1.0)),
                                            fa.Profile1D(fa.CUBIC, (2.0, 3.0, 0.0, -1.0), (-1.0, 0.0))),
                             fa.AxisProfile(fa.Profile1D(fa.ONES)), fa.AxisProfile(fa.Profile1D(fa.SMOOTH_PEAK))), 0.5)
    ref = fa.apply_to_points(fields, oprof, 0.05, 10, pts[:20000])
    err = np.abs(moved[:20000] - ref).max(axis=1)
    assert np.sum(err > 1e-13) <= 5 and err.max() < 0.05 / 10  # a cell-face crossing within rounding may differ for a step
    vals = torch.empty((n, 3), dtype=torch.float64, device="cuda")
    grad = torch.empty((n, 3), dtype=torch.float64, device="cuda")
    frame.evaluate_device(x.data_ptr(), vals.data_ptr(), grad.data_ptr(), n, gradient_of=1, stream=st)
    ov, og = fa.frame_eval(fields, pts[:20000], 1)
    np.testing.assert_allclose(vals[:20000].cpu().numpy(), ov, rtol=0, atol=1e-13, equal_nan=True)
    np.testing.assert_allclose(grad[:20000].cpu().numpy(), og, rtol=0, atol=1e-11)
    for e in engines:
        e.close()


def test_argument_errors():
    from loopstructural_b200 import Engine, build
    from loopstructural_b200 import fault as fl
    from loopstructural_b200._native import Lsb200Error

    build.build()
    engines = [Engine([5, 5, 5], [0, 0, 0], [1, 1, 1]) for _ in range(3)]
    frame = fl.FaultFrame(engines)
    with pytest.raises(Lsb200Error):  # no solution on the handles yet
        frame.evaluate(np.zeros((4, 3)))
    for e in engines:
        e.set_c(np.zeros(125))
    bad = fl.fault_profile([fl.profile1d(7)] * 3)
    with pytest.raises(Lsb200Error):
        frame.apply_to_points(np.zeros((4, 3)), bad, 1.0, 2)
    assert frame.apply_to_points(np.zeros((0, 3)), fl.fault_profile(None), 1.0, 2).shape == (0, 3)
    for e in engines:
        e.close()


@pytest.mark.parametrize("tag", ["world", "zero_origin"])
def test_field_orthogonal_rows_match_the_reference(tag):
    """lsb200_add_field_orthogonal_rows against the '<feature>_orthogonal' families of the unmodified reference
    (tests/golden/orth_*.npz): the rows enter the normal operator and the right-hand side exactly as the reference's."""
    from loopstructural_b200 import Engine, build
    from loopstructural_b200 import engine as en
    from test_oracle_golden_orthogonal import orth_case

    build.build()
    z = np.load(os.path.join(GOLDEN, f"orth_{tag}.npz"))
    rng = np.random.default_rng(2)
    for key in z["families"]:
        grid, other, step, w, B, (m_ref, b_ref, w_ref) = orth_case(z, str(key))
        B = 0.37  # the fixtures all have b = 0: a non-zero target also checks the scaling of b
        fam = fa.field_orthogonal_rows(grid, other, step, w, B)
        m_or = fam.matrix(grid.n_nodes)
        assert abs(m_or - m_ref).max() <= 1e-13
        eo = Engine(other.grid.nsteps, other.grid.origin, other.grid.step)
        eo.set_c(other.c)
        e = Engine(grid.nsteps, grid.origin, grid.step)
        e.set_regularisation({})
        kept = e.add_field_orthogonal_rows(eo, step=step, weight=w, b=B, other_world_to_local=other.world_to_local)
        assert kept == m_ref.shape[0] == e.num_rows
        e.finalize()
        p = rng.normal(size=grid.n_nodes)
        q = e.apply_normal(p, parts=en.PART_ROWS)
        q_ref = m_or.T @ (fam.w**2 * (m_or @ p))
        np.testing.assert_allclose(q, q_ref, rtol=0, atol=1e-12 * np.abs(q_ref).max())
        rhs_rows = m_or.T @ (fam.w**2 * fam.b)
        np.testing.assert_allclose(e.rhs(), rhs_rows, rtol=0, atol=1e-12 * max(np.abs(rhs_rows).max(), 1e-300))
        e.close()
        eo.close()


@pytest.mark.parametrize("tag", ["p1_world", "p1_zero_origin"])
def test_p1_field_orthogonal_rows_match_the_reference(tag):
    """The P1 / TetMesh twin of the entry point (target and source on tet meshes, a fault frame built with
    interpolatortype='PLI', tests/golden/orth_p1_*.npz): rows through the normal operator and the right-hand side as the
    reference's; with the dataset's own origin every barycentre is outside the source mesh, the gradient is NaN and the
    whole family is dropped -- by the reference and by the entry point."""
    from loopstructural_b200 import Engine, build
    from loopstructural_b200 import engine as en

    build.build()
    z = np.load(os.path.join(GOLDEN, f"orth_{tag}.npz"))
    rng = np.random.default_rng(3)
    for key in z["families"]:
        key = str(key)
        ti, oj = int(key[1]), int(key[4])
        grid = fo.Grid.from_cells(z[f"f{ti}_ncells"], z[f"f{ti}_origin"], z[f"f{ti}_step"])
        ogrid = fo.Grid.from_cells(z[f"f{oj}_ncells"], z[f"f{oj}_origin"], z[f"f{oj}_step"])
        other = fa.FrameField(ogrid, z[f"f{oj}_c"], z[f"f{oj}_w2l"], "tetmesh")
        step, w = int(z[f"{key}/step"]), float(z[f"{key}/weight"])
        B = 0.37
        fam = fa.p1_field_orthogonal_rows(grid, other, step, w, B)
        eo = Engine(ogrid.nsteps, ogrid.origin, ogrid.step, kind="tetmesh")
        eo.set_c(other.c)
        e = Engine(grid.nsteps, grid.origin, grid.step, kind="tetmesh")
        e.set_p1_regularisation(0.0)
        kept = e.add_field_orthogonal_rows(eo, step=step, weight=w, b=B, other_world_to_local=other.world_to_local)
        if f"{key}/dropped" in z.files:
            assert fam is None and kept == 0 and e.num_rows == 0
        else:
            m_or = fam.matrix(grid.n_nodes)
            assert kept == m_or.shape[0] == int(z[f"{key}/shape"][0]) == e.num_rows
            e.finalize()
            p = rng.normal(size=grid.n_nodes)
            q = e.apply_normal(p, parts=en.PART_ROWS)
            q_ref = m_or.T @ (fam.w**2 * (m_or @ p))
            np.testing.assert_allclose(q, q_ref, rtol=0, atol=1e-12 * max(np.abs(q_ref).max(), 1e-300))
            rhs_rows = m_or.T @ (fam.w**2 * fam.b)
            np.testing.assert_allclose(e.rhs(), rhs_rows, rtol=0, atol=1e-12 * max(np.abs(rhs_rows).max(), 1e-300))
        e.close()
        eo.close()


def test_lattice_export_matches_the_reference_and_the_oracle():
    """lsb200_eval_lattice against BaseFeature.scalar_field of the unmodified reference (lattice_claudius.npz), then
    lattices that reach outside the support, single-point axes, both flattening orders and the device-pointer entry."""
    import torch

    from loopstructural_b200 import Engine, build

    build.build()
    z = np.load(os.path.join(GOLDEN, "lattice_claudius.npz"))
    grid = fo.Grid.from_cells(z["ncells"], z["origin"], z["step"])
    fld = fa.FrameField(grid, z["c"], z["w2l"])
    e = Engine(grid.nsteps, grid.origin, grid.step)
    e.set_c(z["c"])
    lo, hi, n = z["box_origin"], z["box_maximum"], z["box_nsteps"]
    rng_ = float(np.nanmax(z["node_values_f"]) - np.nanmin(z["node_values_f"]))
    kw = dict(world_to_local=z["w2l"])
    np.testing.assert_allclose(e.evaluate_lattice(lo, hi, n, "F", **kw), z["node_values_f"], rtol=0, atol=1e-13 * rng_)
    np.testing.assert_allclose(e.evaluate_lattice(lo, hi, n, "C", **kw), z["node_values_c"], rtol=0, atol=1e-13 * rng_)
    np.testing.assert_allclose(e.evaluate_lattice(lo, hi, n, "F", gradient=True, **kw), z["node_gradient_f"], rtol=0, atol=1e-12)
    ext = hi - lo
    for start, stop, nn in ((lo - 0.7 * ext, hi + 0.7 * ext, (13, 9, 21)), (lo, hi, (1, 7, 5)), (hi, lo, (6, 1, 4))):
        for order in ("F", "C"):
            v = e.evaluate_lattice(start, stop, nn, order, **kw)
            ov = fa.evaluate_lattice(fld, start, stop, nn, order)
            np.testing.assert_array_equal(np.isnan(v), np.isnan(ov))
            np.testing.assert_allclose(v, ov, rtol=0, atol=1e-13 * rng_, equal_nan=True)
            g = e.evaluate_lattice(start, stop, nn, order, gradient=True, **kw)
            np.testing.assert_allclose(g, fa.evaluate_lattice(fld, start, stop, nn, order, gradient=True), rtol=0, atol=1e-12)
    # device-pointer entry point, a lattice of 2.1e7 points (more than one staging chunk on the host path)
    nn = (300, 280, 250)
    total = int(np.prod(nn))
    out = torch.empty(total, dtype=torch.float64, device="cuda")
    e.evaluate_lattice(lo, hi, nn, "F", out_ptr=out.data_ptr(), stream=torch.cuda.current_stream().cuda_stream, **kw)
    torch.cuda.synchronize()
    host = e.evaluate_lattice(lo, hi, nn, "F", **kw)
    np.testing.assert_array_equal(out.cpu().numpy(), host)
    # against the ordinary gather on explicitly built coordinates (every 997-th point)
    pts = fa.lattice_points(lo, hi, nn, "F")[::997]
    local = pts @ z["w2l"][:3, :3].T + z["w2l"][:3, 3]
    np.testing.assert_allclose(host[::997], e.evaluate_value(local), rtol=0, atol=1e-12 * rng_, equal_nan=True)
    e.close()


@pytest.mark.parametrize("tag", ["base", "none", "wide"])
def test_fault_frame_on_a_tet_mesh_matches_the_reference(tag):
    """Frame solved with interpolatortype='PLI' (tests/golden/faultp1_*.npz): the tet-mesh branch of the frame kernels."""
    from loopstructural_b200 import Engine, build
    from loopstructural_b200.fault import FaultFrame

    build.build()
    z = np.load(os.path.join(GOLDEN, f"faultp1_{tag}.npz"))
    engines = []
    for i in range(3):
        e = Engine(z[f"f{i}_ncells"] + 1, z[f"f{i}_origin"], z[f"f{i}_step"], kind="tetmesh")
        e.set_c(z[f"f{i}_c"])
        engines.append(e)
    frame = FaultFrame(engines, [z[f"f{i}_w2l"] for i in range(3)])
    vals, grad = frame.evaluate(z["points"], gradient_of=1)
    np.testing.assert_array_equal(np.isnan(vals), np.isnan(z["frame_values"]))
    np.testing.assert_allclose(vals, z["frame_values"], rtol=0, atol=1e-11, equal_nan=True)
    np.testing.assert_array_equal(np.isnan(grad), np.isnan(z["frame_gradient1"]))
    np.testing.assert_allclose(grad, z["frame_gradient1"], rtol=0, atol=1e-12, equal_nan=True)
    moved = frame.apply_to_points(z["points"], _profile(z), float(z["displacement"]), int(z["steps"]))
    np.testing.assert_array_equal(np.isnan(moved), np.isnan(z["moved"]))
    np.testing.assert_allclose(moved, z["moved"], rtol=0, atol=1e-9, equal_nan=True)
    disp = float(z["displacement"])
    d = frame.evaluate_displacement(z["points"], _profile(z), disp, float(z["fault_offset"]))
    np.testing.assert_allclose(d, z["displacement_at"], rtol=1e-12, atol=1e-12 * abs(disp))
    for e in engines:
        e.close()


def test_fault_chain_equals_the_faults_applied_one_by_one():
    """lsb200_fault_displace_chain (two frames from the golden fixtures, different profiles) against two separate
    lsb200_fault_displace calls: same bits, one transfer."""
    from loopstructural_b200 import fault as fl

    za = np.load(os.path.join(GOLDEN, "fault_base.npz"))
    zb = np.load(os.path.join(GOLDEN, "fault_wide.npz"))
    fa_, ea = _frame(za)
    fb_, eb = _frame(zb)
    pa, pb = _profile(za), _profile(zb)
    pts = za["points"]
    one = fa_.apply_to_points(pts, pa, float(za["displacement"]), int(za["steps"]))
    two = fb_.apply_to_points(one, pb, float(zb["displacement"]), int(zb["steps"]))
    chain = fl.apply_fault_chain([fa_, fb_], [pa, pb], [float(za["displacement"]), float(zb["displacement"])],
                                 [int(za["steps"]), int(zb["steps"])], pts)
    np.testing.assert_array_equal(chain, two)
    np.testing.assert_array_equal(fl.apply_fault_chain([], [], [], [], pts), pts)
    for e in ea + eb:
        e.close()


def test_faulted_evaluation_in_one_call():
    """lsb200_faulted_eval (fault chain, then a field at the moved points, all on the device) against the same thing done
    in pieces through the entry points checked above, for 0, 1 and 2 faults, value and gradient."""
    from loopstructural_b200 import fault as fl

    za = np.load(os.path.join(GOLDEN, "fault_base.npz"))
    zb = np.load(os.path.join(GOLDEN, "fault_wide.npz"))
    fa_, ea = _frame(za)
    fb_, eb = _frame(zb)
    pa, pb = _profile(za), _profile(zb)
    pts = za["points"]
    field, w2l = ea[2], za["f2_w2l"]  # any solved handle serves as the evaluated feature
    args = ([fa_, fb_], [pa, pb], [float(za["displacement"]), float(zb["displacement"])], [int(za["steps"]), int(zb["steps"])])
    for nf in (0, 1, 2):
        sub = tuple(a[:nf] for a in args)
        moved = fl.apply_fault_chain(*sub, pts)
        local = moved @ w2l[:3, :3].T + w2l[:3, 3]
        v = fl.faulted_evaluate(*sub, field, pts, gradient=False, world_to_local=w2l)
        ref = field.evaluate_value(local)
        np.testing.assert_array_equal(np.isnan(v), np.isnan(ref))
        np.testing.assert_allclose(v, ref, rtol=0, atol=1e-12, equal_nan=True)
        g = fl.faulted_evaluate(*sub, field, pts, gradient=True, world_to_local=w2l)
        np.testing.assert_allclose(g, field.evaluate_gradient(local), rtol=0, atol=1e-13)  # identity rotation: local = world axes
    for e in ea + eb:
        e.close()
