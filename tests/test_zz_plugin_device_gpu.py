"""The plug-in inside the reference's own call chain ON THE DEVICE: GeologicalModel.create_and_add_foliation / _fault /
_fold_frame(solver='b200') -> GeologicalFeatureBuilder -> interpolator_map -> B200*Interpolator -> liblsb200.so.
Needs the unmodified reference importable: baseline/_ref (baseline/install_reference.sh; travels to the GPU box with the
snapshot) or the /root/reference mount.  tests/test_plugin_model_flow.py runs the same flows on the CPU with an oracle
stand-in for the engine; here nothing is replaced."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refenv  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refenv.available(), reason="reference not installed (baseline/install_reference.sh)")]


@pytest.fixture()
def ls():
    refenv.activate()
    import LoopStructural
    from loopstructural_b200 import plugin

    plugin.install()
    yield LoopStructural
    plugin.uninstall()


def rel_l2(a, b):
    b = np.asarray(b)
    return float(np.linalg.norm(np.asarray(a) - b) / np.sqrt(b.size) / float(b.max() - b.min()))


def test_readme_quickstart_at_nelements_1e5_runs_on_the_device(ls, golden_dir):
    """BASELINE configs[0] at its stated size: Claudius 'strati' foliation, nelements=1e5 (175,780 nodes), solver='b200',
    against the stock solver's result on the same inputs (golden full_c1_claudius_1e5.npz)."""
    from LoopStructural.datasets import load_claudius

    z = np.load(os.path.join(golden_dir, "full_c1_claudius_1e5.npz"))
    data, bb = load_claudius()
    model = ls.GeologicalModel(bb[0, :], bb[1, :])
    model.data = data
    model.create_and_add_foliation("strati", nelements=1e5, solver="b200")
    model.update()
    it = model["strati"].interpolator
    assert type(it).__name__ == "B200FiniteDifferenceInterpolator"
    assert it.latest_solve_timing["backend"] == "b200" and it.up_to_date
    assert np.array_equal(np.asarray(it.support.nsteps_cells), z["ncells"])
    assert sorted(it.constraints) == ["norm_0", "norm_1", "norm_2", "value"]  # no stencil family was ever assembled
    its = int(it.latest_solve_timing["iterations"])
    d = abs(its - int(z["iterations"]))
    rng_ = float(z["c_range"])
    assert d <= 3, (its, int(z["iterations"]))
    c = np.array(it.c)
    # default stopping tolerance: 2015 iterations, the iterates of two correct CG runs differ by a fraction of one update
    # (3.2e-5 of the range, measured); the converged comparison below holds the north star's 1e-6
    assert np.linalg.norm(c[z["idx"]] - z["c"]) / np.sqrt(z["idx"].size) / rng_ <= 3.2e-5 * (1 + d)
    val = model.evaluate_feature_value("strati", z["query_world"])
    grad = model.evaluate_feature_gradient("strati", z["query_world"])
    np.testing.assert_array_equal(np.isnan(val), np.isnan(z["value_at"]))
    np.testing.assert_allclose(val, z["value_at"], rtol=0, atol=2e-4 * rng_, equal_nan=True)
    gscale = float(np.nanmax(np.abs(z["gradient_at"])))
    np.testing.assert_allclose(grad, z["gradient_at"], rtol=0, atol=2e-3 * gscale, equal_nan=True)
    # both solvers converged (rtol 1e-10) on the same assembled system: the north star's bar
    assert it.solve_system("b200", solver_kwargs={"rtol": 1e-10, "atol": 0.0}) and it.latest_solve_timing["backend"] == "b200"
    assert np.linalg.norm(np.array(it.c)[z["idx"]] - z["c_tight"]) / np.sqrt(z["idx"].size) / rng_ <= 1e-6
    assert abs(int(it.latest_solve_timing["iterations"]) - int(z["iterations_tight"])) <= max(3, int(z["iterations_tight"]) // 100)


def test_fault_frame_and_faulted_foliation_on_the_device(ls):
    """BASELINE configs[3], fault half (load_intrusion): three coupled FDI solves with device-generated orthogonality rows,
    a foliation whose data are un-faulted through the frame, FaultSegment.apply_to_points / evaluate_displacement and the
    faulted evaluation in one device call -- against the stock build of the same model."""
    from test_plugin_model_flow import _fault_model

    from loopstructural_b200 import plugin

    model, bb = _fault_model(ls, solver="b200")
    fault = model["fault"]
    for i in range(3):
        it = fault[i].interpolator
        assert type(it).__name__ == "B200FiniteDifferenceInterpolator" and it.latest_solve_timing["backend"] == "b200"
        assert not any("orthogonal" in k for k in it.constraints)
    assert model["strati"].interpolator.latest_solve_timing["backend"] == "b200"
    assert fault[1].interpolator.latest_solve_timing["matrix_rows"] > sum(c["matrix"].shape[0] for c in fault[1].interpolator.constraints.values())
    rng = np.random.default_rng(5)
    pts = bb[0] + rng.uniform(-0.2, 1.2, (400, 3)) * (bb[1] - bb[0])
    moved = fault.apply_to_points(pts)
    disp = fault.evaluate_displacement(pts)
    strati = model.evaluate_feature_value("strati", pts)
    # (the gradient of a FAULTED feature is not compared: the stock GeologicalFeature.evaluate_gradient itself fails on
    # it -- it reuses `v` for the (n,4) fault values and then assigns the (n,3) gradient, _geological_feature.py:211-228)
    c_frame = [np.array(fault[i].interpolator.c) for i in range(3)]
    plugin.uninstall()
    stock, _ = _fault_model(ls)
    sf = stock["fault"]
    assert sf[0].interpolator.latest_solve_timing["backend"] == "python"
    # each stage is stopped at the default tolerance and feeds the next: a few 1e-6 of the range per stage
    for i in range(3):
        assert rel_l2(c_frame[i], sf[i].interpolator.c) <= 2e-5, i
    ext = float(np.max(bb[1] - bb[0]))
    np.testing.assert_allclose(moved, sf.apply_to_points(pts), rtol=0, atol=2e-5 * ext)
    np.testing.assert_allclose(disp, sf.evaluate_displacement(pts), rtol=0, atol=2e-5 * abs(sf.displacement))
    s_ref = stock.evaluate_feature_value("strati", pts)
    np.testing.assert_array_equal(np.isnan(strati), np.isnan(s_ref))
    np.testing.assert_allclose(strati, s_ref, rtol=0, atol=5e-5 * float(np.nanmax(s_ref) - np.nanmin(s_ref)), equal_nan=True)
    # unfaulted gradients: the fault frame's own coordinates
    g_b200 = [np.asarray(f.evaluate_gradient(pts)) for f in (fault[0], fault[1])]
    for gb, f in zip(g_b200, (sf[0], sf[1])):
        g_ref = np.asarray(f.evaluate_gradient(pts))
        np.testing.assert_allclose(gb, g_ref, rtol=0, atol=1e-3 * float(np.nanmax(np.abs(g_ref))), equal_nan=True)


def test_refolded_model_on_the_device(ls):
    """BASELINE configs[3], fold half (load_laurent2016): fold frame (FDI), folded fold frame and folded foliation
    (DiscreteFoldInterpolator on a TetMesh, fold rows generated on the device) with solver='b200'; every stage against the
    stock solver ON THE SAME SYSTEM (see tests/test_plugin_model_flow.py for why)."""
    from test_plugin_model_flow import _fold_model

    model, bb = _fold_model(ls, solver="b200")
    it0 = model["s0"].interpolator
    assert type(it0).__name__ == "B200DiscreteFoldInterpolator" and it0.latest_solve_timing["backend"] == "b200"
    assert model["s2"][0].interpolator.latest_solve_timing["backend"] == "b200"
    for it in (model["s1"][0].interpolator, it0):
        # the fold rows exist on the device only: generated there from barycentres / faces and the fold vectors
        generated = [f["name"] for f in it._b200_generated]
        assert len(generated) >= 2 and not set(generated) & set(it.constraints)
        c_default = np.array(it.c)
        tol = it.interpolation_weights.get("tol")
        assert it.solve_system("b200", solver_kwargs={"rtol": 1e-11, "atol": 0.0}) and it.latest_solve_timing["backend"] == "b200"
        c_tight = np.array(it.c)
        # a stock solver materialises the same families on the host (the same elements, the same vectors)
        assert it.solve_system("cg", solver_kwargs={"rtol": 1e-11, "atol": 0.0}) and it.latest_solve_timing["backend"] == "python"
        assert it._b200_generated == [] and set(generated) <= set(it.constraints)
        assert rel_l2(c_tight, it.c) <= 1e-6
        assert it.solve_system("cg", tol=tol) is True
        assert rel_l2(c_default, it.c) <= 3e-5  # default tolerance on systems of thousands of iterations


def test_builder_flow_without_solver_at_setup_keeps_every_constraint_on_the_device(ls):
    """loop_interpolation's InterpolatorBuilder pops `solver` before it calls setup: the explicit set-up is solved on
    the device with the default stencil families applied matrix-free and every later constraint kept."""
    import loop_interpolation as li
    from loop_common.supports import StructuredGrid

    from oracle import fdi_oracle as fo

    grid, value, gradient = fo.synthetic_problem(20, 300, 40, seed=3)
    sg = StructuredGrid(origin=grid.origin, nsteps_cells=grid.nsteps - 1, step_vector=grid.step)
    rng = np.random.default_rng(1)
    pts, vec = rng.uniform(0.05, 0.95, (30, 3)), rng.normal(size=(30, 3))

    def solve(cls, solver):
        it = cls(sg)
        it.set_value_constraints(value.copy())
        it.set_normal_constraints(gradient.copy())
        it.setup_interpolator()
        it.add_gradient_orthogonal_constraints(pts.copy(), vec.copy(), w=0.7, b=0.0, name="after_setup")
        assert it.solve_system(solver, solver_kwargs={"rtol": 1e-10, "atol": 0.0})
        return it

    it = solve(li.interpolator_map[li.InterpolatorType.FINITE_DIFFERENCE], "b200")
    assert it.latest_solve_timing["backend"] == "b200"
    n_data = sum(it.constraints[k]["matrix"].shape[0] for k in ("norm_0", "norm_1", "norm_2", "value", "after_setup"))
    assert it.latest_solve_timing["matrix_rows"] == n_data  # the twelve stencil families did not travel as rows
    stock = solve(li.FiniteDifferenceInterpolator, "cg")
    assert rel_l2(it.c, stock.c) <= 1e-6


def test_admm_inner_solves_on_the_device(ls):
    """SURVEY 8 f rank 3: the reference's ADMM loop (inequality constraints) with every inner least-squares solve on the
    device (solver='admm', solver_kwargs={'linsys_solver': 'b200'}); against the stock ADMM with converged CG inner solves."""
    from test_plugin_model_flow import _inequality_problem

    li, make, pts, f = _inequality_problem(ls, n=20)
    cls = li.interpolator_map[li.InterpolatorType.FINITE_DIFFERENCE]
    kw = {"nmajor": 25, "rtol": 1e-13}
    stock = make(li.FiniteDifferenceInterpolator)
    assert stock.solve_system("admm", solver_kwargs=dict(kw, linsys_solver="cg"))
    it = make(cls, solver="b200")
    assert it.solve_system("admm", solver_kwargs=dict(kw, linsys_solver="b200"))
    t = it.latest_solve_timing
    assert t["inner_backend"] == "b200" and t["inner_solves"] == 25 and t["inner_iterations"] > 25
    assert rel_l2(it.c, stock.c) <= 1e-6
    rng_ = float(stock.c.max() - stock.c.min())
    np.testing.assert_allclose(it.evaluate_value(pts), stock.evaluate_value(pts), rtol=0, atol=1e-6 * rng_)
