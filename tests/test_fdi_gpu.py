"""Parity of the CUDA FDI path (through the C-ABI) with the oracle and the reference's golden outputs."""
import glob
import os

import numpy as np
import pytest

from oracle import fdi_oracle as fo
import golden_helpers as H

pytestmark = pytest.mark.gpu
GOLDEN = H.GOLDEN


@pytest.fixture(scope="module", autouse=True)
def _built():
    from loopstructural_b200 import build

    build.build()


# ---------------------------------------------------------------- K1: regularisation operator
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "fdi_operator_*.npz"))))
def test_stencil_operator_matches_reference_gram(path):
    """R^T R p at EVERY dof incl. the boundary shell == the explicit Gram matrix of the reference
    (the bar of LIT/test_fdi_matrix_free_regularisation.py:468-493, 1e-10)."""
    z = np.load(path)
    grid = fo.Grid(z["nodes"], np.zeros(3), np.ones(3))
    w = H.weights_of(z)
    e = H.engine_for(grid, w)
    scale = np.abs(z["y"]).max()
    for x, y in zip(z["x"], z["y"]):
        q = e.apply_normal(x, parts=1)
        assert np.abs(q - y).max() <= 1e-12 * scale
    # ridge part: + ridge^2 p
    e.set_options(ridge_factor=0.5)
    q = e.apply_normal(z["x"][0], parts=1 | 4)
    assert np.abs(q - (z["y"][0] + 0.25 * z["x"][0])).max() <= 1e-12 * scale
    e.set_options(ridge_factor=0.0, add_ridge=False)
    assert np.abs(e.diagonal() - z["diag"]).max() <= 1e-13
    e.close()


@pytest.mark.parametrize("nodes", [(2, 2, 2), (2, 5, 3), (3, 3, 3), (4, 2, 7), (70, 37, 45), (129, 20, 66), (64, 64, 40)])
@pytest.mark.parametrize("cz", ["0", "5"])
def test_stencil_operator_matches_oracle_two_pass(nodes, cz, monkeypatch):
    """Ragged / tiny / multi-tile grids and odd z-chunking against the oracle's masked two-pass form."""
    if cz != "0":
        monkeypatch.setenv("LSB200_STENCIL_CZ", cz)
    rng = np.random.default_rng(sum(nodes))
    grid = fo.Grid(np.array(nodes), np.zeros(3), np.ones(3))
    w = dict(dxx=1.3, dyy=0.8, dzz=1.1, dxy=2.0, dyz=0.5, dxz=1.5, dx=0.7, dy=1.9, dz=1.2)
    e = H.engine_for(grid, w)
    p = rng.standard_normal(grid.n_nodes)
    ref = fo.regularisation_normal_matvec(grid, w, p)
    q = e.apply_normal(p, parts=1)
    assert np.abs(q - ref).max() <= 1e-12 * max(np.abs(ref).max(), 1.0)
    e.set_options(add_ridge=False)
    assert np.abs(e.diagonal() - fo.regularisation_normal_diagonal(grid, w)).max() <= 1e-13
    e.close()


@pytest.mark.parametrize("nodes", [(70, 37, 45), (129, 20, 66), (64, 64, 40), (16, 16, 5), (200, 100, 9), (66, 70, 130), (8, 16, 3)])
@pytest.mark.parametrize("main", ["normal", "fused", "fused3", "fused4", "fused-cp.async", "fused3-cp.async"])
@pytest.mark.parametrize("cz", ["0", "7"])
def test_main_stencil_kernel_configs(nodes, main, cz, monkeypatch):
    """Every configuration of the single-launch operator (k_fdi_fused: both tile shapes, TMA and cp.async staging)
    against the oracle's masked two-pass form, on ragged grids and odd z-chunking; "normal" is k_fdi_normal everywhere."""
    if main == "normal":
        monkeypatch.setenv("LSB200_STENCIL_MAIN", "-1")
    for tyt in "34":
        if tyt in main:
            monkeypatch.setenv("LSB200_FUSED_CFG", tyt)
    if "cp.async" in main:
        monkeypatch.setenv("LSB200_TMA", "0")
    if cz != "0":
        monkeypatch.setenv("LSB200_STENCIL_CZ", cz)
        monkeypatch.setenv("LSB200_FUSED_CZ", cz)
        monkeypatch.setenv("LSB200_FUSED_SCZ", "9")
    rng = np.random.default_rng(sum(nodes) + 1)
    grid = fo.Grid(np.array(nodes), np.zeros(3), np.ones(3))
    w = dict(dxx=1.3, dyy=0.8, dzz=1.1, dxy=2.0, dyz=0.5, dxz=1.5, dx=0.7, dy=1.9, dz=1.2)
    e = H.engine_for(grid, w)
    e.set_options(ridge_factor=0.3)
    p = rng.standard_normal(grid.n_nodes)
    ref = fo.regularisation_normal_matvec(grid, w, p) + 0.09 * p
    q = e.apply_normal(p, parts=1 | 4)
    assert np.abs(q - ref).max() <= 1e-12 * max(np.abs(ref).max(), 1.0)
    e.close()


def test_stencil_operator_full_size_properties():
    """BASELINE config 2 size (256^3 nodes): the regularisation annihilates affine fields, is symmetric,
    and agrees with the oracle on a z-subsample of planes recomputed from a slab."""
    n = 256
    grid = fo.Grid(np.array([n, n, n]), np.zeros(3), np.ones(3) / (n - 1))
    e = H.engine_for(grid, {})
    e.set_options(add_ridge=False)
    i, j, k = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    lin = (0.3 * i - 1.7 * j + 0.9 * k + 4.0).ravel(order="F")
    q = e.apply_normal(lin, parts=1)
    assert np.abs(q).max() <= 1e-9 * np.abs(lin).max()
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal(n ** 3), rng.standard_normal(n ** 3)
    qa, qb = e.apply_normal(a, parts=1), e.apply_normal(b, parts=1)
    assert abs(np.dot(b, qa) - np.dot(a, qb)) <= 1e-10 * abs(np.dot(b, qa))
    # oracle on the first and last 12 planes (boundary handling) using only the planes they depend on
    m = 12
    sub = fo.Grid(np.array([n, n, m + 2]), np.zeros(3), np.ones(3))
    lo = fo.regularisation_normal_matvec(sub, {}, a[: n * n * (m + 2)])
    # planes 0..m-1 of the slab only depend on planes 0..m+1, but the slab's own upper boundary differs
    assert np.abs(qa[: n * n * (m - 2)] - lo[: n * n * (m - 2)]).max() <= 1e-12 * np.abs(lo).max()
    e.close()


# ---------------------------------------------------------------- K2/K3: data rows
@pytest.mark.parametrize("tag", ["a", "b"])
def test_point_rows_match_reference(tag):
    """Device-generated value / gradient / normal / tangent rows == the reference's constraint
    matrices (fixture from the unmodified reference), compared through A^T W A p and A^T W b."""
    from loopstructural_b200 import engine as en

    z = np.load(os.path.join(GOLDEN, f"fdi_rows_{tag}.npz"))
    grid = fo.Grid.from_cells(z["ncells"], z["origin"], z["step"])
    kw = H.weights_of(z, "kw_keys", "kw_vals")
    dof = grid.n_nodes
    rng = np.random.default_rng(3)
    p = rng.standard_normal(dof)
    kinds = {
        "value": (en.ROWS_VALUE, z["value"], kw["cpw"], ["value"]),
        "gradient": (en.ROWS_GRADIENT, z["gradient"], kw["gpw"], ["gradient", "gradient_0"]),
        "normal": (en.ROWS_NORMAL, z["normal"], kw["npw"], ["norm_0", "norm_1", "norm_2"]),
        "tangent": (en.ROWS_TANGENT, np.column_stack([z["tangent"][:, :6], np.zeros(len(z["tangent"]))]), kw["tpw"],
                    ["gradient orthogonal"]),
    }
    for label, (kind, data, weight, names) in kinds.items():
        e = H.engine_for(grid, dict.fromkeys(("dxx", "dyy", "dzz", "dxy", "dyz", "dxz", "dx", "dy", "dz"), 0.0))
        e.set_options(add_ridge=False, apply_scaling=False)
        nrows = e.add_point_rows(kind, data, weight)
        q_ref = np.zeros(dof)
        rhs_ref = np.zeros(dof)
        diag_ref = np.zeros(dof)
        n_ref = 0
        for name in names:
            m, b, w = H.ref_family(z, name)
            n_ref += m.shape[0]
            q_ref += m.T @ (w * w * (m @ p))
            rhs_ref += m.T @ (w * w * b)
            diag_ref += np.asarray(m.multiply(m).T @ (w * w)).ravel()
        assert nrows == n_ref, label
        q = e.apply_normal(p, parts=2)
        s = max(np.abs(q_ref).max(), 1e-300)
        assert np.abs(q - q_ref).max() <= 1e-12 * s, label
        assert np.abs(e.rhs() - rhs_ref).max() <= 1e-12 * max(np.abs(rhs_ref).max(), 1.0), label
        assert np.abs(e.diagonal() - diag_ref).max() <= 1e-12 * max(np.abs(diag_ref).max(), 1.0), label
        e.close()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_generic_rows_match_reference_including_interface(tag):
    """add_rows (= add_constraints_to_least_squares) with the 16-wide interface rows, whose listed
    duplicate columns merge in the reference (LI/_discrete_interpolator.py:384-386)."""
    z = np.load(os.path.join(GOLDEN, f"fdi_rows_{tag}.npz"))
    grid = fo.Grid.from_cells(z["ncells"], z["origin"], z["step"])
    kw = H.weights_of(z, "kw_keys", "kw_vals")
    dof = grid.n_nodes
    fams = fo.interface_rows(grid, z["interface"], kw["ipw"])
    assert fams
    e = H.engine_for(grid, dict.fromkeys(("dxx", "dyy", "dzz", "dxy", "dyz", "dxz", "dx", "dy", "dz"), 0.0))
    e.set_options(add_ridge=False, apply_scaling=False)
    rng = np.random.default_rng(4)
    p = rng.standard_normal(dof)
    q_ref = np.zeros(dof)
    diag_ref = np.zeros(dof)
    for f in fams:
        # un-normalised rows go in; the device normalises over the listed entries
        scale = rng.uniform(0.5, 2.0, f.A.shape[0])
        e.add_rows(f.A * scale[:, None], f.idx, f.b * scale, f.w, normalise=True)
        m, b, w = H.ref_family(z, f.name)
        q_ref += m.T @ (w * w * (m @ p))
        diag_ref += np.asarray(m.multiply(m).T @ (w * w)).ravel()
    q = e.apply_normal(p, parts=2)
    assert np.abs(q - q_ref).max() <= 1e-12 * np.abs(q_ref).max()
    assert np.abs(e.diagonal() - diag_ref).max() <= 1e-12 * np.abs(diag_ref).max()
    e.close()


def test_rows_edge_cases():
    grid = fo.Grid(np.array([5, 4, 6]), np.zeros(3), np.ones(3))
    e = H.engine_for(grid, {})
    from loopstructural_b200 import engine as en

    # empty input, all points outside, NaN rows
    assert e.add_point_rows(en.ROWS_VALUE, np.zeros((0, 5)), 1.0) == 0
    out = np.array([[-1.0, 0.5, 0.5, 1.0, 1.0], [0.5, 9.0, 0.5, 1.0, 1.0]])
    assert e.add_point_rows(en.ROWS_VALUE, out, 1.0) == 0
    nanrow = np.array([[0.5, 0.5, 0.5, np.nan, 1.0], [1.5, 0.5, 0.5, 2.0, 1.0]])
    assert e.add_point_rows(en.ROWS_VALUE, nanrow, 1.0) == 1
    assert e.num_rows == 1
    with pytest.raises(Exception):
        e.add_rows(np.ones((1, 2)), np.array([[0, 10 ** 6]]), 0.0, 1.0)
    e.clear_rows()
    assert e.num_rows == 0
    e.close()


# ---------------------------------------------------------------- K6: CG solve
@pytest.mark.parametrize("tag", ["unit16", "unit16_noscale", "aniso", "unit24_tight"])
def test_solve_matches_reference(tag):
    """Solved coefficients vs the reference's scipy path on the same inputs: relative L2 of the field
    <= 1e-6 of its range (BASELINE north_star), same stopping rule => iteration count within +-1."""
    z = np.load(os.path.join(GOLDEN, f"fdi_solve_{tag}.npz"))
    grid, weights, fams, rtol, atol = H.solve_case(z)
    e = H.engine_for(grid, weights)
    e.set_options(apply_scaling=bool(z["scaling"]))
    H.add_point_data(e, z, weights)
    c, res = e.solve_cg(rtol=rtol, atol=atol)
    assert res.info == 0
    assert abs(res.iterations - int(z["iterations"])) <= 1
    assert res.residual_norm < res.threshold
    rng_ = float(z["c"].max() - z["c"].min())
    rel = np.linalg.norm(c - z["c"]) / np.sqrt(c.size) / rng_
    assert rel <= 1e-6, rel
    # the oracle's matrix-free restatement follows the same trajectory: rhs, diagonal, scaling agree tightly
    mf = fo.MatrixFreeSystem(grid, fams, weights, apply_scaling=bool(z["scaling"]))
    assert np.abs(e.rhs() - mf.rhs_unscaled).max() <= 1e-12 * np.abs(mf.rhs_unscaled).max()
    assert np.abs(e.diagonal() - mf.diag).max() <= 1e-12 * np.abs(mf.diag).max()
    if mf.S is not None:
        assert np.abs(e.scaling() - mf.S).max() <= 1e-12 * np.abs(mf.S).max()
    e.close()


def test_solve_maxiter_x0_and_zero_rhs():
    z = np.load(os.path.join(GOLDEN, "fdi_solve_unit16.npz"))
    grid, weights, fams, rtol, atol = H.solve_case(z)
    e = H.engine_for(grid, weights)
    H.add_point_data(e, z, weights)
    mf = fo.MatrixFreeSystem(grid, fams, weights)
    # iteration cap: scipy convention info = maxiter, last iterate returned
    c7, res7 = e.solve_cg(rtol=1e-14, maxiter=7)
    ref7 = fo.cg(mf.matvec, mf.rhs, rtol=1e-14, maxiter=7)
    assert res7.info == 7 and res7.iterations == 7
    assert np.abs(c7 - mf.S * ref7.x).max() <= 1e-9 * np.abs(ref7.x).max()
    # warm start from the converged solution: zero iterations
    c, res = e.solve_cg(rtol=1e-10)
    c2, res2 = e.solve_cg(rtol=1e-5, x0=c)
    assert res2.iterations == 0 and np.array_equal(c2, c)
    # no data rows => zero right-hand side => x = 0 immediately (scipy returns b)
    e.clear_rows()
    c0, res0 = e.solve_cg()
    assert res0.iterations == 0 and res0.info == 0 and not c0.any()
    e.close()


# ---------------------------------------------------------------- K7: evaluation
@pytest.mark.parametrize("tag", ["unit16", "aniso"])
def test_evaluate_matches_reference(tag):
    z = np.load(os.path.join(GOLDEN, f"fdi_solve_{tag}.npz"))
    grid = fo.Grid.from_cells(z["ncells"], z["origin"], z["step"])
    e = H.engine_for(grid, {})
    e.set_c(z["c"])
    q = z["query"]
    v = e.evaluate_value(q)
    np.testing.assert_array_equal(np.isnan(v), np.isnan(z["value_at"]))
    np.testing.assert_allclose(v, z["value_at"], rtol=0, atol=1e-13, equal_nan=True)
    assert v[3] == 0.0  # NaN-coordinate row
    fin = np.all(np.isfinite(q), axis=1)
    g = e.evaluate_gradient(q[fin])
    np.testing.assert_allclose(g, z["gradient_at"], rtol=0, atol=1e-11)
    # NaN coordinates in the gradient query are "outside": zeros
    assert not e.evaluate_gradient(q[3:4]).any()
    assert e.evaluate_value(np.zeros((0, 3))).shape == (0,)
    e.close()


def test_evaluate_large_cloud_against_oracle():
    rng = np.random.default_rng(11)
    grid = fo.Grid.from_cells([40, 33, 27], [1.0, -2.0, 0.5], [0.3, 0.2, 0.45])
    c = rng.standard_normal(grid.n_nodes)
    e = H.engine_for(grid, {})
    e.set_c(c)
    lo, hi = grid.origin, grid.maximum
    pts = lo + rng.uniform(-0.02, 1.02, (200_000, 3)) * (hi - lo)
    pts[:8] = [[lo[0], lo[1], lo[2]], [hi[0], hi[1], hi[2]], [hi[0], lo[1], hi[2]], [lo[0], hi[1], lo[2]],
               [hi[0], 0.5 * (lo[1] + hi[1]), lo[2]], [lo[0] + 0.3 * 7, lo[1] + 0.2 * 5, lo[2] + 0.45 * 3],
               [hi[0] + 1e-18, hi[1], hi[2]], [lo[0] - 1e-3, lo[1], lo[2]]]
    v, vo = e.evaluate_value(pts), fo.evaluate_value(grid, pts, c)
    np.testing.assert_array_equal(np.isnan(v), np.isnan(vo))
    np.testing.assert_allclose(v, vo, rtol=0, atol=1e-13, equal_nan=True)
    g, go = e.evaluate_gradient(pts), fo.evaluate_gradient(grid, pts, c)
    np.testing.assert_allclose(g, go, rtol=0, atol=1e-11)
    # LCT/test_structured_grid_boundary_eval.py: evaluation ON the max faces is finite and exact
    lin = (np.arange(grid.n_nodes) % grid.nsteps[0]).astype(float)  # field = i index
    e.set_c(lin)
    face = np.array([[hi[0], 0.5 * (lo[1] + hi[1]), 0.5 * (lo[2] + hi[2])]])
    assert abs(e.evaluate_value(face)[0] - (grid.nsteps[0] - 1)) <= 1e-12
    e.close()


# ---------------------------------------------------------------- host-side mirror of the reference API
@pytest.mark.parametrize("tag", ["a", "b"])
def test_interface_rows_generated_on_device_match_reference(tag):
    """lsb200_add_pair_rows == the reference's interface_<id> families (16 listed entries, pairing on the host)."""
    z = np.load(os.path.join(GOLDEN, f"fdi_rows_{tag}.npz"))
    grid = fo.Grid.from_cells(z["ncells"], z["origin"], z["step"])
    kw = H.weights_of(z, "kw_keys", "kw_vals")
    e = H.engine_for(grid, dict.fromkeys(("dxx", "dyy", "dzz", "dxy", "dyz", "dxz", "dx", "dy", "dz"), 0.0))
    e.set_options(add_ridge=False, apply_scaling=False)
    n = e.add_interface_rows(z["interface"], kw["ipw"])
    names = [str(s) for s in z["names"] if str(s).startswith("interface_")]
    assert names
    dof = grid.n_nodes
    p = np.random.default_rng(8).standard_normal(dof)
    q_ref, diag_ref, n_ref = np.zeros(dof), np.zeros(dof), 0
    for name in names:
        m, b, w = H.ref_family(z, name)
        n_ref += m.shape[0]
        q_ref += m.T @ (w * w * (m @ p))
        diag_ref += np.asarray(m.multiply(m).T @ (w * w)).ravel()
    assert n == n_ref
    assert np.abs(e.apply_normal(p, parts=2) - q_ref).max() <= 1e-12 * np.abs(q_ref).max()
    assert np.abs(e.diagonal() - diag_ref).max() <= 1e-12 * np.abs(diag_ref).max()
    assert not e.rhs().any()
    e.close()


@pytest.mark.parametrize("tag", ["unit16", "aniso", "unit24_tight"])
def test_interpolator_api_matches_reference(tag):
    """B200FiniteDifferenceInterpolator driven like the reference's class (set_*_constraints, setup_interpolator,
    solve_system, evaluate_*) reproduces the reference's coefficients and evaluations on the golden cases."""
    from loopstructural_b200.interpolators import B200FiniteDifferenceInterpolator, Support

    z = np.load(os.path.join(GOLDEN, f"fdi_solve_{tag}.npz"))
    grid, weights, fams, rtol, atol = H.solve_case(z)
    it = B200FiniteDifferenceInterpolator(Support(grid.origin, grid.nsteps - 1, grid.step))
    it.apply_scaling_matrix = bool(z["scaling"])
    for key in ("value", "gradient", "normal", "tangent"):
        if f"data_{key}" in z:
            getattr(it, f"set_{key}_constraints")(z[f"data_{key}"])
    kw = H.weights_of(z, "kw_keys", "kw_vals")
    counts = it.setup_interpolator(**kw)
    assert counts["value"] > 0 and counts["regularisation"] == "matrix-free"
    skw = {"rtol": rtol, "atol": atol}
    assert it.solve_system("b200", solver_kwargs=skw) is True
    t = it.get_last_solve_timing()
    assert t["backend"] == "b200" and abs(t["iterations"] - int(z["iterations"])) <= 1
    rng_ = float(z["c"].max() - z["c"].min())
    assert np.linalg.norm(it.c - z["c"]) / np.sqrt(it.c.size) / rng_ <= 1e-6
    q = z["query"]
    v = it.evaluate_value(q)
    np.testing.assert_array_equal(np.isnan(v), np.isnan(z["value_at"]))
    fin = ~np.isnan(v)
    assert np.abs(v[fin] - z["value_at"][fin]).max() <= 1e-5 * rng_
    # assigning c from outside (the reference lets callers do that) re-syncs the device copy
    it.set_c(z["c"])
    np.testing.assert_allclose(it.evaluate_value(q), z["value_at"], rtol=0, atol=1e-13, equal_nan=True)
    with pytest.raises(ValueError):
        it.solve_system("cg")
    it.close()


def test_interpolator_api_tol_semantics_and_errors():
    """tol -> atol unless both atol and rtol are given (LI/_solver_strategy.py:59-60); solve before setup fails."""
    from loopstructural_b200.interpolators import B200FiniteDifferenceInterpolator, Support

    grid, value, gradient = fo.synthetic_problem(12, 200, 20, seed=2)
    it = B200FiniteDifferenceInterpolator(Support(grid.origin, grid.nsteps - 1, grid.step))
    with pytest.raises(ValueError):
        it.solve_system("b200")
    with pytest.raises(ValueError):
        it.set_value_constraints(np.zeros((3, 4)))
    it.set_value_constraints(value)
    it.set_gradient_constraints(gradient)
    it.setup_interpolator()
    it.solve_system("b200", tol=1e-3)
    loose = it.get_last_solve_timing()
    assert loose["threshold"] >= 1e-3  # atol = tol
    it.solve_system("b200", tol=1e-3, solver_kwargs={"atol": 0.0, "rtol": 1e-9})
    tight = it.get_last_solve_timing()
    assert tight["iterations"] > loose["iterations"] and tight["threshold"] < loose["threshold"]
    assert it.update() is True
    it.close()


def test_full_size_solve_properties():
    """BASELINE configs[1] (256^3 nodes, 1e5 value + 1e4 gradient points) through the public class: the device
    CG stops by the reference's rule, the final residual recomputed from the returned c through the operator
    agrees with the recurrence, and the field honours the data."""
    from loopstructural_b200.interpolators import B200FiniteDifferenceInterpolator, Support

    n = 256
    grid, value, gradient = fo.synthetic_problem(n, 100_000, 10_000, seed=0)
    it = B200FiniteDifferenceInterpolator(Support(grid.origin, grid.nsteps - 1, grid.step))
    it.set_value_constraints(value)
    it.set_gradient_constraints(gradient)
    it.setup_interpolator()
    assert it.solve_system("b200") is True
    t = it.get_last_solve_timing()
    assert t["info"] == 0 and t["residual_norm"] < t["threshold"]
    e = it.engine
    S = e.scaling()
    r = e.rhs() - e.apply_normal(it.c)
    assert abs(np.linalg.norm(S * r) - t["residual_norm"]) <= 1e-6 * t["threshold"] + 1e-3 * t["residual_norm"]
    # values at the data points are close to the data (cpw = 1 against unit regularisation)
    v = it.evaluate_value(value[:, :3])
    assert np.nanmax(np.abs(v - value[:, 3])) < 0.05
    g = it.evaluate_gradient(gradient[:2000, :3])
    cosang = np.sum(g * gradient[:2000, 3:6], axis=1) / (np.linalg.norm(g, axis=1) * np.linalg.norm(gradient[:2000, 3:6], axis=1))
    assert np.median(cosang) > 0.95
    it.close()


def test_deterministic_mode_is_bitwise_repeatable_and_matches():
    """set_options(deterministic=True): the data rows are applied through the node-major transpose (stable sort,
    fixed-order sums, no atomics).  Two solves give identical bits; the answer matches the reference like the
    atomic path does; rhs / diagonal / operator agree with it to rounding."""
    z = np.load(os.path.join(GOLDEN, "fdi_solve_aniso.npz"))
    grid, weights, fams, rtol, atol = H.solve_case(z)
    runs = []
    for det in (True, True, False):
        e = H.engine_for(grid, weights)
        e.set_options(apply_scaling=bool(z["scaling"]), deterministic=det)
        H.add_point_data(e, z, weights)
        c, res = e.solve_cg(rtol=rtol, atol=atol)
        runs.append((c, res.iterations, e.rhs(), e.diagonal(), e.apply_normal(z["c"], parts=2)))
        e.close()
    (c1, it1, rhs1, d1, q1), (c2, it2, rhs2, d2, q2), (c3, it3, rhs3, d3, q3) = runs
    assert it1 == it2 and np.array_equal(c1, c2) and np.array_equal(rhs1, rhs2) and np.array_equal(d1, d2) and np.array_equal(q1, q2)
    rng_ = float(z["c"].max() - z["c"].min())
    assert np.linalg.norm(c1 - z["c"]) / np.sqrt(c1.size) / rng_ <= 1e-6 and abs(it1 - int(z["iterations"])) <= 1
    for a, b in ((rhs1, rhs3), (d1, d3), (q1, q3)):
        assert np.abs(a - b).max() <= 1e-13 * max(np.abs(b).max(), 1.0)


def test_deterministic_mode_with_interface_and_generic_rows():
    """Width-16 rows with duplicate columns (interface) through the transpose: diagonal merges duplicates per row."""
    z = np.load(os.path.join(GOLDEN, "fdi_rows_a.npz"))
    grid = fo.Grid.from_cells(z["ncells"], z["origin"], z["step"])
    kw = H.weights_of(z, "kw_keys", "kw_vals")
    out = []
    for det in (True, False):
        e = H.engine_for(grid, dict.fromkeys(("dxx", "dyy", "dzz", "dxy", "dyz", "dxz", "dx", "dy", "dz"), 0.0))
        e.set_options(add_ridge=False, apply_scaling=False, deterministic=det)
        from loopstructural_b200 import engine as en
        e.add_interface_rows(z["interface"], kw["ipw"])
        e.add_point_rows(en.ROWS_VALUE, z["value"], kw["cpw"])
        e.add_point_rows(en.ROWS_NORMAL, z["normal"], kw["npw"])
        p = np.random.default_rng(5).standard_normal(grid.n_nodes)
        out.append((e.apply_normal(p, parts=2), e.diagonal(), e.rhs()))
        e.close()
    for a, b in zip(*out):
        assert np.abs(a - b).max() <= 1e-12 * max(np.abs(b).max(), 1.0)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_directional_regularisation_rows_match_reference(tag):
    """lsb200_add_directional_rows (k_fdi_directional_rows) against the families of the unmodified reference's
    minimise_directional_gradient_change (tests/golden/fdi_directional_*.npz): row count after the zero-weight filter and
    the normal operator of the six families together."""
    import golden_helpers as H
    from loopstructural_b200 import Engine

    z = np.load(os.path.join(H.GOLDEN, f"fdi_directional_{tag}.npz"))
    grid = fo.Grid.from_cells(z["ncells"], z["origin"], z["step"])
    e = Engine(grid.nsteps, grid.origin, grid.step)
    e.set_regularisation({k: 0.0 for k in ("dxx", "dyy", "dzz", "dxy", "dxz", "dyz", "dx", "dy", "dz")})
    e.set_options(add_ridge=False, apply_scaling=False)
    kept = e.add_directional_rows(z["vectors"], float(z["w"]))
    p = np.random.default_rng(3).standard_normal(grid.n_nodes)
    q_ref, rows = np.zeros(grid.n_nodes), 0
    for name in z["names"]:
        m, _b, w = H.ref_family(z, str(name))
        q_ref += m.T @ (w * w * (m @ p))
        rows += m.shape[0]
    assert kept == rows
    assert np.abs(e.apply_normal(p, parts=2) - q_ref).max() <= 1e-11 * np.abs(q_ref).max()
    assert not np.any(e.rhs())
    e.close()
