"""Pin the numpy oracle against outputs of the reference itself (tests/golden/*.npz,
made by tests/golden/make_golden.py from the unmodified reference)."""
import glob
import os

import numpy as np
import pytest
import scipy.sparse.linalg as spla
from scipy import sparse

from oracle import fdi_oracle as fo
import golden_helpers as H

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _weights(z, kk="weights_keys", kv="weights_vals"):
    return {str(k): float(v) for k, v in zip(z[kk], z[kv])}


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "fdi_operator_*.npz"))))
def test_regularisation_operator_matches_reference(path):
    z = np.load(path)
    nodes = z["nodes"]
    grid = fo.Grid(nodes, np.zeros(3), np.ones(3))
    w = _weights(z)
    fams = fo.regularisation_families(grid, w)
    A, b = fo.assemble(fams, grid.n_nodes)
    assert A.shape[0] == int(z["n_rows"])
    assert A.nnz == int(z["nnz"])
    N = (A.T @ A).tocsr()
    scale = np.abs(z["y"]).max()
    for x, y in zip(z["x"], z["y"]):
        assert np.abs(N @ x - y).max() <= 1e-12 * scale
        # matrix-free restatement, every dof including the boundary shell
        assert np.abs(fo.regularisation_normal_matvec(grid, w, x) - y).max() <= 1e-12 * scale
    assert np.abs(N.diagonal() - z["diag"]).max() <= 1e-13
    assert np.abs(fo.regularisation_normal_diagonal(grid, w) - z["diag"]).max() <= 1e-13


def _ref_family(z, name):
    shape = tuple(z[f"{name}/shape"])
    m = sparse.coo_matrix((z[f"{name}/vals"], (z[f"{name}/rows"], z[f"{name}/cols"])), shape=shape).tocsr()
    return m, z[f"{name}/b"], z[f"{name}/w"]


@pytest.mark.parametrize("tag", ["a", "b"])
def test_data_rows_match_reference(tag):
    z = np.load(os.path.join(GOLDEN, f"fdi_rows_{tag}.npz"))
    grid = fo.Grid.from_cells(z["ncells"], z["origin"], z["step"])
    kw = _weights(z, "kw_keys", "kw_vals")
    fams = []
    fams += fo.normal_rows(grid, z["normal"], kw["npw"])
    fams += fo.gradient_rows(grid, z["gradient"], kw["gpw"])
    fams.append(fo.value_rows(grid, z["value"], kw["cpw"]))
    fams.append(fo.orthogonal_rows(grid, z["tangent"][:, :3], z["tangent"][:, 3:6], kw["tpw"]))
    fams += fo.interface_rows(grid, z["interface"], kw["ipw"])
    # the reference also lists the six interior stencil families here (weight 0 in this fixture)
    ref_names = [str(n) for n in z["names"] if str(n) not in fo.STENCILS]
    mine = {f.name: f for f in fams}
    assert sorted(mine) == sorted(ref_names)
    for name in ref_names:
        m, b, w = _ref_family(z, name)
        f = mine[name]
        mm = f.matrix(grid.n_nodes)
        assert mm.shape == m.shape, name
        assert abs(mm - m).max() <= 1e-13, name
        np.testing.assert_allclose(f.b, b, rtol=0, atol=1e-13, err_msg=name)
        np.testing.assert_allclose(f.w, w, rtol=0, atol=1e-15, err_msg=name)


def _solve_case(z):
    grid = fo.Grid.from_cells(z["ncells"], z["origin"], z["step"])
    kw = _weights(z, "kw_keys", "kw_vals")
    weights = dict(fo.DEFAULT_WEIGHTS)
    if "regularisation" in kw:
        for k in ("dxy", "dyz", "dxz", "dxx", "dyy", "dzz"):
            weights[k] = kw["regularisation"]
    weights.update({k: v for k, v in kw.items() if k != "regularisation"})
    fams = []
    if "data_normal" in z:
        fams += fo.normal_rows(grid, z["data_normal"], weights["npw"])
    if "data_gradient" in z:
        fams += fo.gradient_rows(grid, z["data_gradient"], weights["gpw"])
    fams.append(fo.value_rows(grid, z["data_value"], weights["cpw"]))
    if "data_tangent" in z:
        fams.append(fo.orthogonal_rows(grid, z["data_tangent"][:, :3], z["data_tangent"][:, 3:6], weights["tpw"]))
    rtol = 1e-5 if float(z["rtol"]) < 0 else float(z["rtol"])
    atol = 0.0 if float(z["tol"]) < 0 else float(z["tol"])
    return grid, weights, fams, rtol, atol


@pytest.mark.parametrize("tag", ["unit16", "unit16_noscale", "aniso", "unit24_tight"])
def test_solve_matches_reference(tag):
    z = np.load(os.path.join(GOLDEN, f"fdi_solve_{tag}.npz"))
    grid, weights, fams, rtol, atol = _solve_case(z)
    scaling = bool(z["scaling"])
    reg = fo.regularisation_families(grid, weights)
    timing = {}
    c, res, S = fo.solve_reference_path(fams + reg, grid.n_nodes, rtol=rtol, atol=atol, apply_scaling=scaling,
                                        timing=timing)
    assert timing["matrix_rows"] == int(z["matrix_rows"]) + grid.n_nodes  # reference counts rows before the ridge
    assert res.iterations == int(z["iterations"])
    rng_ = float(z["c"].max() - z["c"].min())
    rel = np.linalg.norm(c - z["c"]) / np.sqrt(c.size) / rng_
    assert rel <= 1e-7, rel  # measured 1e-12..6e-9: rounding-order noise amplified by CG
    # matrix-free restatement follows the same Krylov trajectory
    mf = fo.MatrixFreeSystem(grid, fams, weights, apply_scaling=scaling)
    c2, res2 = mf.solve(rtol=rtol, atol=atol)
    assert abs(res2.iterations - int(z["iterations"])) <= 1
    rel2 = np.linalg.norm(c2 - z["c"]) / np.sqrt(c.size) / rng_
    assert rel2 <= 1e-6, rel2


@pytest.mark.parametrize("tag", ["unit16", "aniso"])
def test_evaluate_matches_reference(tag):
    z = np.load(os.path.join(GOLDEN, f"fdi_solve_{tag}.npz"))
    grid = fo.Grid.from_cells(z["ncells"], z["origin"], z["step"])
    q = z["query"]
    v = fo.evaluate_value(grid, q, z["c"])
    np.testing.assert_array_equal(np.isnan(v), np.isnan(z["value_at"]))
    np.testing.assert_allclose(v, z["value_at"], rtol=0, atol=1e-13, equal_nan=True)
    assert np.isnan(v).sum() > 10 and v[3] == 0.0  # outside rows NaN, NaN-coordinate row 0.0
    fin = np.all(np.isfinite(q), axis=1)
    g = fo.evaluate_gradient(grid, q[fin], z["c"])
    np.testing.assert_allclose(g, z["gradient_at"], rtol=0, atol=1e-11)


def test_cg_restatement_matches_scipy():
    rng = np.random.default_rng(5)
    n = 400
    M = sparse.random(n, n, density=0.02, random_state=3, format="csr")
    N = (M.T @ M + sparse.eye(n) * 0.1).tocsr()
    b = rng.standard_normal(n)
    its = []
    x_ref, info = spla.cg(N, b, rtol=1e-8, atol=0.0, callback=lambda xk: its.append(1))
    res = fo.cg(lambda v: N @ v, b, rtol=1e-8)
    assert info == 0 and res.info == 0
    assert res.iterations == len(its)
    np.testing.assert_allclose(res.x, x_ref, rtol=0, atol=1e-12 * np.abs(x_ref).max())
    # iteration cap: info = maxiter, last iterate returned
    x_ref, info = spla.cg(N, b, rtol=1e-14, atol=0.0, maxiter=7)
    res = fo.cg(lambda v: N @ v, b, rtol=1e-14, maxiter=7)
    assert info == 7 and res.info == 7
    np.testing.assert_allclose(res.x, x_ref, rtol=0, atol=1e-12 * np.abs(x_ref).max())
    # zero right-hand side returns immediately
    assert fo.cg(lambda v: N @ v, np.zeros(n)).iterations == 0


def test_claudius_quickstart_matches_reference():
    """BASELINE configs[0] (README quickstart: Claudius 'strati' foliation through GeologicalModel, FDI, scipy cg) at
    nelements=2e4: the oracle fed with what reached the reference's interpolator reproduces its iteration count and
    coefficients, and the evaluation of the reference's field."""
    z = np.load(os.path.join(GOLDEN, "fdi_claudius_2e4.npz"))
    grid = fo.Grid.from_cells(z["ncells"], z["origin"], z["step"])
    kw = _weights(z, "kw_keys", "kw_vals")
    w = dict(fo.DEFAULT_WEIGHTS)
    w.update({k: v for k, v in kw.items() if k in w})
    fams = fo.normal_rows(grid, z["data_normal"], w["npw"]) + [fo.value_rows(grid, z["data_value"], w["cpw"])]
    fams += fo.regularisation_families(grid, w)
    c, res, _ = fo.solve_reference_path(fams, grid.n_nodes, rtol=1e-5, atol=float(z["tol"]))
    assert res.iterations == int(z["iterations"])
    rng_ = float(z["c"].max() - z["c"].min())
    assert np.linalg.norm(c - z["c"]) / np.sqrt(c.size) / rng_ <= 1e-9
    v = fo.evaluate_value(grid, z["query"], z["c"])
    np.testing.assert_array_equal(np.isnan(v), np.isnan(z["value_at"]))
    np.testing.assert_allclose(v, z["value_at"], rtol=0, atol=1e-9, equal_nan=True)
    np.testing.assert_allclose(fo.evaluate_gradient(grid, z["query"], z["c"]), z["gradient_at"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_directional_regularisation_families_match_reference(tag):
    """fdi_oracle.directional_families against FDI.minimise_directional_gradient_change of the unmodified reference
    (tests/golden/fdi_directional_*.npz): the same rows survive the zero-weight filter, the same normal equations."""
    z = np.load(os.path.join(GOLDEN, f"fdi_directional_{tag}.npz"))
    grid = fo.Grid.from_cells(z["ncells"], z["origin"], z["step"])
    fams = {f.name: f for f in fo.directional_families(grid, z["vectors"], float(z["w"]), name="fold regularisation 1")}
    assert set(fams) == set(str(n) for n in z["names"]) and len(fams) == 6
    p = np.random.default_rng(8).standard_normal(grid.n_nodes)
    for name, fam in fams.items():
        m, b, w = H.ref_family(z, name)
        assert fam.A.shape[0] == m.shape[0], name
        M = fam.matrix(grid.n_nodes)
        ref = m.T @ (w * w * (m @ p))
        got = M.T @ (fam.w * fam.w * (M @ p))
        assert np.abs(ref - got).max() <= 1e-12 * np.abs(ref).max(), name
        assert not np.any(b)
