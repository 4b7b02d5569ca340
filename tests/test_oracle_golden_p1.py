"""The P1 / TetMesh oracle against the outputs of the unmodified reference (tests/golden/p1_*.npz)."""
import glob
import os

import numpy as np
import pytest
from scipy import sparse

import golden_helpers as H
from oracle import fdi_oracle as fo
from oracle import p1_oracle as po

GOLDEN = H.GOLDEN


def _normal_ops(fam, dof, p):
    m = fam.matrix(dof)
    w2 = fam.w * fam.w
    return m.T @ (w2 * (m @ p)), np.asarray(m.multiply(m).T @ w2).ravel(), m.T @ (w2 * fam.b)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "p1_operator_*.npz"))))
def test_face_jump_operator_matches_reference(path):
    z = np.load(path)
    grid = fo.Grid.from_cells(z["ncells"], np.zeros(3), z["step"])
    fam = po.face_rows(grid, float(z["cgw"]))
    assert fam.A.shape[0] == int(z["n_rows"])
    scale = np.abs(z["y"]).max()
    for x, y in zip(z["x"], z["y"]):
        q, diag, _ = _normal_ops(fam, grid.n_nodes, x)
        assert np.abs(q - y).max() <= 1e-12 * scale
    assert np.abs(diag - z["diag"]).max() <= 1e-12 * np.abs(z["diag"]).max()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_point_rows_match_reference(tag):
    z = np.load(os.path.join(GOLDEN, f"p1_rows_{tag}.npz"))
    grid = fo.Grid.from_cells(z["ncells"], z["origin"], z["step"])
    kw = H.weights_of(z, "kw_keys", "kw_vals")
    dof = grid.n_nodes
    p = np.random.default_rng(3).standard_normal(dof)
    fams = {"value": po.value_rows(grid, z["value"], kw["cpw"]), "norm": po.normal_rows(grid, z["normal"], kw["npw"]),
            "undefined gradient orthogonal constraint": po.orthogonal_rows(grid, z["tangent"][:, :3], z["tangent"][:, 3:6], kw["tpw"])}
    assert sorted(fams) == sorted(str(s) for s in z["names"])
    for name, fam in fams.items():
        m, b, w = H.ref_family(z, name)
        assert fam.A.shape[0] == m.shape[0], name
        q, diag, rhs = _normal_ops(fam, dof, p)
        q_ref = m.T @ (w * w * (m @ p))
        assert np.abs(q - q_ref).max() <= 1e-11 * max(np.abs(q_ref).max(), 1e-300), name
        assert np.abs(rhs - m.T @ (w * w * b)).max() <= 1e-11 * max(np.abs(m.T @ (w * w * b)).max(), 1.0), name


@pytest.mark.parametrize("tag", ["unit12", "aniso", "unit16_tight"])
def test_solve_and_evaluation_match_reference(tag):
    z = np.load(os.path.join(GOLDEN, f"p1_solve_{tag}.npz"))
    grid = fo.Grid.from_cells(z["ncells"], z["origin"], z["step"])
    kw = H.weights_of(z, "kw_keys", "kw_vals")
    w = dict(po.P1_DEFAULT_WEIGHTS)
    if "regularisation" in kw:
        w["cgw"] = kw["regularisation"]
    w.update({k: v for k, v in kw.items() if k != "regularisation"})
    fams = [po.face_rows(grid, w["cgw"])]
    if "data_normal" in z:
        fams.append(po.normal_rows(grid, z["data_normal"], w["npw"]))
    fams.append(po.value_rows(grid, z["data_value"], w["cpw"]))
    if "data_tangent" in z:
        fams.append(po.orthogonal_rows(grid, z["data_tangent"][:, :3], z["data_tangent"][:, 3:6], w["tpw"]))
    rtol = 1e-5 if float(z["rtol"]) < 0 else float(z["rtol"])
    atol = 0.0 if float(z["tol"]) < 0 else float(z["tol"])
    c, res, _S = fo.solve_reference_path(fams, grid.n_nodes, rtol=rtol, atol=atol, apply_scaling=bool(z["scaling"]))
    assert abs(res.iterations - int(z["iterations"])) <= 1
    rng_ = float(z["c"].max() - z["c"].min())
    assert np.linalg.norm(c - z["c"]) / np.sqrt(c.size) / rng_ <= 1e-6
    q = z["query"]
    v = po.evaluate_value(grid, q, z["c"])
    np.testing.assert_array_equal(np.isnan(v), np.isnan(z["value_at"]))
    np.testing.assert_allclose(v, z["value_at"], rtol=0, atol=1e-12, equal_nan=True)
    fin = np.all(np.isfinite(q), axis=1)
    g = po.evaluate_gradient(grid, q[fin], z["c"])
    np.testing.assert_array_equal(np.isnan(g), np.isnan(z["gradient_at"]))
    np.testing.assert_allclose(g, z["gradient_at"], rtol=0, atol=1e-9, equal_nan=True)
