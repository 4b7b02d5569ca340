"""The fold rows of the oracle (oracle/p1_oracle.py: element_vector_rows, face_vector_rows) against the families the
unmodified reference's DiscreteFoldInterpolator.add_fold_constraints produced (tests/golden/fold_rows_*.npz, written by
tests/golden/make_golden_fold.py)."""
import os

import numpy as np
import pytest

import golden_helpers as H
from oracle import fdi_oracle as fo
from oracle import p1_oracle as po

FAMILIES = (("fold orientation", "deformed_orientation", "fold_orientation", None),
            ("fold axis", "fold_axis", "fold_axis_w", None),
            ("fold normalisation", "dgz", "fold_normalisation", "fold_norm"))


def normal_ops(m, b, w, p):
    w2 = w * w
    return m.T @ (w2 * (m @ p)), np.asarray(m.multiply(m).T @ w2).ravel(), m.T @ (w2 * b)


def fold_case(tag):
    z = np.load(os.path.join(H.GOLDEN, f"fold_rows_{tag}.npz"))
    grid = fo.Grid.from_cells(z["ncells"], z["origin"], z["step"])
    return z, grid, H.weights_of(z, "w_keys", "w_vals")


@pytest.mark.parametrize("tag", ["a", "b"])
def test_element_rows_match_reference(tag):
    z, grid, w = fold_case(tag)
    rng = np.random.default_rng(5)
    p = rng.standard_normal(grid.n_nodes)
    seen = 0
    for name, vec, wkey, bkey in FAMILIES:
        if name not in z["names"]:
            assert wkey not in w
            continue
        seen += 1
        m, b, rw = H.ref_family(z, name)
        sel = z[f"{name}/elements"]
        target = w[bkey] * w[wkey] if bkey else 0.0  # B = fold_norm * fold_normalisation * vol (:170-172)
        fam = po.element_vector_rows(grid, z["barycentre"][sel], z[vec][sel], w[wkey], target, name)
        assert fam.A.shape[0] == m.shape[0] == len(sel)
        ref = normal_ops(m, b, rw, p)
        got = normal_ops(fam.matrix(grid.n_nodes), fam.b, fam.w, p)
        for r, g in zip(ref, got):
            assert np.abs(r - g).max() <= 1e-12 * max(np.abs(r).max(), 1e-300)
    assert seen >= 2


def test_direction_weighted_face_rows_match_reference():
    z, grid, _w = fold_case("a")
    assert int(z["n_shared_faces"]) == len(z["face_nodes_a"]) != int(z["n_elements"])
    p = np.random.default_rng(6).standard_normal(grid.n_nodes)
    for i, weight in enumerate(z["fold_regularisation"], start=1):
        m, b, rw = H.ref_family(z, f"fold regularisation {i}")
        fam = po.face_vector_rows(grid, z["face_nodes_a"], z["face_nodes_b"], z[f"face_vectors_{i}"], float(weight))
        assert fam.A.shape[0] == m.shape[0]
        ref = normal_ops(m, b, rw, p)
        got = normal_ops(fam.matrix(grid.n_nodes), fam.b, fam.w, p)
        for r, g in zip(ref[:2], got[:2]):
            assert np.abs(r - g).max() <= 1e-12 * np.abs(r).max()
        assert not np.any(ref[2]) and not np.any(got[2])


def test_shared_face_count_of_the_structured_tet_mesh():
    """10 abc - 2 (ab + bc + ca) interior faces for a x b x c cells: what the plug-in compares a vector's length with."""
    for tag in "ab":
        z, _g, _w = fold_case(tag)
        a, b, c = (int(v) for v in z["ncells"])
        assert int(z["n_shared_faces"]) == 10 * a * b * c - 2 * (a * b + b * c + c * a)
        assert int(z["n_elements"]) == 5 * a * b * c
