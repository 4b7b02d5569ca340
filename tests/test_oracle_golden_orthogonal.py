"""oracle/fault_oracle.field_orthogonal_rows against the '<feature>_orthogonal' families of the unmodified reference
(tests/golden/orth_*.npz, made by tests/golden/make_golden_orthogonal.py)."""
import os

import numpy as np
import pytest

from golden_helpers import GOLDEN, ref_family
from oracle import fault_oracle as fa
from oracle import fdi_oracle as fo


def orth_case(z, key):
    """target grid, the other field (oracle FrameField), step, weight, B and the reference family of one key 't<i>_o<j>'."""
    ti, oj = int(key[1]), int(key[4])
    grid = fo.Grid.from_cells(z[f"f{ti}_ncells"], z[f"f{ti}_origin"], z[f"f{ti}_step"])
    other = fa.FrameField(fo.Grid.from_cells(z[f"f{oj}_ncells"], z[f"f{oj}_origin"], z[f"f{oj}_step"]), z[f"f{oj}_c"], z[f"f{oj}_w2l"])
    return grid, other, int(z[f"{key}/step"]), float(z[f"{key}/weight"]), float(z[f"{key}/B"]), ref_family(z, key)


@pytest.mark.parametrize("tag", ["world", "zero_origin"])
def test_field_orthogonal_rows_match_the_reference(tag):
    z = np.load(os.path.join(GOLDEN, f"orth_{tag}.npz"))
    assert len(z["families"]) == 3
    for key in z["families"]:
        grid, other, step, w, B, (m_ref, b_ref, w_ref) = orth_case(z, str(key))
        fam = fa.field_orthogonal_rows(grid, other, step, w, B)
        assert fam.A.shape[0] == m_ref.shape[0]
        diff = abs(fam.matrix(grid.n_nodes) - m_ref)
        assert diff.max() <= 1e-13
        np.testing.assert_allclose(fam.b, b_ref, rtol=0, atol=1e-14)
        np.testing.assert_array_equal(fam.w, w_ref)
    if tag == "world":
        # the reference evaluates the other feature at LOCAL barycentres as if they were world points: with a
        # translated local frame more than half of the rows come out empty -- restated as is
        assert int(np.sum(np.abs(fam.A).sum(axis=1) == 0)) == 106


@pytest.mark.parametrize("tag", ["p1_world", "p1_zero_origin"])
def test_p1_field_orthogonal_rows_match_the_reference(tag):
    """The P1 / TetMesh twin (fault frame built with interpolatortype='PLI'): with the dataset's own origin every
    barycentre lands outside the other feature's mesh (local coordinates read as world coordinates), the gradient is NaN and
    the reference drops all three families; with a zero origin it keeps 5400 rows each."""
    z = np.load(os.path.join(GOLDEN, f"orth_{tag}.npz"))
    assert str(z["kind"]) == "PLI" and len(z["families"]) == 3
    for key in z["families"]:
        key = str(key)
        ti, oj = int(key[1]), int(key[4])
        grid = fo.Grid.from_cells(z[f"f{ti}_ncells"], z[f"f{ti}_origin"], z[f"f{ti}_step"])
        other = fa.FrameField(fo.Grid.from_cells(z[f"f{oj}_ncells"], z[f"f{oj}_origin"], z[f"f{oj}_step"]), z[f"f{oj}_c"], z[f"f{oj}_w2l"],
                              "tetmesh")
        fam = fa.p1_field_orthogonal_rows(grid, other, int(z[f"{key}/step"]), float(z[f"{key}/weight"]), float(z[f"{key}/B"]))
        if f"{key}/dropped" in z.files:
            assert fam is None
            continue
        m_ref, b_ref, w_ref = ref_family(z, key)
        assert fam is not None and fam.A.shape[0] == m_ref.shape[0]
        assert abs(fam.matrix(grid.n_nodes) - m_ref).max() <= 1e-12
        np.testing.assert_allclose(fam.b, b_ref, rtol=0, atol=1e-13)
        np.testing.assert_allclose(fam.w, w_ref, rtol=1e-12, atol=0)
    assert (tag == "p1_world") == all(f"{str(k)}/dropped" in z.files for k in z["families"])
