"""Host-side logic of the package that needs no GPU: weights, interface pairing, the P1 stencil table and its class
lookup (against the oracle's explicit face rows), argument validation."""
import numpy as np
import pytest

from loopstructural_b200 import engine as en
from loopstructural_b200 import p1_tables as pt
from loopstructural_b200.interpolators import Support
from oracle import fdi_oracle as fo
from oracle import p1_oracle as po


def test_effective_stencil_weights_divide_mixed_terms_by_four():
    w = en.effective_stencil_weights(dict(dxy=2.0, dx=0.5, dzz=3.0))
    ref = fo.effective_stencil_weights(dict(fo.DEFAULT_WEIGHTS, dxy=2.0, dx=0.5, dzz=3.0))
    assert w.tolist() == [ref[k] for k in ("dxx", "dyy", "dzz", "dxy", "dxz", "dyz", "dx", "dy", "dz")]
    assert w[3] == 0.5 and w[8] == 1.0


@pytest.mark.parametrize("n,mode,threshold,fallback", [(0, "auto", 200, "star"), (1, "all", 200, "star"), (5, "auto", 200, "star"),
                                                       (7, "auto", 3, "chain"), (6, "star", 200, "x"), (4, "chain", 200, "star")])
def test_interface_pair_indices_match_the_oracle(n, mode, threshold, fallback):
    a = en.interface_pair_indices(n, mode, threshold, fallback)
    b = fo.interface_pairs(n, mode, threshold, fallback if fallback in ("star", "chain") else "star")
    assert a.shape == b.shape and np.array_equal(a, b)
    with pytest.raises(ValueError):
        en.interface_pair_indices(5, "nonsense")


def test_support_validation():
    s = Support([0, 1, 2], [3, 4, 5], [0.5, 1.0, 2.0])
    assert s.nsteps.tolist() == [4, 5, 6] and s.n_nodes == 120 and s.maximum.tolist() == [1.5, 5.0, 12.0]
    with pytest.raises(ValueError):
        Support([0, 0, 0], [0, 4, 5], [1, 1, 1])
    with pytest.raises(ValueError):
        Support([0, 0, 0], [3, 4, 5], [1, 0, 1])


def _rep(i, nc, m):
    if nc == m:
        return i
    if i < 2:
        return i
    if i > nc - 2:
        return m - (nc - i)
    return 2 + ((i - 2) & 1)


@pytest.mark.parametrize("ncells,step", [((9, 8, 7), (0.5, 1.25, 2.0)), ((3, 3, 3), (1, 1, 1)), ((12, 7, 10), (1, 1, 1)),
                                         ((1, 2, 1), (1, 1, 1)), ((6, 5, 7), (1, 1, 1))])
def test_p1_stencil_table_reproduces_the_explicit_gram(ncells, step):
    """The table of a <= 6-cell mesh + the class lookup of csrc/p1.cuh (p1_rep) == R^T R of the full mesh, row by
    row, and a non-zero coefficient never points outside the mesh."""
    m, T = pt.regularisation_table(ncells, step)
    assert m.tolist() == pt.small_mesh_cells(ncells).tolist() and T.shape[1] == 125
    assert int((T != 0).sum(axis=1).max()) <= 36  # P1_TAPS of the compacted device table
    grid = fo.Grid.from_cells(ncells, np.zeros(3), step)
    M = po.face_rows(grid, 1.0).matrix(grid.n_nodes)
    N = (M.T @ M).toarray()
    n = grid.nsteps
    for k in range(n[2]):
        for j in range(n[1]):
            for i in range(n[0]):
                r = (_rep(k, ncells[2], m[2]) * (m[1] + 1) + _rep(j, ncells[1], m[1])) * (m[0] + 1) + _rep(i, ncells[0], m[0])
                row = np.zeros(grid.n_nodes)
                t = 0
                for dz in range(-2, 3):
                    for dy in range(-2, 3):
                        for dx in range(-2, 3):
                            c = T[r, t]
                            t += 1
                            if c != 0.0:
                                assert 0 <= i + dx < n[0] and 0 <= j + dy < n[1] and 0 <= k + dz < n[2]
                                row[(i + dx) + n[0] * ((j + dy) + n[1] * (k + dz))] = c
                assert np.abs(row - N[i + n[0] * (j + n[1] * k)]).max() <= 1e-13


def test_mirror_drops_a_whole_family_that_holds_a_nan_like_the_reference(monkeypatch):
    """add_constraints_to_least_squares returns without adding anything when A, B or the indices of a family contain a
    NaN (LI/_discrete_interpolator.py:355-359) -- a NaN value constraint inside the support silences ALL value
    constraints; one outside the support never reaches the check.  The mirror classes follow that rule."""
    import warnings

    from loopstructural_b200 import interpolators as ip

    calls = []

    class FakeEngine:
        def __init__(self, *a, **k):
            self.kind = k.get("kind", "grid")

        def clear_rows(self):
            calls.clear()

        def set_regularisation(self, w):
            pass

        def set_p1_regularisation(self, w):
            pass

        def add_point_rows(self, kind, data, weight):
            calls.append((kind, len(data)))
            return len(data)

        def add_interface_rows(self, data, weight):
            return 0

    monkeypatch.setattr(ip, "Engine", FakeEngine)
    sup = ip.Support(np.zeros(3), np.array([4, 4, 4]), np.ones(3))
    good = np.array([[0.5, 0.5, 0.5, 1.0, 1.0], [1.5, 2.5, 0.5, 2.0, 1.0]])
    nan_inside = np.vstack([good, [[2.5, 2.5, 2.5, np.nan, 1.0]]])
    nan_outside = np.vstack([good, [[9.0, 2.5, 2.5, np.nan, 1.0]]])
    for cls in (ip.B200FiniteDifferenceInterpolator, ip.B200P1Interpolator):
        it = cls(sup)
        it.set_value_constraints(nan_outside)
        assert it.setup_interpolator()["value"] == 3 and calls == [(ip.en.ROWS_VALUE, 3)]  # the engine drops the outside row itself
        it.set_value_constraints(nan_inside)
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            assert it.setup_interpolator()["value"] == 0 and calls == []
        assert any("contain nan" in str(x.message) for x in w)
        nrm = np.array([[0.5, 0.5, 0.5, 0.0, np.nan, 1.0, 1.0]])
        it.set_value_constraints(good)
        it.set_normal_constraints(nrm)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            rc = it.setup_interpolator()
        assert rc["normal"] == 0 and rc["value"] == 2
