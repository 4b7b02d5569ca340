"""The LoopStructural plug-in against the UNMODIFIED reference (CPU; needs /root/reference, which exists in the
build container only).  The device solve itself is covered by the -m gpu tests through the same Engine calls."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refenv  # noqa: E402

pytestmark = pytest.mark.skipif(not refenv.available(), reason="reference neither installed (baseline/_ref) nor mounted")


@pytest.fixture(scope="module")
def li():
    refenv.activate()
    import loop_interpolation

    return loop_interpolation


def _problem(li, n=9, seed=0):
    from loop_common.supports import StructuredGrid
    from oracle import fdi_oracle as fo

    grid, value, gradient = fo.synthetic_problem(n, 60, 12, seed=seed)
    sg = StructuredGrid(origin=grid.origin, nsteps_cells=grid.nsteps - 1, step_vector=grid.step)
    return grid, sg, value, gradient


def test_install_registers_subclass_and_uninstall_restores(li):
    from loopstructural_b200 import plugin

    stock = li.interpolator_map[li.InterpolatorType.FINITE_DIFFERENCE]
    cls = plugin.install()["FDI"]
    try:
        assert li.interpolator_map[li.InterpolatorType.FINITE_DIFFERENCE] is cls
        assert issubclass(cls, li.FiniteDifferenceInterpolator)
        assert plugin.install()["FDI"] is cls  # idempotent
    finally:
        plugin.uninstall()
    assert li.interpolator_map[li.InterpolatorType.FINITE_DIFFERENCE] is stock


def test_b200_setup_diverts_only_the_stencil_families(li):
    """With solver='b200' the data families are the reference's own; the twelve stencil families are reduced to
    their effective weights (mixed terms / 4), exactly what lsb200_set_regularisation takes."""
    from loopstructural_b200 import plugin
    from oracle import fdi_oracle as fo

    cls = plugin.install()["FDI"]
    plugin.uninstall()
    grid, sg, value, gradient = _problem(li)
    it = cls(sg)
    it.set_value_constraints(value)
    it.set_gradient_constraints(gradient)
    it.setup_interpolator(solver="b200", dxy=2.0, dx=0.5)
    stock = li.FiniteDifferenceInterpolator(sg)
    stock.set_value_constraints(value)
    stock.set_gradient_constraints(gradient)
    stock.setup_interpolator(dxy=2.0, dx=0.5)
    data_names = [k for k in stock.constraints if k.split("_")[0] not in ("dxx", "dyy", "dzz", "dxy", "dxz", "dyz", "dx", "dy", "dz")]
    assert sorted(it.constraints) == sorted(data_names)
    ew = fo.effective_stencil_weights(dict(fo.DEFAULT_WEIGHTS, dxy=2.0, dx=0.5))
    assert it._b200_stencil == {k: pytest.approx(ew[k]) for k in ("dxx", "dyy", "dzz", "dxy", "dxz", "dyz", "dx", "dy", "dz")}
    # ELL conversion reproduces the reference's matrices
    rows = plugin.constraints_to_rows(it.constraints)
    for name, vals, cols, b, w in rows:
        m = stock.constraints[name]["matrix"].toarray()
        dense = np.zeros_like(m)
        np.add.at(dense, (np.repeat(np.arange(vals.shape[0]), vals.shape[1]), cols.ravel()), vals.ravel())
        assert np.array_equal(dense, m), name
        assert np.array_equal(b, stock.constraints[name]["b"]) and np.array_equal(w, stock.constraints[name]["w"])


def test_other_solvers_still_match_the_stock_class(li):
    """A diverted setup followed by solver='cg' re-assembles the stencil rows and gives the stock answer."""
    from loopstructural_b200 import plugin

    cls = plugin.install()["FDI"]
    plugin.uninstall()
    grid, sg, value, gradient = _problem(li, n=8, seed=3)
    it = cls(sg)
    it.set_value_constraints(value)
    it.set_gradient_constraints(gradient)
    it.setup_interpolator(solver="b200")
    assert it.solve_system(solver="cg", tol=1e-10)
    stock = li.FiniteDifferenceInterpolator(sg)
    stock.set_value_constraints(value)
    stock.set_gradient_constraints(gradient)
    stock.setup_interpolator()
    assert stock.solve_system(solver="cg", tol=1e-10)
    assert np.allclose(it.c, stock.c, rtol=0, atol=1e-8)  # both stop at ||r|| < 1e-10


def test_b200_solve_fails_loudly_without_a_gpu(li):
    from loopstructural_b200 import _native as nat, plugin

    if nat.load().lsb200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    cls = plugin.install()["FDI"]
    plugin.uninstall()
    grid, sg, value, gradient = _problem(li, n=6)
    it = cls(sg)
    it.set_value_constraints(value)
    it.setup_interpolator(solver="b200")
    with pytest.raises(nat.Lsb200Error):
        it.solve_system(solver="b200")


def test_p1_subclass_diverts_only_the_face_rows(li):
    """P1: with solver='b200' the edge-jump family is reduced to its weight; every data family is the reference's;
    any other solver re-assembles it and gives the stock answer."""
    from loop_common.supports import TetMesh
    from loopstructural_b200 import plugin
    from oracle import fdi_oracle as fo

    classes = plugin.install()
    assert li.interpolator_map[li.InterpolatorType.PIECEWISE_LINEAR] is classes["P1"]
    plugin.uninstall()
    assert li.interpolator_map[li.InterpolatorType.PIECEWISE_LINEAR] is li.P1Interpolator
    grid, value, gradient = fo.synthetic_problem(7, 40, 10, seed=4)
    value[:, :3] = np.clip(value[:, :3], 0.01, 0.99)
    gradient[:, :3] = np.clip(gradient[:, :3], 0.01, 0.99)
    mesh = TetMesh(origin=grid.origin, nsteps_cells=grid.nsteps - 1, step_vector=grid.step)
    it = classes["P1"](mesh)
    it.set_value_constraints(value)
    it.set_normal_constraints(gradient)
    it.setup_interpolator(solver="b200", cgw=0.25)
    assert it._b200_reg == {"cgw": 0.25} and sorted(it.constraints) == ["norm", "value"]
    stock = li.P1Interpolator(mesh)
    stock.set_value_constraints(value)
    stock.set_normal_constraints(gradient)
    stock.setup_interpolator(cgw=0.25)
    for name, vals, cols, b, w in plugin.constraints_to_rows(it.constraints):
        m = stock.constraints[name]["matrix"].toarray()
        dense = np.zeros_like(m)
        np.add.at(dense, (np.repeat(np.arange(vals.shape[0]), vals.shape[1]), cols.ravel()), vals.ravel())
        assert np.array_equal(dense, m), name
    assert it.solve_system(solver="cg", tol=1e-10) and stock.solve_system(solver="cg", tol=1e-10)
    assert "edge jump" in it.constraints
    assert np.allclose(it.c, stock.c, rtol=0, atol=1e-8)
