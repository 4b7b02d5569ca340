"""oracle/fault_oracle.py against outputs of the unmodified reference (tests/golden/fault_*.npz,
made by tests/golden/make_golden_fault.py): FaultSegment.apply_to_points / evaluate_displacement and the
fault-frame gathers behind them."""
import glob
import os

import numpy as np
import pytest

from golden_helpers import GOLDEN, fault_case
from oracle import fault_oracle as fa

CASES = sorted(os.path.basename(p)[6:-4] for p in glob.glob(os.path.join(GOLDEN, "fault_*.npz")))
P1_CASES = sorted(os.path.basename(p)[8:-4] for p in glob.glob(os.path.join(GOLDEN, "faultp1_*.npz")))


def test_fixtures_present():
    assert CASES == ["base", "base3d", "halfzero", "none", "wide"]


@pytest.mark.parametrize("tag", CASES)
def test_frame_gathers_match_the_reference(tag):
    z = np.load(os.path.join(GOLDEN, f"fault_{tag}.npz"))
    fields, *_ = fault_case(z)
    vals, grad = fa.frame_eval(fields, z["points"])
    np.testing.assert_array_equal(np.isnan(vals), np.isnan(z["frame_values"]))
    np.testing.assert_allclose(vals, z["frame_values"], rtol=0, atol=1e-12, equal_nan=True)
    # GeologicalFeature.evaluate_gradient leaves NaN only where a mask excludes the point: none here
    np.testing.assert_allclose(grad, z["frame_gradient1"], rtol=0, atol=1e-14)


@pytest.mark.parametrize("tag", CASES)
def test_apply_to_points_matches_the_reference(tag):
    z = np.load(os.path.join(GOLDEN, f"fault_{tag}.npz"))
    fields, profile, disp, steps, _ = fault_case(z)
    moved = fa.apply_to_points(fields, profile, disp, steps, z["points"])
    np.testing.assert_array_equal(np.isnan(moved), np.isnan(z["moved"]))
    np.testing.assert_allclose(moved, z["moved"], rtol=0, atol=1e-9, equal_nan=True)  # coordinates are O(1e4)
    # `reverse` does not change the reference's result (the reversed d is overwritten inside the loop)
    np.testing.assert_array_equal(z["moved"], z["moved_reverse"])


@pytest.mark.parametrize("tag", CASES)
def test_evaluate_displacement_matches_the_reference(tag):
    z = np.load(os.path.join(GOLDEN, f"fault_{tag}.npz"))
    fields, profile, disp, _, offset = fault_case(z)
    d = fa.evaluate_displacement(fields, profile, disp, offset, z["points"])
    np.testing.assert_allclose(d, z["displacement_at"], rtol=1e-13, atol=1e-13 * abs(disp))


def test_lattice_export_matches_the_reference():
    """oracle evaluate_lattice against BaseFeature.scalar_field of the unmodified reference (lattice_claudius.npz)."""
    from oracle import fdi_oracle as fo

    z = np.load(os.path.join(GOLDEN, "lattice_claudius.npz"))
    fld = fa.FrameField(fo.Grid.from_cells(z["ncells"], z["origin"], z["step"]), z["c"], z["w2l"])
    pts = fa.lattice_points(z["box_origin"], z["box_maximum"], z["box_nsteps"], "F")
    np.testing.assert_array_equal(pts[:50], z["points_f_head"])
    np.testing.assert_array_equal(pts[-50:], z["points_f_tail"])
    rng_ = float(np.nanmax(z["node_values_f"]) - np.nanmin(z["node_values_f"]))
    for order, key in (("F", "node_values_f"), ("C", "node_values_c")):
        v = fa.evaluate_lattice(fld, z["box_origin"], z["box_maximum"], z["box_nsteps"], order)
        np.testing.assert_allclose(v, z[key], rtol=0, atol=1e-13 * rng_, equal_nan=True)
    g = fa.evaluate_lattice(fld, z["box_origin"], z["box_maximum"], z["box_nsteps"], "F", gradient=True)
    np.testing.assert_allclose(g, z["node_gradient_f"], rtol=0, atol=1e-13)


@pytest.mark.parametrize("tag", P1_CASES)
def test_fault_frame_on_a_tet_mesh_matches_the_reference(tag):
    """The same frame solved with interpolatortype='PLI' (TetMesh): value AND gradient are NaN outside the mesh, so a
    point that leaves it becomes NaN (faultp1_none: 7 NaN coordinates in the reference's output)."""
    z = np.load(os.path.join(GOLDEN, f"faultp1_{tag}.npz"))
    assert P1_CASES == ["base", "none", "wide"] and int(z["f1_kind"]) == 1
    fields, profile, disp, steps, offset = fault_case(z)
    vals, grad = fa.frame_eval(fields, z["points"])
    np.testing.assert_array_equal(np.isnan(vals), np.isnan(z["frame_values"]))
    np.testing.assert_allclose(vals, z["frame_values"], rtol=0, atol=1e-11, equal_nan=True)
    np.testing.assert_array_equal(np.isnan(grad), np.isnan(z["frame_gradient1"]))
    np.testing.assert_allclose(grad, z["frame_gradient1"], rtol=0, atol=1e-13, equal_nan=True)
    moved = fa.apply_to_points(fields, profile, disp, steps, z["points"])
    np.testing.assert_array_equal(np.isnan(moved), np.isnan(z["moved"]))
    np.testing.assert_allclose(moved, z["moved"], rtol=0, atol=1e-9, equal_nan=True)
    d = fa.evaluate_displacement(fields, profile, disp, offset, z["points"])
    np.testing.assert_allclose(d, z["displacement_at"], rtol=1e-13, atol=1e-13 * abs(disp))
