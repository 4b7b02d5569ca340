"""Golden fixtures of the fault-frame evaluation (SURVEY.md section 8 f, rank 1) from the UNMODIFIED reference:
a faulted model (LoopStructural.datasets.load_intrusion) is built through GeologicalModel, then
FaultSegment.apply_to_points / evaluate_displacement are run on a point cloud for several fault profiles.

    python tests/golden/make_golden_fault.py      (build container only)

Stored: the three solved fault-frame fields (support geometry, c, world -> local matrix), the profile as plain
numbers, displacement / steps / offset, the points and the reference's outputs.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402,F401  (sets up the reference import path and stubs)

from LoopStructural import GeologicalModel  # noqa: E402
from LoopStructural.datasets import load_intrusion  # noqa: E402
from LoopStructural.modelling.features.fault import _fault_function as ff  # noqa: E402

ZEROS, ONES, CUBIC, SMOOTH_PEAK = 0, 1, 2, 3


def profile1d(fn):
    """(kind, w[4], lim[2], min_v, max_v) of one reference profile object."""
    if fn is ff.smooth_peak:
        return [SMOOTH_PEAK, 0, 0, 0, 0, -1, 1, 0, 0]
    if isinstance(fn, ff.Ones):
        return [ONES, 0, 0, 0, 0, -1, 1, 0, 0]
    if isinstance(fn, ff.Zeros):
        return [ZEROS, 0, 0, 0, 0, -1, 1, 0, 0]
    if isinstance(fn, ff.CubicFunction):
        fn.solve()
        return [CUBIC, *[float(x) for x in fn.w], float(fn.lim[0]), float(fn.lim[1]), float(fn.min_v), float(fn.max_v)]
    raise TypeError(type(fn))


def profile_table(fd):
    """(3, 2, 9) table [axis][pos|neg] plus the per-axis composite flags of a FaultDisplacement."""
    tab = np.zeros((3, 2, 9))
    comp = np.zeros(3, dtype=np.int64)
    for a, fn in enumerate((fd.gx, fd.gy, fd.gz)):
        if isinstance(fn, ff.Composite):
            comp[a] = 1
            tab[a, 0] = profile1d(fn.positive)
            tab[a, 1] = profile1d(fn.negative)
        else:
            tab[a, 0] = tab[a, 1] = profile1d(fn)
    return tab, comp


def main(interpolatortype="FDI", prefix="fault_", only=None):
    data, bb = load_intrusion()
    model = GeologicalModel(bb[0, :], bb[1, :])
    model.data = data
    model.create_and_add_fault("fault", 500, nelements=2000, interpolatortype=interpolatortype, buffer=0.3)
    model.update()
    fault = model["fault"]
    fields = {}
    for i in range(3):
        feat = fault[i]
        assert len(feat.faults) == 0 and len(feat.regions) == 0
        it = feat.interpolator
        sup = it.support
        fields[f"f{i}_ncells"] = np.asarray(sup.nsteps_cells)
        fields[f"f{i}_origin"] = np.asarray(sup.origin, float)
        fields[f"f{i}_step"] = np.asarray(sup.step_vector, float)
        fields[f"f{i}_c"] = np.array(it.c)
        fields[f"f{i}_w2l"] = np.array(it.bounding_box.world_to_local_matrix, float)
        fields[f"f{i}_kind"] = np.array(1 if type(sup).__name__ == "TetMesh" else 0)
    rng = np.random.default_rng(7)
    ext = bb[1] - bb[0]
    pts = bb[0] + rng.uniform(-0.6, 1.6, (2000, 3)) * ext  # well beyond the buffered supports on every side
    near = bb[0] + rng.uniform(0, 1, (1000, 3)) * ext
    pts = np.vstack([pts, near])
    pts[5] = np.nan
    pts[17, 1] = np.nan

    # profiles spread over the ranges the three coordinates really take in this model (gx +-8, gy +-700, gz -1..3)
    cubic_y = ff.CubicFunction()
    for x, y in ((-600, 0), (600, 0), (-100, 1), (100, 1)):
        cubic_y.add_cstr(x, y)
    cubic_y.add_grad(0, 0)
    cubic_y.add_min(-1)
    cubic_y.add_max(0.8)
    cubic_y.set_lim(-650, 500)
    hw_wide, fw_wide = ff.CubicFunction(), ff.CubicFunction()
    for fn, s in ((hw_wide, 1.0), (fw_wide, -1.0)):
        fn.add_cstr(0, s)
        fn.add_grad(0, 0)
        fn.add_cstr(6 * s, 0)
        fn.add_grad(6 * s, 0)
    hw_wide.set_lim(0, 6)
    fw_wide.set_lim(-6, 0)
    cubic_z = ff.CubicFunction()
    for x, y in ((-1, 0.2), (3, 0.1), (1, 1)):
        cubic_z.add_cstr(x, y)
    cubic_z.add_grad(1, 0)
    cubic_z.set_lim(-0.5, 2.5)
    cases = {
        "base": (ff.BaseFault().fault_displacement, 500.0, 10, 0.0),
        "base3d": (ff.BaseFault3D().fault_displacement, -350.0, 4, 0.3),
        "wide": (ff.FaultDisplacement(hw=hw_wide, fw=fw_wide, gy=cubic_y, gz=cubic_z, scale=1.5), 200.0, 3, -0.1),
        "halfzero": (ff.FaultDisplacement(hw=hw_wide, fw=ff.Zeros(), gy=ff.Ones(), gz=ff.Ones(), scale=0.5), 300.0, 6, 0.0),
        "none": (None, 120.0, 5, 0.0),
    }
    for tag, (fd, disp, steps, offset) in cases.items():
        if only is not None and tag not in only:
            continue
        fault._faultfunction = fd
        fault.displacement = disp
        fault.steps = steps
        fault.fault_offset = offset
        moved = fault.apply_to_points(pts)
        moved_rev = fault.apply_to_points(pts, reverse=True)
        dsp = fault.evaluate_displacement(pts)
        out = dict(fields)
        if fd is not None:
            tab, comp = profile_table(fd)
            out.update(profile_table=tab, profile_composite=comp, profile_scale=float(fd.scale), has_function=1)
        else:
            out.update(profile_table=np.zeros((3, 2, 9)), profile_composite=np.zeros(3, dtype=np.int64), profile_scale=0.0,
                       has_function=0)
        vals = np.stack([fault[i].evaluate_value(pts) for i in range(3)], axis=1)
        grad1 = fault[1].evaluate_gradient(pts, ignore_regions=True)
        np.savez_compressed(os.path.join(HERE, f"{prefix}{tag}.npz"), points=pts, moved=moved, moved_reverse=moved_rev,
                            displacement_at=dsp, frame_values=vals, frame_gradient1=grad1, displacement=disp, steps=steps,
                            fault_offset=offset, **out)
        n_moved = int(np.sum(np.any(moved != pts, axis=1) & ~np.any(np.isnan(pts), axis=1)))
        print(tag, "points", pts.shape[0], "moved", n_moved, "max shift", float(np.nanmax(np.abs(moved - pts))),
              "nan out", int(np.isnan(moved).sum()), "reverse identical", bool(np.array_equal(moved, moved_rev, equal_nan=True)))


if __name__ == "__main__":
    main()
    # the same frame on a TetMesh (P1): evaluation is NaN outside for value AND gradient
    main("PLI", "faultp1_", only=("base", "wide", "none"))
