"""Golden fixture of the regular-grid field export (SURVEY.md section 8 f, rank 4) from the UNMODIFIED reference:
BaseFeature.scalar_field (node values on BoundingBox.regular_grid(local=False, order='F')) of the Claudius 'strati'
foliation built through GeologicalModel, plus evaluate_gradient on the same lattice and the 'C'-ordered lattice.

    python tests/golden/make_golden_lattice.py      (build container only)
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402,F401  (sets up the reference import path and stubs)

from LoopStructural import GeologicalModel  # noqa: E402
from LoopStructural.datasets import load_claudius  # noqa: E402


def main():
    data, bb = load_claudius()
    model = GeologicalModel(bb[0, :], bb[1, :])
    model.data = data
    model.create_and_add_foliation("strati", nelements=3e3)
    model.update()
    feat = model["strati"]
    it = feat.interpolator
    sup = it.support
    box = model.bounding_box
    box.nsteps = np.array([23, 17, 11])
    field = feat.scalar_field()
    pts_f = box.regular_grid(local=False, order="F")
    pts_c = box.regular_grid(local=False, order="C")
    assert np.array_equal(np.asarray(field.properties["strati"]), feat.evaluate_value(pts_f), equal_nan=True)
    np.savez_compressed(
        os.path.join(HERE, "lattice_claudius.npz"), ncells=np.asarray(sup.nsteps_cells), origin=np.asarray(sup.origin, float),
        step=np.asarray(sup.step_vector, float), c=np.array(it.c), w2l=np.array(it.bounding_box.world_to_local_matrix, float),
        box_origin=np.asarray(box.origin, float), box_maximum=np.asarray(box.maximum, float), box_nsteps=np.asarray(box.nsteps),
        node_values_f=np.asarray(field.properties["strati"], float), node_values_c=feat.evaluate_value(pts_c),
        node_gradient_f=feat.evaluate_gradient(pts_f), points_f_head=pts_f[:50], points_f_tail=pts_f[-50:])
    v = np.asarray(field.properties["strati"], float)
    print("lattice", box.nsteps, "points", v.size, "nan", int(np.isnan(v).sum()), "range", float(np.nanmin(v)), float(np.nanmax(v)))


if __name__ == "__main__":
    main()
