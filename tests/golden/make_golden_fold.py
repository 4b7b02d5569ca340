"""Golden fixtures for the fold rows of DiscreteFoldInterpolator, from the UNMODIFIED reference.

    python tests/golden/make_golden_fold.py      (build container only)

Writes tests/golden/fold_rows_*.npz: the families `add_fold_constraints` (LI/_discrete_fold_interpolator.py:66-204) puts
into the constraint dict for a fold whose geometry is three analytic vector fields -- 'fold orientation', 'fold axis',
'fold normalisation' (one row per selected tetrahedron) and the three 'fold regularisation n' families (face-jump rows) --
together with what determines them: the element barycentres, the three vector fields there and, per family, the
elements the reference's shuffles selected (the module-level generator is seeded here and its shuffles replayed).
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import dump_families  # noqa: E402  (sets up the reference import path)

from loop_common.math import rng  # noqa: E402  the generator add_fold_constraints shuffles with
from loop_common.supports import TetMesh  # noqa: E402
from loop_interpolation._discrete_fold_interpolator import DiscreteFoldInterpolator  # noqa: E402


class AnalyticFold:
    """Stands in for FoldEvent: get_deformed_orientation(points) -> (deformed orientation, fold axis, dgz)."""

    def __init__(self, seed):
        r = np.random.default_rng(seed)
        self.k = r.uniform(0.3, 1.1, (3, 3))
        self.p = r.uniform(0.0, 3.0, (3, 3))

    def field(self, i, x):
        k, p = self.k[i], self.p[i]
        return np.column_stack([np.sin(k[0] * x[:, 1] + p[0]) + 0.3, np.cos(k[1] * x[:, 2] + p[1]) - 0.2 * x[:, 0],
                                0.5 + np.sin(k[2] * x[:, 0] + p[2]) * np.cos(0.5 * x[:, 1])])

    def get_deformed_orientation(self, points):
        x = np.asarray(points, float)
        return self.field(0, x), self.field(1, x), self.field(2, x)


def case(tag, ncells, origin, step, weights, element_step, seed):
    mesh = TetMesh(origin=np.array(origin, float), nsteps_cells=np.array(ncells), step_vector=np.array(step, float))
    fold = AnalyticFold(seed)
    it = DiscreteFoldInterpolator(mesh, fold)
    it.reset()
    state = np.random.default_rng(1000 + seed).bit_generator.state
    rng.bit_generator.state = state
    it.add_fold_constraints(step=element_step, **weights)
    out = dump_families(it)
    # replay the shuffles (LI/_discrete_fold_interpolator.py:104-105,116,135,153): one before, one per family
    rng.bit_generator.state = state
    idx = np.arange(mesh.n_elements)
    rng.shuffle(idx)
    for name, key in (("fold orientation", "fold_orientation"), ("fold axis", "fold_axis_w"), ("fold normalisation", "fold_normalisation")):
        if weights.get(key) is not None:
            rng.shuffle(idx)
            out[f"{name}/elements"] = idx[::element_step].copy()
    # what the direction-weighted face rows are functions of (LI/_p1interpolator.py:115-166,
    # LI/_discrete_interpolator.py:432-449): the two tetrahedra of every shared face and the vectors at the face centres
    rel = np.asarray(mesh.shared_element_relationships)
    el = np.asarray(mesh.elements)
    centres = np.asarray(it.get_regularisation_sample_points(), float)
    fdo, ffa, fdgz = fold.get_deformed_orientation(centres)
    out.update(face_nodes_a=el[rel[:, 0]].astype(np.int64), face_nodes_b=el[rel[:, 1]].astype(np.int64), face_centres=centres,
               face_vectors_1=fdgz, face_vectors_2=fdo, face_vectors_3=ffa)
    bary = np.asarray(mesh.barycentre, float)
    do, fa, dgz = fold.get_deformed_orientation(bary)
    out.update(ncells=np.array(ncells), origin=np.array(origin, float), step=np.array(step, float), barycentre=bary,
               deformed_orientation=do, fold_axis=fa, dgz=dgz, element_step=element_step,
               n_shared_faces=int(mesh.shared_element_relationships.shape[0]), n_elements=int(mesh.n_elements),
               w_keys=np.array([k for k, v in weights.items() if np.ndim(v) == 0 and v is not None]),
               w_vals=np.array([float(v) for k, v in weights.items() if np.ndim(v) == 0 and v is not None]),
               fold_regularisation=np.array(weights.get("fold_regularisation") or [0.1, 0.01, 0.01], float))
    np.savez_compressed(os.path.join(HERE, f"fold_rows_{tag}.npz"), **out)
    print("fold rows", tag, {k: it.constraints[k]["matrix"].shape for k in it.constraints})


if __name__ == "__main__":
    case("a", (4, 5, 3), (0.0, 0.0, 0.0), (1.0, 1.0, 1.0),
         dict(fold_orientation=10.0, fold_axis_w=7.0, fold_normalisation=1.5, fold_norm=0.8, fold_regularisation=[0.1, 0.02, 0.03]), 2, 1)
    case("b", (5, 4, 6), (-2.0, 1.0, 0.5), (0.8, 1.25, 1.1),
         dict(fold_orientation=3.0, fold_axis_w=None, fold_normalisation=2.0, fold_norm=-1.0, fold_regularisation=None), 3, 2)
