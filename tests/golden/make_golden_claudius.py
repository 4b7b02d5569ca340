"""Golden fixture from the reference's README quickstart (BASELINE configs[0]: Claudius 'strati' foliation, FDI, cg),
run through the UNMODIFIED GeologicalModel at a reduced nelements so that the fixture stays small.

    python tests/golden/make_golden_claudius.py      (build container only)

Stores what reaches the interpolator (support geometry, constraint arrays in local coordinates, weights as the
feature builder injects them, tolerance) and what comes out (c, iterations, evaluations in LOCAL coordinates).
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


def main(nelements=2e4, tag="claudius_2e4"):
    data, bb = load_claudius()
    model = GeologicalModel(bb[0, :], bb[1, :])
    model.data = data
    its = []
    model.create_and_add_foliation("strati", nelements=nelements, solver_kwargs={"callback": lambda xk: its.append(1)})
    model.update()
    feat = model["strati"]
    it = feat.interpolator
    sup = it.support
    c = np.array(it.c)
    rng = np.random.default_rng(0)
    lo, hi = np.asarray(sup.origin, float), np.asarray(sup.maximum, float)
    q = lo + rng.uniform(-0.05, 1.05, (300, 3)) * (hi - lo)
    val = sup.evaluate_value(q, c)
    grad = sup.evaluate_gradient(q, c)
    kw = {k: float(v) for k, v in it.interpolation_weights.items() if isinstance(v, (int, float))}
    timing = it.get_last_solve_timing()
    solver_kwargs = dict(feat.builder.build_arguments.get("solver_kwargs", {}))
    tol = feat.builder.build_arguments.get("tol", None)
    np.savez_compressed(
        os.path.join(HERE, f"fdi_{tag}.npz"), ncells=np.asarray(sup.nsteps_cells), origin=lo,
        step=np.asarray(sup.step_vector, float), c=c, iterations=len(its), scaling=bool(it.apply_scaling_matrix),
        tol=-1.0 if tol is None else float(tol), query=q, value_at=val, gradient_at=grad,
        kw_keys=np.array(list(kw.keys())), kw_vals=np.array(list(kw.values()), float),
        data_value=it.get_value_constraints(), data_normal=it.get_norm_constraints(),
        data_gradient=it.get_gradient_constraints(), data_tangent=it.get_tangent_constraints(),
        data_interface=it.get_interface_constraints(), matrix_rows=timing.get("matrix_rows", -1),
        families=np.array(list(it.constraints.keys())))
    print(tag, "nodes", sup.nsteps, "dof", it.dof, "its", len(its), "tol", tol, "skw", {k: v for k, v in solver_kwargs.items() if k != "callback"},
          "families", list(it.constraints.keys()), "n_value", it.get_value_constraints().shape, "n_normal", it.get_norm_constraints().shape,
          "range", float(c.max() - c.min()), "weights", kw)


if __name__ == "__main__":
    main()
