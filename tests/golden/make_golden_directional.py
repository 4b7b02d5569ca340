"""Golden fixtures for the directional regularisation of FiniteDifferenceInterpolator, from the UNMODIFIED reference.

    python tests/golden/make_golden_directional.py      (build container only)

Writes tests/golden/fdi_directional_*.npz: the families `minimise_directional_gradient_change`
(LI/_finite_difference_interpolator.py:1207-1340) puts into the constraint dict for an analytic direction field sampled at
the nodes (with a plane of zero vectors and one zero component, so that rows are dropped), and the field itself.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import dump_families, make_interp  # noqa: E402  (sets up the reference import path)


def direction(x, seed):
    r = np.random.default_rng(seed)
    k, p = r.uniform(0.3, 1.1, 3), r.uniform(0.0, 3.0, 3)
    v = np.column_stack([np.sin(k[0] * x[:, 1] + p[0]) + 0.3, np.cos(k[1] * x[:, 2] + p[1]) - 0.2 * x[:, 0],
                         0.5 + np.sin(k[2] * x[:, 0] + p[2]) * np.cos(0.5 * x[:, 1])])
    return v


def case(tag, ncells, origin, step, w, seed):
    it = make_interp(ncells, origin, step)
    it.reset()
    nodes = np.asarray(it.support.nodes, float)
    v = direction(nodes, seed)
    zmid = nodes[:, 2] == nodes[:, 2][len(nodes) // 2]
    v[zmid] = 0.0                      # a whole plane without direction: its rows are dropped
    v[nodes[:, 0] > nodes[:, 0].mean(), 1] = 0.0   # vy = 0 on half of the grid: dyy, dxy, dyz rows dropped there
    it.minimise_directional_gradient_change(w, v, name="fold regularisation 1")
    out = dump_families(it)
    out.update(ncells=np.array(ncells), origin=np.array(origin, float), step=np.array(step, float), vectors=v, w=w)
    np.savez_compressed(os.path.join(HERE, f"fdi_directional_{tag}.npz"), **out)
    print("fdi directional", tag, {k: it.constraints[k]["matrix"].shape for k in it.constraints})


if __name__ == "__main__":
    case("a", (6, 5, 7), (0.0, 0.0, 0.0), (1.0, 1.0, 1.0), 0.3, 1)
    case("b", (5, 8, 4), (-2.0, 1.0, 0.5), (0.8, 1.25, 1.1), 1.7, 2)
