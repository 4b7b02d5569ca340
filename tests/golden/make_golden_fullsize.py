"""Golden fixtures of the BASELINE configurations at (or near) their stated sizes, from the UNMODIFIED reference.

    python tests/golden/make_golden_fullsize.py [c1] [c2] [c3]      (build container only; c3 needs ~18 GB and ~5 min)

  c1  configs[0]: README quickstart, Claudius 'strati' foliation through GeologicalModel at nelements=1e5 (175,780 nodes),
      default tolerance, and the same system at rtol 1e-10 (where two correct solvers must agree to the 1e-6 bar).
  c2  configs[1] at 128^3 nodes (the largest size the reference solves in about half a minute; 256^3 needs 5 GB for A^T A
      and ~1 s per iteration): synthetic problem of SURVEY.md 8(d), counts scaled with the grid.
  c3  configs[2] at full size: P1Interpolator on TetMesh 117^3 cells (8,008,065 tets), 1e5 normal + 1e5 tangent + 100 value.

The inputs are functions of a seed (`fullsize_inputs.py`, shared with the tests), so a fixture holds only what comes OUT:
the node values on a random subsample of nodes, the iteration count, and evaluation probes.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import make_golden  # noqa: E402,F401  (sets up the reference import path and stubs)
import fullsize_inputs as fi  # noqa: E402


def subsample(c, n, seed=99):
    idx = np.sort(np.random.default_rng(seed).choice(c.size, size=min(n, c.size), replace=False))
    return idx.astype(np.int64), np.asarray(c)[idx]


def counted():
    its = []
    return its, {"callback": lambda xk: its.append(1)}


def c1():
    from LoopStructural import GeologicalModel
    from LoopStructural.datasets import load_claudius

    data, bb = load_claudius()
    model = GeologicalModel(bb[0, :], bb[1, :])
    model.data = data
    its, skw = counted()
    t0 = time.perf_counter()
    model.create_and_add_foliation("strati", nelements=1e5, solver_kwargs=skw)
    model.update()
    dt = time.perf_counter() - t0
    it = model["strati"].interpolator
    sup = it.support
    c = np.array(it.c)
    tol = model["strati"].builder.build_arguments.get("tol", None)
    rng = np.random.default_rng(0)
    q = bb[0] + rng.uniform(-0.02, 1.02, (2000, 3)) * (bb[1] - bb[0])  # WORLD coordinates
    val = model.evaluate_feature_value("strati", q)
    grad = model.evaluate_feature_gradient("strati", q)
    # the same assembled system, solved tightly by the stock solver
    its2, skw2 = counted()
    skw2.update(rtol=1e-10, atol=0.0)
    assert it.solve_system("cg", solver_kwargs=skw2)
    c_tight = np.array(it.c)
    idx, cs = subsample(c, 40000)
    np.savez_compressed(os.path.join(HERE, "full_c1_claudius_1e5.npz"), ncells=np.asarray(sup.nsteps_cells), origin=np.asarray(sup.origin, float),
                        step=np.asarray(sup.step_vector, float), tol=-1.0 if tol is None else float(tol), iterations=len(its),
                        iterations_tight=len(its2), idx=idx, c=cs, c_tight=c_tight[idx], c_range=float(c.max() - c.min()),
                        query_world=q, value_at=val, gradient_at=grad, bb=bb, seconds=dt)
    print("c1", "nodes", sup.nsteps, "its", len(its), "tight", len(its2), "tol", tol, f"{dt:.1f} s")


def c2(n=128):
    from loop_common.supports import StructuredGrid
    from loop_interpolation import FiniteDifferenceInterpolator

    value, gradient = fi.c2_points(n)
    sg = StructuredGrid(origin=np.zeros(3), nsteps_cells=np.array([n - 1] * 3), step_vector=np.full(3, 1.0 / (n - 1)))
    it = FiniteDifferenceInterpolator(sg)
    it.set_value_constraints(value.copy())
    it.set_gradient_constraints(gradient.copy())
    t0 = time.perf_counter()
    it.setup_interpolator()
    its, skw = counted()
    assert it.solve_system("cg", solver_kwargs=skw)
    dt = time.perf_counter() - t0
    c = np.array(it.c)
    q = fi.probe_points(3000, seed=5)
    val, grad = it.evaluate_value(q), it.evaluate_gradient(q)
    its2, skw2 = counted()
    skw2.update(rtol=1e-10, atol=0.0)
    assert it.solve_system("cg", solver_kwargs=skw2)
    c_tight = np.array(it.c)
    idx, cs = subsample(c, 100000)
    np.savez_compressed(os.path.join(HERE, f"full_c2_fdi_{n}.npz"), n=n, iterations=len(its), iterations_tight=len(its2), idx=idx, c=cs,
                        c_tight=c_tight[idx], c_range=float(c.max() - c.min()), query=q, value_at=val, gradient_at=grad, seconds=dt)
    print("c2", n, "its", len(its), "tight", len(its2), f"{dt:.1f} s")


def c3(nc=117):
    from loop_common.supports import TetMesh
    from loop_interpolation import P1Interpolator

    value, normal, tangent = fi.c3_points()
    mesh = TetMesh(origin=np.zeros(3), nsteps_cells=np.array([nc] * 3), step_vector=np.full(3, 1.0 / nc))
    it = P1Interpolator(mesh)
    it.set_value_constraints(value.copy())
    it.set_normal_constraints(normal.copy())
    it.set_tangent_constraints(tangent.copy())
    t0 = time.perf_counter()
    it.setup_interpolator(cgw=0.1)
    t1 = time.perf_counter()
    its, skw = counted()
    assert it.solve_system("cg", solver_kwargs=skw)
    dt = time.perf_counter() - t0
    c = np.array(it.c)
    q = fi.probe_points(3000, seed=6, lo=0.001, hi=0.999)
    val, grad = it.evaluate_value(q), it.evaluate_gradient(q)
    idx, cs = subsample(c, 100000)
    np.savez_compressed(os.path.join(HERE, f"full_c3_p1_{nc}.npz"), nc=nc, iterations=len(its), idx=idx, c=cs, c_range=float(c.max() - c.min()),
                        query=q, value_at=val, gradient_at=grad, seconds=dt, setup_seconds=t1 - t0,
                        n_tets=int(mesh.n_elements), n_nodes=int(mesh.n_nodes))
    print("c3", nc, "tets", mesh.n_elements, "its", len(its), f"setup {t1 - t0:.1f} s, total {dt:.1f} s")


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c2"]
    for w in which:
        {"c1": c1, "c2": c2, "c3": c3}[w]()
