"""Golden fixtures for the P1 / TetMesh twin of the path, from the UNMODIFIED reference.

    python tests/golden/make_golden_p1.py      (build container only)

Writes tests/golden/p1_*.npz: the explicit face-jump regularisation operator, the value / norm / tangent
families as the reference's own sparse matrices, full solves, and evaluation queries (incl. points on faces,
outside, NaN).  Same import recipe as make_golden.py.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import dump_families, field, grad  # noqa: E402  (sets up the reference import path)

from loop_common.supports import TetMesh  # noqa: E402
from loop_interpolation import P1Interpolator  # noqa: E402
from scipy import sparse  # noqa: E402


def make_interp(ncells, origin, step):
    m = TetMesh(origin=np.array(origin, float), nsteps_cells=np.array(ncells), step_vector=np.array(step, float))
    return P1Interpolator(m)


def case_operator(tag, ncells, step=(1.0, 1.0, 1.0), cgw=0.1, seed=0):
    rng = np.random.default_rng(seed)
    it = make_interp(ncells, [0, 0, 0], step)
    it.setup_interpolator(cgw=cgw)
    fam = it.constraints["edge jump"]
    M = sparse.csr_matrix(fam["matrix"])
    w = np.asarray(fam["w"], float)
    N = (M.T @ sparse.diags(w * w) @ M).tocsr()
    x = rng.standard_normal((3, it.dof))
    y = np.stack([N @ v for v in x])
    np.savez_compressed(os.path.join(HERE, f"p1_operator_{tag}.npz"), ncells=np.array(ncells), step=np.array(step, float),
                        cgw=cgw, x=x, y=y, diag=N.diagonal(), n_rows=M.shape[0], nnz=M.nnz)
    print("p1 operator", tag, M.shape, M.nnz)


def cloud(rng, lo, ext, step, n):
    p = lo + rng.uniform(-0.05, 1.05, (n, 3)) * ext
    p[0] = lo + 0.5 * np.array(step)            # cell centre: on the central tet
    p[1] = lo + ext - 1e-9 * ext                # just inside the far corner
    p[2] = lo + np.array(step) * [1.0, 1.5, 0.5]   # on a cell face
    p[3] = lo + np.array(step) * [2.0, 1.0, 1.0]   # exactly on a node
    p[4] = lo + np.array(step) * [1.25, 1.25, 1.5]  # on an internal tet face / edge
    p[5] = lo                                     # the origin corner: strict `inside` drops it
    return p


def case_rows(tag, ncells, origin, step, seed=1):
    rng = np.random.default_rng(seed)
    it = make_interp(ncells, origin, step)
    lo = np.array(origin, float)
    ext = np.array(ncells) * np.array(step, float)
    pv = cloud(rng, lo, ext, step, 40)
    value = np.column_stack([pv, field(pv), rng.uniform(0.5, 2.0, 40)])
    pn = cloud(rng, lo, ext, step, 25)
    normal = np.column_stack([pn, grad(pn), rng.uniform(0.5, 2.0, 25)])
    pt = cloud(rng, lo, ext, step, 20)
    tangent = np.column_stack([pt, np.cross(grad(pt), [1.0, 0.0, 0.0]), np.ones(20)])
    it.set_value_constraints(value.copy())
    it.set_normal_constraints(normal.copy())
    it.set_tangent_constraints(tangent.copy())
    it.setup_interpolator(cgw=0.0, cpw=1.5, npw=3.0, tpw=0.9)
    out = dump_families(it)
    out.update(ncells=np.array(ncells), origin=lo, step=np.array(step, float), value=value, normal=normal,
               tangent=tangent, kw_keys=np.array(["cpw", "npw", "tpw"]), kw_vals=np.array([1.5, 3.0, 0.9]))
    np.savez_compressed(os.path.join(HERE, f"p1_rows_{tag}.npz"), **out)
    print("p1 rows", tag, {k: it.constraints[k]["matrix"].shape for k in it.constraints})


def case_solve(tag, ncells, origin, step, n_value, n_normal, n_tangent=0, scaling=True, seed=2, rtol=None, tol=None,
               weights=None):
    rng = np.random.default_rng(seed)
    it = make_interp(ncells, origin, step)
    lo = np.array(origin, float)
    ext = np.array(ncells) * np.array(step, float)
    unit = lambda n: rng.uniform(0.01, 0.99, (n, 3))  # noqa: E731
    pv = unit(n_value)
    value = np.column_stack([lo + pv * ext, field(pv), np.ones(n_value)])
    it.set_value_constraints(value.copy())
    data = dict(value=value)
    if n_normal:
        pn = unit(n_normal)
        normal = np.column_stack([lo + pn * ext, grad(pn) / ext, np.ones(n_normal)])
        it.set_normal_constraints(normal.copy())
        data["normal"] = normal
    if n_tangent:
        pt = unit(n_tangent)
        tangent = np.column_stack([lo + pt * ext, np.cross(grad(pt) / ext, [1.0, 0.0, 0.0]), np.ones(n_tangent)])
        it.set_tangent_constraints(tangent.copy())
        data["tangent"] = tangent
    it.apply_scaling_matrix = scaling
    kw = dict(weights or {})
    it.setup_interpolator(**kw)
    its = []
    skw = {"callback": lambda xk: its.append(1)}
    if rtol is not None:
        skw["rtol"] = rtol
        skw["atol"] = 0.0
    assert it.solve_system("cg", tol=tol, solver_kwargs=skw)
    c = np.array(it.c)
    q = lo + rng.uniform(-0.1, 1.1, (300, 3)) * ext
    q[:6] = cloud(rng, lo, ext, step, 6)
    q[6] = [np.nan, lo[1], lo[2]]
    val = it.evaluate_value(q)
    fin = np.all(np.isfinite(q), axis=1)
    g = it.evaluate_gradient(q[fin])
    np.savez_compressed(os.path.join(HERE, f"p1_solve_{tag}.npz"), ncells=np.array(ncells), origin=lo,
                        step=np.array(step, float), c=c, iterations=len(its), scaling=scaling,
                        rtol=-1.0 if rtol is None else rtol, tol=-1.0 if tol is None else tol, query=q, value_at=val,
                        gradient_at=g, kw_keys=np.array(list(kw.keys())), kw_vals=np.array(list(kw.values()), float),
                        **{f"data_{k}": v for k, v in data.items()})
    print("p1 solve", tag, "its", len(its), "range", float(c.max() - c.min()))


if __name__ == "__main__":
    case_operator("3x3x3", (3, 3, 3))
    case_operator("1x2x1", (1, 2, 1))
    case_operator("4x5x3", (4, 5, 3), step=(0.5, 1.25, 2.0), cgw=0.7)
    case_operator("6x5x7", (6, 5, 7))
    case_operator("8x8x8", (8, 8, 8), step=(0.125, 0.125, 0.125), cgw=1.0)
    case_rows("a", [5, 4, 6], [0.0, 0.0, 0.0], [1.0, 1.0, 1.0])
    case_rows("b", [4, 7, 3], [-3.0, 10.0, 2.5], [0.5, 0.25, 2.0])
    case_solve("unit12", [12, 12, 12], [0, 0, 0], [1 / 12] * 3, 200, 30)
    case_solve("aniso", [14, 9, 11], [-5.0, 3.0, 100.0], [0.7, 1.1, 0.4], 300, 40, n_tangent=25,
               weights=dict(regularisation=0.4, npw=2.0), tol=1e-9)
    case_solve("unit16_tight", [16, 15, 17], [0, 0, 0], [1 / 16, 1 / 15, 1 / 17], 500, 60, rtol=1e-10, scaling=False)
