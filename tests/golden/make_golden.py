"""Generate golden fixtures by running the UNMODIFIED reference (LoopStructural 1.8.0).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Writes tests/golden/fdi_*.npz.  The reference is imported from /root/reference with three
import stubs (tests/golden/_stubs: pyvista, skimage.measure, LoopStructural dist-info), see
SURVEY.md section 8c.  Nothing here is imported by the product or by the GPU tests.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.dont_write_bytecode = True
for p in (
    os.path.join(HERE, "_stubs"),
    f"{REF}/packages/loop_common/src",
    f"{REF}/packages/loop_interpolation/src",
    REF,
):
    if p not in sys.path:
        sys.path.insert(0, p)

import logging  # noqa: E402

logging.disable(logging.CRITICAL)

from loop_common.supports import StructuredGrid  # noqa: E402
from loop_interpolation import FiniteDifferenceInterpolator  # noqa: E402
from scipy import sparse  # noqa: E402


def field(p):
    return p[:, 2] + 0.1 * np.sin(2 * np.pi * p[:, 0]) + 0.05 * np.cos(2 * np.pi * p[:, 1])


def grad(p):
    return np.stack([0.2 * np.pi * np.cos(2 * np.pi * p[:, 0]), -0.1 * np.pi * np.sin(2 * np.pi * p[:, 1]),
                     np.ones(p.shape[0])], axis=1)


def make_interp(ncells, origin, step):
    g = StructuredGrid(origin=np.array(origin, float), nsteps_cells=np.array(ncells), step_vector=np.array(step, float))
    return FiniteDifferenceInterpolator(g)


def dump_families(it):
    out = {}
    names = list(it.constraints.keys())
    out["names"] = np.array(names)
    for n in names:
        c = it.constraints[n]
        m = c["matrix"].tocoo()
        out[f"{n}/rows"] = m.row.astype(np.int64)
        out[f"{n}/cols"] = m.col.astype(np.int64)
        out[f"{n}/vals"] = m.data
        out[f"{n}/shape"] = np.array(m.shape)
        out[f"{n}/b"] = np.asarray(c["b"], float)
        out[f"{n}/w"] = np.asarray(c["w"], float)
    return out


def case_operator(tag, nodes, weights=None, seed=0):
    """Regularisation-only system: explicit R^T R applied to random vectors + its diagonal."""
    rng = np.random.default_rng(seed)
    ncells = [n - 1 for n in nodes]
    it = make_interp(ncells, [0, 0, 0], [1.0, 1.0, 1.0])
    kw = dict(weights or {})
    it.setup_interpolator(**kw)
    A, b = it.build_matrix()
    A = sparse.csr_matrix(A)
    N = (A.T @ A).tocsr()
    x = rng.standard_normal((3, it.dof))
    y = np.stack([N @ v for v in x])
    np.savez_compressed(os.path.join(HERE, f"fdi_operator_{tag}.npz"), nodes=np.array(nodes), x=x, y=y,
                        diag=N.diagonal(), n_rows=A.shape[0], nnz=A.nnz,
                        weights_keys=np.array(list(kw.keys())), weights_vals=np.array(list(kw.values()), float))
    print("operator", tag, A.shape, A.nnz)


def case_rows(tag, ncells, origin, step, seed=1):
    """Rows produced for every data family, incl. points outside and on the faces."""
    rng = np.random.default_rng(seed)
    it = make_interp(ncells, origin, step)
    lo = np.array(origin, float)
    hi = lo + np.array(ncells) * np.array(step, float)
    ext = hi - lo

    def cloud(n):
        p = lo + rng.uniform(-0.05, 1.05, (n, 3)) * ext
        p[0] = lo  # corners / faces
        p[1] = hi
        p[2] = [hi[0], lo[1] + 0.5 * ext[1], lo[2] + 0.25 * ext[2]]
        p[3] = lo + np.array(step) * np.array([2, 3, 1])  # exactly on a node
        return p

    pv = cloud(40)
    value = np.column_stack([pv, field(pv), rng.uniform(0.5, 2.0, 40)])
    pg = cloud(30)
    gradient = np.column_stack([pg, grad(pg) * rng.uniform(0.5, 3.0, (30, 1)), np.ones(30)])
    gradient[5, 3:6] = [0.0, 0.0, 2.0]  # vertical normal: strike undefined branch
    pn = cloud(25)
    normal = np.column_stack([pn, grad(pn), rng.uniform(0.5, 2.0, 25)])
    pt = cloud(20)
    tangent = np.column_stack([pt, np.cross(grad(pt), [1.0, 0.0, 0.0]), np.ones(20)])
    pi_ = cloud(24)
    interface = np.column_stack([pi_, np.repeat([0.0, 1.0, 2.0], 8), np.ones(24)])
    it.set_value_constraints(value.copy())
    it.set_gradient_constraints(gradient.copy())
    it.set_normal_constraints(normal.copy())
    it.set_tangent_constraints(tangent.copy())
    it.set_interface_constraints(interface.copy())
    it.setup_interpolator(cpw=1.5, gpw=0.7, npw=3.0, tpw=0.9, ipw=1.1, dxx=0.0, dyy=0.0, dzz=0.0, dxy=0.0,
                          dyz=0.0, dxz=0.0, dx=0.0, dy=0.0, dz=0.0)
    out = dump_families(it)
    out.update(ncells=np.array(ncells), origin=np.array(origin, float), step=np.array(step, float), value=value,
               gradient=gradient, normal=normal, tangent=tangent, interface=interface,
               kw_keys=np.array(["cpw", "gpw", "npw", "tpw", "ipw"]), kw_vals=np.array([1.5, 0.7, 3.0, 0.9, 1.1]))
    np.savez_compressed(os.path.join(HERE, f"fdi_rows_{tag}.npz"), **out)
    print("rows", tag, list(it.constraints.keys()))


def case_solve(tag, ncells, origin, step, n_value, n_gradient, n_normal=0, n_tangent=0, scaling=True, seed=2,
               rtol=None, weights=None, tol=None):
    rng = np.random.default_rng(seed)
    it = make_interp(ncells, origin, step)
    lo = np.array(origin, float)
    ext = np.array(ncells) * np.array(step, float)
    unit = lambda n: rng.uniform(0, 1, (n, 3))  # noqa: E731
    pv = unit(n_value)
    value = np.column_stack([lo + pv * ext, field(pv), np.ones(n_value)])
    it.set_value_constraints(value.copy())
    data = dict(value=value)
    if n_gradient:
        pg = unit(n_gradient)
        gradient = np.column_stack([lo + pg * ext, grad(pg) / ext, np.ones(n_gradient)])
        it.set_gradient_constraints(gradient.copy())
        data["gradient"] = gradient
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
    residuals = []
    skw = {"callback": lambda xk: residuals.append(1)}
    if rtol is not None:
        skw["rtol"] = rtol
        skw["atol"] = 0.0
    ok = it.solve_system("cg", tol=tol, solver_kwargs=skw)
    assert ok
    c = np.array(it.c)
    # evaluation cloud: inside, outside, on faces, NaN row
    q = lo + rng.uniform(-0.1, 1.1, (200, 3)) * ext
    q[0] = lo
    q[1] = lo + ext
    q[2] = lo + ext * [1.0, 0.5, 0.0]
    q[3] = [np.nan, lo[1], lo[2]]
    val = it.evaluate_value(q)
    g = it.evaluate_gradient(q[np.all(np.isfinite(q), axis=1)])
    timing = it.get_last_solve_timing()
    np.savez_compressed(os.path.join(HERE, f"fdi_solve_{tag}.npz"), ncells=np.array(ncells), origin=lo,
                        step=np.array(step, float), c=c, iterations=len(residuals), scaling=scaling,
                        rtol=-1.0 if rtol is None else rtol, tol=-1.0 if tol is None else tol, query=q, value_at=val,
                        gradient_at=g, kw_keys=np.array(list(kw.keys())), kw_vals=np.array(list(kw.values()), float),
                        matrix_rows=timing["matrix_rows"], matrix_nnz=timing["matrix_nnz"],
                        **{f"data_{k}": v for k, v in data.items()})
    print("solve", tag, "its", len(residuals), "rows", timing["matrix_rows"], "range", float(c.max() - c.min()))


if __name__ == "__main__":
    case_operator("6x6x6", (6, 6, 6))
    case_operator("9x7x13", (9, 7, 13))
    case_operator("15x15x15", (15, 15, 15))
    case_operator("3x5x4", (3, 5, 4))
    case_operator("9x7x13_w", (9, 7, 13), weights=dict(dxx=2.0, dyy=0.5, dzz=1.5, dxy=3.0, dyz=0.25, dxz=1.25, dx=0.3,
                                                       dy=2.5, dz=1.7))
    case_rows("a", [7, 5, 6], [0.0, 0.0, 0.0], [1.0, 1.0, 1.0])
    case_rows("b", [6, 9, 4], [-3.0, 10.0, 2.5], [0.5, 0.25, 2.0])
    case_solve("unit16", [15, 15, 15], [0, 0, 0], [1 / 15] * 3, 300, 40)
    case_solve("unit16_noscale", [15, 15, 15], [0, 0, 0], [1 / 15] * 3, 300, 40, scaling=False)
    case_solve("aniso", [20, 14, 11], [-5.0, 3.0, 100.0], [0.7, 1.1, 0.4], 400, 60, n_normal=30, n_tangent=25,
               weights=dict(regularisation=0.7, npw=2.0), tol=1e-9)
    case_solve("unit24_tight", [23, 23, 23], [0, 0, 0], [1 / 23] * 3, 800, 100, rtol=1e-10)
