"""CPU oracle (numpy) of the P1 / TetMesh twin of the hot path.  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's CPU legs, never by the product.

Restates, with explicit arrays and brute-force topology:
  TetMesh (LC/supports/_3d_structured_tetra.py): 5 tets per cell with parity masks (:28-34), element order
    (:376-419), point location with the "last matching tet" tie-break (:274-365), strict inside (:264-272),
    element gradients (M^-1)^T I (:421-486), element_size (:85-100);
  P1Interpolator (LI/_p1interpolator.py): value rows (:98-113), norm rows (:73-96, flattened (n,3,4) block),
    gradient-orthogonal / tangent rows with w * tet volume (:248-296), face gradient-jump regularisation
    minimise_edge_jumps (:115-166) over every interior shared face;
  row normalisation of add_constraints_to_least_squares (LI/_discrete_interpolator.py:361-366).
  DiscreteFoldInterpolator.add_fold_constraints (LI/_discrete_fold_interpolator.py:116-242): one row vol (v . eg) = vol b
    per selected tetrahedron (element_vector_rows) and direction-weighted face-jump rows (face_vector_rows:
    minimise_edge_jumps with a vector per shared face, LI/_p1interpolator.py:115-166).
Pinned against tests/golden/p1_*.npz and fold_rows_*.npz, which tests/golden/make_golden_p1.py / make_golden_fold.py
wrote from the unmodified reference.
"""
from __future__ import annotations

import numpy as np

from .fdi_oracle import Family, Grid, make_family

TET_EVEN = np.array([[7, 1, 2, 4], [6, 2, 4, 7], [5, 1, 4, 7], [0, 1, 2, 4], [3, 1, 2, 7]])
TET_ODD = np.array([[0, 6, 5, 3], [7, 3, 5, 6], [4, 0, 5, 6], [2, 0, 3, 6], [1, 0, 3, 5]])
CORNER = np.array([[c & 1, (c >> 1) & 1, (c >> 2) & 1] for c in range(8)])
I34 = np.array([[-1.0, 1.0, 0.0, 0.0], [-1.0, 0.0, 1.0, 0.0], [-1.0, 0.0, 0.0, 1.0]])
P1_DEFAULT_WEIGHTS = dict(cgw=0.1, cpw=1.0, npw=1.0, gpw=1.0, tpw=1.0, ipw=1.0)


def cell_tets(grid: Grid, cells: np.ndarray) -> np.ndarray:
    """(n,5,4) global node ids of the five tets of each cell (n,3)."""
    nx, ny = int(grid.nsteps[0]), int(grid.nsteps[1])
    corner = cells[:, None, :] + CORNER[None, :, :]
    gi = corner[..., 0] + nx * corner[..., 1] + nx * ny * corner[..., 2]
    even = cells.sum(axis=1) % 2 == 0
    out = np.empty((cells.shape[0], 5, 4), dtype=np.int64)
    out[even] = gi[even][:, TET_EVEN]
    out[~even] = gi[~even][:, TET_ODD]
    return out


def elements(grid: Grid) -> np.ndarray:
    ncx, ncy, ncz = (int(v) for v in grid.nsteps - 1)
    zz, yy, xx = np.meshgrid(np.arange(ncz), np.arange(ncy), np.arange(ncx), indexing="ij")
    cells = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1)
    return cell_tets(grid, cells).reshape(-1, 4)


def node_xyz(grid: Grid, ids: np.ndarray) -> np.ndarray:
    nx, ny = int(grid.nsteps[0]), int(grid.nsteps[1])
    k = ids // (nx * ny)
    j = (ids - k * nx * ny) // nx
    i = ids - k * nx * ny - j * nx
    return grid.origin + np.stack([i, j, k], axis=-1) * grid.step


def tet_gradients(verts: np.ndarray) -> np.ndarray:
    """(n,4,3) vertices -> (n,3,4) shape-function gradients."""
    E = verts[:, 1:, :] - verts[:, :1, :]
    return np.linalg.inv(E) @ I34


def tet_volume(verts: np.ndarray) -> np.ndarray:
    return np.abs(np.linalg.det(verts[:, 1:, :] - verts[:, :1, :])) / 6.0


def face_rows(grid: Grid, w: float) -> Family:
    """minimise_edge_jumps: one row n . (grad phi_T1 - grad phi_T2) = 0 per interior shared face."""
    el = elements(grid)
    nt = el.shape[0]
    drop = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])
    faces = np.sort(el[:, drop].reshape(-1, 3), axis=1)
    owner = np.repeat(np.arange(nt), 4)
    order = np.lexsort((faces[:, 2], faces[:, 1], faces[:, 0]))
    fs, ow = faces[order], owner[order]
    same = np.all(fs[1:] == fs[:-1], axis=1)
    t1, t2, shared = ow[:-1][same], ow[1:][same], fs[:-1][same]
    xyz = node_xyz(grid, shared)
    n = np.cross(xyz[:, 1] - xyz[:, 0], xyz[:, 2] - xyz[:, 0])
    g1 = tet_gradients(node_xyz(grid, el[t1]))
    g2 = tet_gradients(node_xyz(grid, el[t2]))
    A = np.hstack([np.einsum("ij,ijk->ik", n, g1), -np.einsum("ij,ijk->ik", n, g2)])
    idx = np.hstack([el[t1], el[t2]])
    return make_family("edge jump", A, np.zeros(A.shape[0]), idx, w)


def locate(grid: Grid, pts: np.ndarray):
    """get_element_for_location: (barycentric c (n,4), node ids (n,4), tet vertices (n,4,3), tet id, inside)."""
    pts = np.asarray(pts, dtype=float)
    nc = grid.nsteps - 1
    maximum = grid.origin + nc * grid.step
    tol = np.finfo(float).eps * 32.0 * np.maximum(np.abs(maximum - grid.origin), 1.0)
    with np.errstate(invalid="ignore"):
        ins_tol = np.all((pts >= grid.origin - tol) & (pts <= maximum + tol), axis=1)
        ins_strict = np.all((pts > grid.origin) & (pts < maximum), axis=1)
    cells = np.zeros((pts.shape[0], 3), dtype=np.int64)
    with np.errstate(invalid="ignore"):
        f = np.floor((pts[ins_tol] - grid.origin) / grid.step).astype(np.int64)
    cells[ins_tol] = np.clip(f, 0, nc - 1)
    tets = cell_tets(grid, cells)                     # (n,5,4)
    v = node_xyz(grid, tets)                          # (n,5,4,3)
    vap = pts[:, None, :] - v[:, :, 0]
    vbp = pts[:, None, :] - v[:, :, 1]
    vab, vac, vad = v[:, :, 1] - v[:, :, 0], v[:, :, 2] - v[:, :, 0], v[:, :, 3] - v[:, :, 0]
    vbc, vbd = v[:, :, 2] - v[:, :, 1], v[:, :, 3] - v[:, :, 1]
    dot = lambda a, b: np.einsum("ikj,ikj->ik", a, b)  # noqa: E731
    with np.errstate(invalid="ignore", divide="ignore"):
        vol = dot(vab, np.cross(vac, vad)) / 6.0
        c = np.stack([dot(vbp, np.cross(vbd, vbc)) / 6.0 / vol, dot(vap, np.cross(vac, vad)) / 6.0 / vol,
                      dot(vap, np.cross(vad, vab)) / 6.0 / vol, dot(vap, np.cross(vab, vac)) / 6.0 / vol], axis=2)
        ok = np.all(c >= 0, axis=2)
    # the last tet whose barycentrics are all >= 0 wins (dict(zip(i, j)) keeps the last j)
    local = np.where(ok.any(axis=1), 4 - np.argmax(ok[:, ::-1], axis=1), -1)
    inside = ins_tol & ins_strict & (local >= 0)
    sel = np.clip(local, 0, 4)
    r = np.arange(pts.shape[0])
    ncx, ncy = int(nc[0]), int(nc[1])
    tet_id = sel + 5 * (cells[:, 0] + ncx * cells[:, 1] + ncx * ncy * cells[:, 2])
    return c[r, sel], tets[r, sel], v[r, sel], tet_id, inside


def value_rows(grid: Grid, pts: np.ndarray, cpw: float = 1.0):
    if pts.shape[0] == 0:
        return None
    c, ids, _v, _t, ins = locate(grid, pts[:, :3])
    return make_family("value", c[ins], pts[ins, 3], ids[ins], cpw)  # the per-point weight column is not used


def normal_rows(grid: Grid, pts: np.ndarray, npw: float = 1.0):
    if pts.shape[0] == 0:
        return None
    _c, ids, v, _t, ins = locate(grid, pts[:, :3])
    g = tet_gradients(v[ins])                                   # (n,3,4)
    A = g.reshape(-1, 4)                                        # point-major flattening of the (n,3,4) block
    idx = np.repeat(ids[ins], 3, axis=0)
    return make_family("norm", A, pts[ins, 3:6].reshape(-1), idx, npw)


def orthogonal_rows(grid: Grid, xyz: np.ndarray, vectors: np.ndarray, w: float = 1.0, b: float = 0.0,
                    name: str = "undefined gradient orthogonal constraint"):
    if xyz.shape[0] == 0:
        return None
    _c, ids, v, _t, ins = locate(grid, xyz[:, :3])
    vec = np.array(vectors, dtype=float)
    nrm = np.linalg.norm(vec, axis=1)
    vec[nrm > 0] /= nrm[nrm > 0, None]
    g = tet_gradients(v[ins])
    A = np.einsum("ij,ijk->ik", vec[ins], g)
    return make_family(name, A, np.zeros(int(ins.sum())) + b, ids[ins], w * tet_volume(v[ins]))


def element_vector_rows(grid: Grid, xyz: np.ndarray, vectors: np.ndarray, w: float = 1.0, b: float = 0.0,
                        name: str = "fold orientation"):
    """LI/_discrete_fold_interpolator.py:116-190 for the tetrahedra that hold ``xyz`` (their barycentres in the reference):
    A = vol (v . eg), B = vol b, weight w -- neither is v normalised nor the weight scaled by the volume."""
    if xyz.shape[0] == 0:
        return None
    _c, ids, v, _t, ins = locate(grid, xyz[:, :3])
    vol = tet_volume(v[ins])
    A = np.einsum("ij,ijk->ik", np.asarray(vectors, float)[ins], tet_gradients(v[ins])) * vol[:, None]
    return make_family(name, A, (np.zeros(int(ins.sum())) + b) * vol, ids[ins], w)


def face_vector_rows(grid: Grid, nodes_a: np.ndarray, nodes_b: np.ndarray, vectors: np.ndarray, w: float,
                     name: str = "directional edge jump"):
    """minimise_edge_jumps with ``vector`` (LI/_p1interpolator.py:115-166): v . (grad phi_a - grad phi_b) = 0 over the 4 + 4
    listed nodes of two face-adjacent tetrahedra given by their global node ids."""
    g1 = tet_gradients(node_xyz(grid, np.asarray(nodes_a)))
    g2 = tet_gradients(node_xyz(grid, np.asarray(nodes_b)))
    v = np.asarray(vectors, float)
    A = np.hstack([np.einsum("ij,ijk->ik", v, g1), -np.einsum("ij,ijk->ik", v, g2)])
    return make_family(name, A, np.zeros(A.shape[0]), np.hstack([nodes_a, nodes_b]), w)


def evaluate_value(grid: Grid, pts: np.ndarray, cvec: np.ndarray) -> np.ndarray:
    pts = np.asarray(pts, dtype=float)
    out = np.zeros(pts.shape[0])
    fin = ~np.any(np.isnan(pts), axis=1)                       # _evaluate_value_local: NaN rows give 0
    c, ids, _v, _t, ins = locate(grid, pts[fin])
    val = np.full(int(fin.sum()), np.nan)
    val[ins] = np.sum(c[ins] * cvec[ids[ins]], axis=1)
    out[fin] = val
    return out


def evaluate_gradient(grid: Grid, pts: np.ndarray, cvec: np.ndarray) -> np.ndarray:
    pts = np.asarray(pts, dtype=float)
    _c, ids, v, _t, ins = locate(grid, pts)
    out = np.full(pts.shape, np.nan)
    g = tet_gradients(v[ins])
    out[ins] = np.einsum("ijk,ik->ij", g, cvec[ids[ins]])
    return out
