"""Numpy/scipy restatement of the reference FiniteDifferenceInterpolator path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): the checker the CUDA
path is compared with, and the timed CPU baseline of ``bench.py``.  Never
imported by the product package.

Citations use the SURVEY.md prefixes
  LI/ = /root/reference/packages/loop_interpolation/src/loop_interpolation/
  LC/ = /root/reference/packages/loop_common/src/loop_common/
The CG iteration itself lives in scipy (pinned scipy 1.18.0 in
/root/reference/uv.lock:3141; 1.18.1 installed); ``cg`` below restates the
published algorithm of ``scipy.sparse.linalg.cg`` and is checked against the
installed scipy in tests/test_oracle_golden.py.

Parity status: PINNED -- every function below is checked against outputs of
the reference run in the build container (tests/golden/make_golden.py ->
tests/golden/fdi_*.npz).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
from scipy import sparse

# --------------------------------------------------------------------------
# support (StructuredGrid) index math
# --------------------------------------------------------------------------


@dataclass
class Grid:
    """Regular grid; ``nsteps`` are NODE counts (cells + 1), LC/supports/_3d_base_structured.py:66."""

    nsteps: np.ndarray
    origin: np.ndarray
    step: np.ndarray

    def __post_init__(self):
        self.nsteps = np.asarray(self.nsteps, dtype=np.int64)
        self.origin = np.asarray(self.origin, dtype=float)
        self.step = np.asarray(self.step, dtype=float)

    @classmethod
    def from_cells(cls, ncells, origin=(0.0, 0.0, 0.0), step=(1.0, 1.0, 1.0)):
        return cls(np.asarray(ncells, dtype=np.int64) + 1, origin, step)

    @property
    def n_nodes(self) -> int:
        return int(np.prod(self.nsteps))

    @property
    def ncells(self):
        return self.nsteps - 1

    @property
    def maximum(self):
        # LC/supports/_3d_base_structured.py:187-189
        return self.origin + self.ncells * self.step


def inside(grid: Grid, pts: np.ndarray) -> np.ndarray:
    """LC/supports/_3d_base_structured.py:287-293 (tolerance 32 eps max(extent,1))."""
    extent = np.maximum(np.abs(grid.maximum - grid.origin), 1.0)
    tol = np.finfo(float).eps * 32.0 * extent
    return np.all((pts >= grid.origin - tol) & (pts <= grid.maximum + tol), axis=1)


def cell_index(grid: Grid, pts: np.ndarray):
    """floor((p - origin)/step) clipped to [0, ncell-1]; LC/supports/_3d_base_structured.py:251-282."""
    ins = inside(grid, pts)
    ci = np.zeros((pts.shape[0], 3), dtype=np.int64)
    rel = pts - grid.origin[None, :]
    for a in range(3):
        v = np.floor(rel[ins, a] / grid.step[a]).astype(np.int64)
        ci[ins, a] = np.clip(v, 0, grid.ncells[a] - 1)
    return ci, ins


def cell_corners(grid: Grid, pts: np.ndarray):
    """8 corner node ids, corner c has offsets (c&1, (c>>1)&1, (c>>2)&1);
    LC/supports/_3d_base_structured.py:386-412.  -1 rows for outside points."""
    ci, ins = cell_index(grid, pts)
    nx, ny = int(grid.nsteps[0]), int(grid.nsteps[1])
    off = np.array([0, 1, nx, nx + 1, nx * ny, nx * ny + 1, nx * ny + nx, nx * ny + nx + 1], dtype=np.int64)
    g = ci[:, 0] + nx * ci[:, 1] + nx * ny * ci[:, 2]
    idx = g[:, None] + off[None, :]
    idx[~ins] = -1
    return idx, ins


def local_coords(grid: Grid, pts: np.ndarray) -> np.ndarray:
    """LC/supports/_3d_structured_grid.py:126-146 (clip to [0,1])."""
    ci, _ = cell_index(grid, pts)
    cell_origin = grid.origin[None, :] + grid.step[None, :] * ci
    return np.clip((pts - cell_origin) / grid.step[None, :], 0.0, 1.0)


def trilinear(lc: np.ndarray) -> np.ndarray:
    """LC/supports/_3d_structured_grid.py:90-124, corner order 000,100,010,110,001,101,011,111."""
    x, y, z = lc[:, 0], lc[:, 1], lc[:, 2]
    w = np.empty((lc.shape[0], 8))
    w[:, 0] = (1 - x) * (1 - y) * (1 - z)
    w[:, 1] = x * (1 - y) * (1 - z)
    w[:, 2] = (1 - x) * y * (1 - z)
    w[:, 3] = x * y * (1 - z)
    w[:, 4] = (1 - x) * (1 - y) * z
    w[:, 5] = x * (1 - y) * z
    w[:, 6] = (1 - x) * y * z
    w[:, 7] = x * y * z
    return w


def gradient_matrix(grid: Grid, lc: np.ndarray) -> np.ndarray:
    """T[n,3,8]: derivative of the trilinear shape functions / step;
    LC/supports/_3d_structured_grid.py:393-454."""
    x, y, z = lc[:, 0], lc[:, 1], lc[:, 2]
    T = np.zeros((lc.shape[0], 3, 8))
    T[:, 0, 0] = (1 - z) * (y - 1)
    T[:, 0, 1] = (1 - y) * (1 - z)
    T[:, 0, 2] = -y * (1 - z)
    T[:, 0, 3] = y * (1 - z)
    T[:, 0, 4] = -(1 - y) * z
    T[:, 0, 5] = (1 - y) * z
    T[:, 0, 6] = -y * z
    T[:, 0, 7] = y * z
    T[:, 1, 0] = (x - 1) * (1 - z)
    T[:, 1, 1] = -x * (1 - z)
    T[:, 1, 2] = (1 - x) * (1 - z)
    T[:, 1, 3] = x * (1 - z)
    T[:, 1, 4] = -(1 - x) * z
    T[:, 1, 5] = -x * z
    T[:, 1, 6] = (1 - x) * z
    T[:, 1, 7] = x * z
    T[:, 2, 0] = -(1 - x) * (1 - y)
    T[:, 2, 1] = -x * (1 - y)
    T[:, 2, 2] = -(1 - x) * y
    T[:, 2, 3] = -x * y
    T[:, 2, 4] = (1 - x) * (1 - y)
    T[:, 2, 5] = x * (1 - y)
    T[:, 2, 6] = (1 - x) * y
    T[:, 2, 7] = x * y
    for a in range(3):
        T[:, a, :] /= grid.step[a]
    return T


# --------------------------------------------------------------------------
# constraint rows
# --------------------------------------------------------------------------


@dataclass
class Family:
    """One entry of ``interpolator.constraints`` (LI/_discrete_interpolator.py:383-389):
    row-normalised coefficients ``A`` (n,k), column ids ``idx`` (n,k), targets ``b`` (n,), weights ``w`` (n,)."""

    name: str
    A: np.ndarray
    idx: np.ndarray
    b: np.ndarray
    w: np.ndarray

    def matrix(self, dof: int) -> sparse.csr_matrix:
        n, k = self.A.shape
        rows = np.repeat(np.arange(n), k)
        # duplicates are summed on conversion, as coo->csc does in the reference
        return sparse.coo_matrix((self.A.ravel(), (rows, self.idx.ravel())), shape=(n, dof)).tocsr()


def make_family(name, A, b, idx, w) -> Family:
    """Row ingestion rule of add_constraints_to_least_squares, LI/_discrete_interpolator.py:316-389:
    every row is divided by its L2 norm over the LISTED entries (before duplicate
    columns merge), b likewise; scalar w broadcasts."""
    A = np.array(A, dtype=float)
    b = np.array(b, dtype=float).reshape(-1)
    idx = np.array(idx, dtype=np.int64)
    length = np.linalg.norm(A, axis=1)
    nz = length > 0
    b[nz] /= length[nz]
    A[nz, :] /= length[nz, None]
    if np.isscalar(w):
        w = np.ones(A.shape[0]) * float(w)
    return Family(name, A, idx, b, np.asarray(w, dtype=float))


def value_rows(grid: Grid, pts: np.ndarray, cpw: float = 1.0) -> Family | None:
    """LI/_finite_difference_interpolator.py:236-269.  pts columns x,y,z,val,w (local coords)."""
    if pts.shape[0] == 0:
        return None
    idx, ins = cell_corners(grid, pts[:, :3])
    a = trilinear(local_coords(grid, pts[ins, :3]))
    return make_family("value", a, pts[ins, 3], idx[ins], cpw * pts[ins, 4])


def strike_dip_vectors(normal: np.ndarray):
    """get_vectors, LC/math/_maths.py:276-298 (+ :126-160, :301-332).  normal: (n,3) unit vectors.
    Returns (3,n) strike and dip vectors."""
    normal = np.array(normal, dtype=float)
    length = np.linalg.norm(normal, axis=1)[:, None]
    normal = normal / length
    phi = np.arctan2(normal[:, 1], normal[:, 0])  # == -strike
    # the reference round-trips the angle through degrees
    strike = np.rad2deg(-phi)
    s = np.deg2rad(strike)
    strike_vec = np.array([np.sin(-s), -np.cos(-s), np.zeros(s.shape[0])])
    strike_vec /= np.linalg.norm(strike_vec, axis=0)[None, :]
    dip_vec = np.cross(strike_vec, normal, axisa=0, axisb=1).T
    dip_vec /= np.linalg.norm(dip_vec, axis=0)[None, :]
    return strike_vec * length.T, dip_vec * length.T


def gradient_rows(grid: Grid, pts: np.ndarray, gpw: float = 1.0) -> list[Family]:
    """LI/_finite_difference_interpolator.py:342-401: two rows per point (strike.T, dip.T) = 0,
    weight = gpw (the per-point weight column is ignored)."""
    if pts.shape[0] == 0:
        return []
    pts = np.array(pts, dtype=float)
    idx, ins = cell_corners(grid, pts[:, :3])
    T = gradient_matrix(grid, local_coords(grid, pts[ins, :3]))
    norm = np.linalg.norm(pts[:, 3:6], axis=1)
    vec = pts[:, 3:6] / norm[:, None]
    T = T / norm[ins, None, None]
    strike, dip = strike_dip_vectors(vec[ins])
    out = []
    zeros = np.zeros(int(ins.sum()))
    for name, v in (("gradient", strike), ("gradient_0", dip)):
        A = np.einsum("ij,ijk->ik", v.T, T)
        out.append(make_family(name, A, zeros.copy(), idx[ins], gpw))
    return out


def normal_rows(grid: Grid, pts: np.ndarray, npw: float = 1.0) -> list[Family]:
    """LI/_finite_difference_interpolator.py:403-476: T[:,d,:] c = n_d, w = npw/3 * w_point."""
    if pts.shape[0] == 0:
        return []
    idx, ins = cell_corners(grid, pts[:, :3])
    T = gradient_matrix(grid, local_coords(grid, pts[ins, :3]))
    w = (npw / 3.0) * pts[ins, 6]
    return [make_family(f"norm_{d}", T[:, d, :], pts[ins, 3 + d], idx[ins], w) for d in range(3)]


def orthogonal_rows(grid: Grid, xyz: np.ndarray, vectors: np.ndarray, w: float = 1.0, b: float = 0.0,
                    name: str = "gradient orthogonal") -> Family | None:
    """LI/_finite_difference_interpolator.py:478-539 (tangent constraints go through this,
    LI/_discrete_interpolator.py:668-682)."""
    if xyz.shape[0] == 0:
        return None
    vectors = np.array(vectors, dtype=float)
    idx, ins = cell_corners(grid, xyz)
    norm = np.linalg.norm(vectors, axis=1)
    pos = norm > 0
    vectors[pos] /= norm[pos, None]
    T = gradient_matrix(grid, local_coords(grid, xyz[ins]))
    norm_in = norm[ins]
    T[norm_in > 0] /= norm_in[norm_in > 0, None, None]
    A = np.einsum("ij,ijk->ik", vectors[ins], T)
    return make_family(name, A, np.zeros(int(ins.sum())) + b, idx[ins], w)


def interface_pairs(n: int, mode: str = "auto", threshold: int = 200, fallback: str = "star") -> np.ndarray:
    """LI/_finite_difference_interpolator.py:271-295."""
    if n < 2:
        return np.zeros((0, 2), dtype=np.int64)
    if mode == "auto":
        mode = fallback if n > max(2, threshold) else "all"
    if mode == "all":
        ii, jj = np.triu_indices(n, k=1)
        return np.column_stack([ii, jj])
    if mode == "star":
        return np.column_stack([np.zeros(n - 1, dtype=np.int64), np.arange(1, n)])
    if mode == "chain":
        return np.column_stack([np.arange(0, n - 1), np.arange(1, n)])
    raise ValueError(mode)


def interface_rows(grid: Grid, pts: np.ndarray, ipw: float = 1.0) -> list[Family]:
    """LI/_finite_difference_interpolator.py:297-340: N(x_i) - N(x_j) = 0, 16 listed entries per row."""
    if pts.shape[0] < 2:
        return []
    idx, ins = cell_corners(grid, pts[:, :3])
    idc = idx[ins]
    A = trilinear(local_coords(grid, pts[ins, :3]))
    ids = pts[ins, 3]
    out = []
    for uid in np.unique(pts[np.logical_and(~np.isnan(pts[:, 3]), ins), 3]):
        m = ids == uid
        pair = interface_pairs(int(m.sum()))
        if pair.shape[0] == 0:
            continue
        Am, im = A[m], idc[m]
        rowsA = np.hstack([Am[pair[:, 0]], -Am[pair[:, 1]]])
        rowsI = np.hstack([im[pair[:, 0]], im[pair[:, 1]]])
        out.append(make_family(f"interface_{uid}", rowsA, np.zeros(rowsA.shape[0]), rowsI, ipw))
    return out


# --------------------------------------------------------------------------
# regularisation stencil rows
# --------------------------------------------------------------------------

# (dx,dy,dz) taps and raw mask coefficients of the six interior families,
# LI/_operator.py:29-35 with mask index [a,b,c] <-> (dz,dy,dx) = (a-1,b-1,c-1)
# (LI/_finite_difference_interpolator.py:541-634).
_S2 = 0.25 / np.sqrt(2.0)
STENCILS = {
    "dxx": ([(-1, 0, 0), (0, 0, 0), (1, 0, 0)], [1.0, -2.0, 1.0]),
    "dyy": ([(0, -1, 0), (0, 0, 0), (0, 1, 0)], [1.0, -2.0, 1.0]),
    "dzz": ([(0, 0, -1), (0, 0, 0), (0, 0, 1)], [1.0, -2.0, 1.0]),
    "dxy": ([(-1, -1, 0), (1, -1, 0), (-1, 1, 0), (1, 1, 0)], [-_S2, _S2, _S2, -_S2]),
    "dxz": ([(-1, 0, -1), (1, 0, -1), (-1, 0, 1), (1, 0, 1)], [-_S2, _S2, _S2, -_S2]),
    "dyz": ([(0, -1, -1), (0, 1, -1), (0, -1, 1), (0, 1, 1)], [-_S2, _S2, _S2, -_S2]),
}
# default FDI weights, LI/_finite_difference_interpolator.py:102-119
DEFAULT_WEIGHTS = dict(dxy=1.0, dyz=1.0, dxz=1.0, dxx=1.0, dyy=1.0, dzz=1.0, dx=1.0, dy=1.0, dz=1.0,
                       cpw=1.0, gpw=1.0, npw=10.0, tpw=1.0, ipw=1.0)


def effective_stencil_weights(weights: dict) -> dict:
    """Row weights actually applied: mixed terms are divided by 4 by get_operators,
    LC/supports/_3d_structured_grid.py:489-510; border families use dx/dy/dz,
    LI/_finite_difference_interpolator.py:1102-1139."""
    w = dict(DEFAULT_WEIGHTS)
    w.update(weights or {})
    return {
        "dxy": w["dxy"] / 4.0, "dyz": w["dyz"] / 4.0, "dxz": w["dxz"] / 4.0,
        "dxx": w["dxx"], "dyy": w["dyy"], "dzz": w["dzz"],
        "dx": w["dx"], "dy": w["dy"], "dz": w["dz"],
    }


def _node_ijk(grid: Grid, sel):
    """(i,j,k) of all nodes selected by the boolean function ``sel(i,j,k)``, in F-order of gi."""
    nx, ny, nz = (int(v) for v in grid.nsteps)
    gi = np.arange(grid.n_nodes)
    i = gi % nx
    j = gi // nx % ny
    k = gi // nx // ny
    m = sel(i, j, k)
    return i[m], j[m], k[m]


def stencil_family(grid: Grid, name: str, taps, coefs, w: float, sel) -> Family | None:
    """One regularisation family: a row per selected centre node, LI/_finite_difference_interpolator.py:643-692."""
    nx, ny, nz = (int(v) for v in grid.nsteps)
    i, j, k = _node_ijk(grid, sel)
    if i.shape[0] == 0:
        return None
    cols = np.stack([(i + dx) + nx * (j + dy) + nx * ny * (k + dz) for dx, dy, dz in taps], axis=1)
    A = np.tile(np.asarray(coefs, dtype=float), (i.shape[0], 1))
    return make_family(name, A, np.zeros(i.shape[0]), cols, w)


def regularisation_families(grid: Grid, weights: dict | None = None) -> list[Family]:
    """Interior families in get_operators order (dxy,dyz,dxz,dxx,dyy,dzz) on nodes interior on ALL axes
    (LC/supports/_3d_structured_grid.py:179-190) followed by the six border families on the planes
    index==1 / index==n-2 (LI/_finite_difference_interpolator.py:636-641,1052-1144)."""
    nx, ny, nz = (int(v) for v in grid.nsteps)
    ew = effective_stencil_weights(weights)

    def interior(i, j, k):
        return (i > 0) & (i < nx - 1) & (j > 0) & (j < ny - 1) & (k > 0) & (k < nz - 1)

    fams = []
    for name in ("dxy", "dyz", "dxz", "dxx", "dyy", "dzz"):
        taps, coefs = STENCILS[name]
        if ew[name] is None:
            continue
        f = stencil_family(grid, name, taps, coefs, ew[name], interior)
        if f is not None:
            fams.append(f)
    n = (nx, ny, nz)
    for axis, (key, sten) in enumerate((("dx", "dxx"), ("dy", "dyy"), ("dz", "dzz"))):
        if ew[key] == 0 or n[axis] < 3:
            continue
        taps, coefs = STENCILS[sten]
        for tag, plane in (("lower", 1), ("upper", n[axis] - 2)):
            sel = (lambda i, j, k, a=axis, p=plane: (i, j, k)[a] == p)
            f = stencil_family(grid, f"{key}_{tag}", taps, coefs, ew[key], sel)
            if f is not None:
                fams.append(f)
    return fams


def directional_families(grid: Grid, vectors: np.ndarray, w: float, name: str = "directional regularisation") -> list[Family]:
    """minimise_directional_gradient_change, LI/_finite_difference_interpolator.py:1207-1340 (StructuredGrid path): the six
    second-derivative masks at the nodes interior on all axes, row weights w * (vx^2, vy^2, vz^2, 2 vx vy, 2 vx vz, 2 vy vz)
    of the direction at the centre node, rows of weight 0 dropped (:1325-1329)."""
    nx, ny, nz = (int(v) for v in grid.nsteps)
    vectors = np.asarray(vectors, dtype=float)

    def interior(i, j, k):
        return (i > 0) & (i < nx - 1) & (j > 0) & (j < ny - 1) & (k > 0) & (k < nz - 1)

    i, j, k = _node_ijk(grid, interior)
    centre = i + nx * j + nx * ny * k
    vx, vy, vz = vectors[centre, 0], vectors[centre, 1], vectors[centre, 2]
    comp = {"dxx": vx ** 2, "dyy": vy ** 2, "dzz": vz ** 2, "dxy": 2.0 * vx * vy, "dxz": 2.0 * vx * vz, "dyz": 2.0 * vy * vz}
    fams = []
    for key in ("dxx", "dyy", "dzz", "dxy", "dxz", "dyz"):
        taps, coefs = STENCILS[key]
        full = stencil_family(grid, f"{name}_{key}", taps, coefs, 1.0, interior)
        if full is None:
            continue
        row_w = w * comp[key]
        with np.errstate(invalid="ignore"):
            nz_ = np.abs(row_w) > 0.0
        if not np.any(nz_):
            continue
        fams.append(Family(full.name, full.A[nz_], full.idx[nz_], full.b[nz_], row_w[nz_]))
    return fams


# --------------------------------------------------------------------------
# system assembly, preprocessing, CG
# --------------------------------------------------------------------------


def assemble(families: list[Family], dof: int):
    """A = vstack(diag(w_i) M_i), B = hstack(w_i b_i); LI/_discrete_interpolator.py:684-705."""
    mats, bs = [], []
    for f in families:
        if f is None or f.w.shape[0] == 0:
            continue
        mats.append(f.matrix(dof).multiply(f.w[:, None]))
        bs.append(f.b * f.w)
    if not mats:
        return sparse.csr_matrix((0, dof)), np.zeros(0)
    return sparse.vstack(mats).tocsr(), np.hstack(bs)


def preprocess(A, b, ridge_factor: float = 1e-8, add_ridge: bool = True, apply_scaling: bool = True):
    """Ridge rows + column scaling; LI/_solver_pipeline.py:86-124, LI/_discrete_interpolator.py:815-833."""
    dof = A.shape[1]
    if add_ridge:
        A = sparse.vstack([A, sparse.eye(dof) * ridge_factor]).tocsr()
        b = np.hstack([b, np.zeros(dof)])
    S = None
    if apply_scaling:
        col = np.sqrt(np.asarray(A.multiply(A).sum(axis=0)).ravel())
        s = np.ones(dof)
        s[col > 0] = 1.0 / col[col > 0]
        S = s
        A = (A @ sparse.diags(s)).tocsr()
    return A, b, S


@dataclass
class CGResult:
    x: np.ndarray
    info: int
    iterations: int
    residual_norm: float
    threshold: float
    history: list = field(default_factory=list)


def cg(matvec, b: np.ndarray, rtol: float = 1e-5, atol: float = 0.0, maxiter: int | None = None,
       x0: np.ndarray | None = None, keep_history: bool = False) -> CGResult:
    """scipy.sparse.linalg.cg (scipy 1.18, _isolve/iterative.py) restated, M = identity:
    threshold = max(atol, rtol*||b||); test ||r|| < threshold at the TOP of every iteration;
    rho = r.r ; p = r + (rho/rho_prev) p ; q = A p ; alpha = rho / p.q ; x += alpha p ; r -= alpha q.
    Exhausting maxiter (default 10 n) returns info = maxiter."""
    b = np.asarray(b, dtype=float)
    n = b.shape[0]
    bn = float(np.linalg.norm(b))
    thr = max(float(atol), float(rtol) * bn)
    if bn == 0:
        return CGResult(b.copy(), 0, 0, 0.0, thr)
    if maxiter is None:
        maxiter = 10 * n
    x = np.zeros(n) if x0 is None else np.array(x0, dtype=float)
    r = b - matvec(x) if x.any() else b.copy()
    p = None
    rho_prev = None
    hist = []
    it = 0
    for it in range(maxiter):
        rn = float(np.linalg.norm(r))
        if keep_history:
            hist.append(rn)
        if rn < thr:
            return CGResult(x, 0, it, rn, thr, hist)
        rho = float(np.dot(r, r))
        if it > 0:
            p *= rho / rho_prev
            p += r
        else:
            p = r.copy()
        q = matvec(p)
        alpha = rho / float(np.dot(p, q))
        x += alpha * p
        r -= alpha * q
        rho_prev = rho
    return CGResult(x, maxiter, maxiter, float(np.linalg.norm(r)), thr, hist)


def solve_reference_path(families: list[Family], dof: int, rtol: float = 1e-5, atol: float = 0.0,
                         ridge_factor: float = 1e-8, apply_scaling: bool = True, maxiter=None,
                         keep_history: bool = False, timing: dict | None = None):
    """The reference's default solve: explicit A, ridge, column scaling, N = A^T A as CSR, CG, c = S y.
    LI/_discrete_interpolator.py:1221-1319, LI/_solver_strategy.py:48-71."""
    from time import perf_counter

    t0 = perf_counter()
    A, b = assemble(families, dof)
    t1 = perf_counter()
    A, b, S = preprocess(A, b, ridge_factor=ridge_factor, apply_scaling=apply_scaling)
    t2 = perf_counter()
    N = (A.T @ A).tocsr()
    rhs = A.T @ b
    t3 = perf_counter()
    res = cg(lambda v: N @ v, rhs, rtol=rtol, atol=atol, maxiter=maxiter, keep_history=keep_history)
    t4 = perf_counter()
    c = res.x if S is None else S * res.x
    if timing is not None:
        timing.update(assembly_seconds=t1 - t0, preprocess_seconds=t2 - t1, gram_seconds=t3 - t2,
                      solve_seconds=t4 - t3, matrix_rows=int(A.shape[0]), matrix_nnz=int(A.nnz),
                      gram_nnz=int(N.nnz), cg_iterations=res.iterations)
    return c, res, S


# --------------------------------------------------------------------------
# matrix-free restatement of the normal operator (used for grids where the
# explicit Gram matrix is too slow to build in a test)
# --------------------------------------------------------------------------


def regularisation_normal_matvec(grid: Grid, weights: dict | None, p: np.ndarray) -> np.ndarray:
    """q = R^T R p for the 12 regularisation families, two-pass with row masks
    (the exact form the reference's matrix-free blocks compute,
    LI/_finite_difference_interpolator.py:760-850), done with array slicing."""
    nx, ny, nz = (int(v) for v in grid.nsteps)
    ew = effective_stencil_weights(weights)
    P = p.reshape((nx, ny, nz), order="F")
    Q = np.zeros_like(P)
    n = (nx, ny, nz)

    def sl(lo, hi, d):
        # slice of centres [lo,hi) shifted by d along one axis
        return slice(lo + d, hi + d)

    # interior families: centres 1..n-2 on every axis
    ctr = [(1, m - 1) for m in n]
    if all(hi > lo for lo, hi in ctr):
        for name in ("dxy", "dyz", "dxz", "dxx", "dyy", "dzz"):
            taps, coefs = STENCILS[name]
            coefs = np.asarray(coefs) / np.linalg.norm(coefs)
            w2 = ew[name] ** 2
            if w2 == 0:
                continue
            y = np.zeros([hi - lo for lo, hi in ctr])
            views = []
            for (dx, dy, dz), c in zip(taps, coefs):
                v = (sl(*ctr[0], dx), sl(*ctr[1], dy), sl(*ctr[2], dz))
                views.append((v, c))
                y += c * P[v]
            y *= w2
            for v, c in views:
                Q[v] += c * y
    # border families
    c3 = np.array([1.0, -2.0, 1.0]) / np.sqrt(6.0)
    for axis, key in enumerate(("dx", "dy", "dz")):
        w2 = ew[key] ** 2
        if w2 == 0 or n[axis] < 3:
            continue
        for plane in (1, n[axis] - 2):
            idx = [slice(None)] * 3
            y = 0.0
            views = []
            for d, c in zip((-1, 0, 1), c3):
                v = list(idx)
                v[axis] = plane + d
                views.append((tuple(v), c))
                y = y + c * P[tuple(v)]
            y = w2 * y
            for v, c in views:
                Q[v] += c * y
    return Q.ravel(order="F")


def regularisation_normal_diagonal(grid: Grid, weights: dict | None) -> np.ndarray:
    """diag(R^T R): sum over rows of (w coef)^2 landing on each node."""
    nx, ny, nz = (int(v) for v in grid.nsteps)
    ew = effective_stencil_weights(weights)
    D = np.zeros((nx, ny, nz))
    n = (nx, ny, nz)
    ctr = [(1, m - 1) for m in n]
    if all(hi > lo for lo, hi in ctr):
        for name in ("dxy", "dyz", "dxz", "dxx", "dyy", "dzz"):
            taps, coefs = STENCILS[name]
            coefs = np.asarray(coefs) / np.linalg.norm(coefs)
            w2 = ew[name] ** 2
            for (dx, dy, dz), c in zip(taps, coefs):
                D[slice(ctr[0][0] + dx, ctr[0][1] + dx), slice(ctr[1][0] + dy, ctr[1][1] + dy),
                  slice(ctr[2][0] + dz, ctr[2][1] + dz)] += w2 * c * c
    c3 = np.array([1.0, -2.0, 1.0]) / np.sqrt(6.0)
    for axis, key in enumerate(("dx", "dy", "dz")):
        w2 = ew[key] ** 2
        if w2 == 0 or n[axis] < 3:
            continue
        for plane in (1, n[axis] - 2):
            for d, c in zip((-1, 0, 1), c3):
                v = [slice(None)] * 3
                v[axis] = plane + d
                D[tuple(v)] += w2 * c * c
    return D.ravel(order="F")


class MatrixFreeSystem:
    """N = A_d^T W A_d + R^T R + ridge^2 I applied matrix-free, optionally column scaled
    (S N S) exactly as the explicit path scales it (LI/_solver_pipeline.py:99-122)."""

    def __init__(self, grid: Grid, data_families: list[Family], weights: dict | None = None,
                 ridge_factor: float = 1e-8, apply_scaling: bool = True):
        self.grid = grid
        self.weights = weights
        self.ridge2 = ridge_factor ** 2
        dof = grid.n_nodes
        Ad, bd = assemble([f for f in data_families if f is not None], dof)
        self.Ad = Ad.tocsr()
        self.AdT = self.Ad.T.tocsr()
        self.rhs_unscaled = self.AdT @ bd
        diag = np.asarray(self.Ad.multiply(self.Ad).sum(axis=0)).ravel()
        diag = diag + regularisation_normal_diagonal(grid, weights) + self.ridge2
        self.diag = diag
        if apply_scaling:
            s = np.ones(dof)
            nz = diag > 0
            s[nz] = 1.0 / np.sqrt(diag[nz])
            self.S = s
        else:
            self.S = None

    @property
    def rhs(self):
        return self.rhs_unscaled if self.S is None else self.S * self.rhs_unscaled

    def matvec(self, y: np.ndarray) -> np.ndarray:
        p = y if self.S is None else self.S * y
        q = self.AdT @ (self.Ad @ p) + regularisation_normal_matvec(self.grid, self.weights, p) + self.ridge2 * p
        return q if self.S is None else self.S * q

    def solve(self, rtol=1e-5, atol=0.0, maxiter=None, keep_history=False):
        res = cg(self.matvec, self.rhs, rtol=rtol, atol=atol, maxiter=maxiter, keep_history=keep_history)
        c = res.x if self.S is None else self.S * res.x
        return c, res


# --------------------------------------------------------------------------
# evaluation
# --------------------------------------------------------------------------


def evaluate_value(grid: Grid, pts: np.ndarray, c: np.ndarray) -> np.ndarray:
    """LC/supports/_3d_structured_grid.py:295-322 + LI/_discrete_interpolator.py:1342-1362:
    NaN outside the grid, 0.0 for query rows holding NaN coordinates."""
    pts = np.asarray(pts, dtype=float)
    out = np.zeros(pts.shape[0])
    bad = np.any(np.isnan(pts), axis=1)
    good = pts[~bad]
    if good.shape[0]:
        idx, ins = cell_corners(grid, good)
        v = np.full(good.shape[0], np.nan)
        w = trilinear(local_coords(grid, good[ins]))
        v[ins] = np.sum(w * c[idx[ins]], axis=1)
        out[~bad] = v
    return out


def evaluate_gradient(grid: Grid, pts: np.ndarray, c: np.ndarray) -> np.ndarray:
    """LC/supports/_3d_structured_grid.py:324-391: sum_i T[d,i] c[idx_i]; rows outside the grid are 0
    (T is zero-initialised, :356-357); rows with NaN coordinates are outside (comparisons false)."""
    pts = np.asarray(pts, dtype=float)
    out = np.zeros((pts.shape[0], 3))
    if pts.shape[0] == 0:
        return out
    idx, ins = cell_corners(grid, pts)
    T = gradient_matrix(grid, local_coords(grid, pts[ins]))
    out[ins] = np.einsum("ijk,ik->ij", T, c[idx[ins]])
    return out


# --------------------------------------------------------------------------
# synthetic workload of SURVEY.md section 8(d) / BASELINE.md section 3
# --------------------------------------------------------------------------


def synthetic_field(p: np.ndarray) -> np.ndarray:
    return p[:, 2] + 0.1 * np.sin(2 * np.pi * p[:, 0]) + 0.05 * np.cos(2 * np.pi * p[:, 1])


def synthetic_gradient(p: np.ndarray) -> np.ndarray:
    return np.stack([0.2 * np.pi * np.cos(2 * np.pi * p[:, 0]),
                     -0.1 * np.pi * np.sin(2 * np.pi * p[:, 1]),
                     np.ones(p.shape[0])], axis=1)


def synthetic_problem(n: int, n_value: int, n_gradient: int, seed: int = 0):
    """Unit-cube grid with n nodes per axis, value rows [x,y,z,f,1], gradient rows [x,y,z,grad f,1]."""
    grid = Grid.from_cells([n - 1] * 3, origin=np.zeros(3), step=np.ones(3) / (n - 1))
    rng = np.random.default_rng(seed)
    pv = rng.uniform(0, 1, (n_value, 3))
    pg = rng.uniform(0, 1, (n_gradient, 3))
    value = np.column_stack([pv, synthetic_field(pv), np.ones(n_value)])
    gradient = np.column_stack([pg, synthetic_gradient(pg), np.ones(n_gradient)])
    return grid, value, gradient
