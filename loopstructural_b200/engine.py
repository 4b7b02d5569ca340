"""Thin Python owner of one ``lsb200`` device handle.

``Engine`` is the host-side object the interpolator classes (``interpolators.py``) and the
LoopStructural plug-in (``plugin.py``) drive.  It holds no numerics of its own: every method is one
call through the C-ABI of ``include/lsb200.h`` into the CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
import sys
from dataclasses import dataclass

import numpy as np

from . import _native as nat

ROWS_VALUE, ROWS_GRADIENT, ROWS_NORMAL, ROWS_TANGENT, ROWS_ELEMENT_VECTOR = 0, 1, 2, 3, 4
PART_REGULARISATION, PART_ROWS, PART_RIDGE = 1, 2, 4

# FDI weight defaults, LI/_finite_difference_interpolator.py:102-119
FDI_DEFAULT_WEIGHTS = dict(dxy=1.0, dyz=1.0, dxz=1.0, dxx=1.0, dyy=1.0, dzz=1.0, dx=1.0, dy=1.0, dz=1.0,
                           cpw=1.0, gpw=1.0, npw=10.0, tpw=1.0, ipw=1.0)


def effective_stencil_weights(weights: dict) -> np.ndarray:
    """Row weights the twelve stencil families actually get, in LSB200_W_* order.

    get_operators divides the mixed-derivative weights by 4
    (LC/supports/_3d_structured_grid.py:489-510); border families use dx/dy/dz
    (LI/_finite_difference_interpolator.py:1102-1139)."""
    w = dict(FDI_DEFAULT_WEIGHTS)
    w.update({k: v for k, v in (weights or {}).items() if v is not None})
    return np.array([w["dxx"], w["dyy"], w["dzz"], w["dxy"] / 4.0, w["dxz"] / 4.0, w["dyz"] / 4.0,
                     w["dx"], w["dy"], w["dz"]], dtype=np.float64)


def interface_pair_indices(n: int, mode: str = "auto", threshold: int = 200, fallback: str = "star") -> np.ndarray:
    """FiniteDifferenceInterpolator._interface_pair_indices (LI/_finite_difference_interpolator.py:271-295)."""
    if n < 2:
        return np.zeros((0, 2), dtype=np.int64)
    mode = str(mode).lower()
    threshold = max(2, int(threshold))
    fallback = str(fallback).lower()
    if fallback not in ("star", "chain"):
        fallback = "star"
    if mode == "auto":
        mode = fallback if n > threshold else "all"
    if mode == "all":
        ii, jj = np.triu_indices(n, k=1)
        return np.column_stack([ii, jj]).astype(np.int64)
    if mode == "star":
        return np.column_stack([np.zeros(n - 1, dtype=np.int64), np.arange(1, n, dtype=np.int64)])
    if mode == "chain":
        return np.column_stack([np.arange(0, n - 1, dtype=np.int64), np.arange(1, n, dtype=np.int64)])
    raise ValueError(f"Unknown interface_pair_mode {mode!r}. Use one of: auto, all, star, chain")


@dataclass
class SolveResult:
    info: int
    iterations: int
    residual_norm: float
    threshold: float
    rhs_norm: float
    setup_seconds: float
    solve_seconds: float
    kernel_launches: int


class _PinnedBlock:
    """Owner of one cudaHostAlloc block (lsb200_host_alloc / lsb200_host_free)."""

    def __init__(self, lib, nbytes: int):
        self.lib = lib
        self.ptr = C.c_void_p()
        nat.check(lib.lsb200_host_alloc(int(nbytes), C.byref(self.ptr)))

    def __del__(self):
        try:
            if self.ptr:
                self.lib.lsb200_host_free(self.ptr)
        except Exception:
            pass


class Engine:
    """One support (regular grid or tet mesh on the same lattice) on one GPU.

    ``nsteps`` are NODE counts per axis (LC/supports/_3d_base_structured.py:66); ``slab=(k0,k1)``
    restricts the handle to the z-planes [k0,k1) of a multi-GPU solve."""

    def __init__(self, nsteps, origin, step, kind: str = "grid", device: int = 0, slab=None):
        self._lib = nat.load()
        self._h = nat.Handle()
        self.nsteps = np.asarray(nsteps, dtype=np.int64).reshape(3).copy()
        self.origin = nat.as_f64(origin, (3,)).copy()
        self.step = nat.as_f64(step, (3,)).copy()
        self.kind = kind
        self.device = int(device)
        k0, k1 = (0, int(self.nsteps[2])) if slab is None else (int(slab[0]), int(slab[1]))
        self.slab = (k0, k1)
        if kind == "grid":
            rc = self._lib.lsb200_create_grid(nat.iptr(self.nsteps), nat.dptr(self.origin), nat.dptr(self.step),
                                              k0, k1, self.device, C.byref(self._h))
        elif kind == "tetmesh":
            if slab is None:
                rc = self._lib.lsb200_create_tetmesh(nat.iptr(self.nsteps), nat.dptr(self.origin), nat.dptr(self.step),
                                                     self.device, C.byref(self._h))
            else:
                rc = self._lib.lsb200_create_tetmesh_slab(nat.iptr(self.nsteps), nat.dptr(self.origin), nat.dptr(self.step),
                                                          k0, k1, self.device, C.byref(self._h))
        else:
            raise ValueError(f"unknown support kind {kind!r}")
        nat.check(rc)

    # -- lifetime --------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.lsb200_destroy(self._h)
            self._h = nat.Handle()
        self._pinned_buf = None  # the block itself lives as long as any array over it (see pinned_vector)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n_nodes(self) -> int:
        return int(np.prod(self.nsteps))

    @property
    def local_dof(self) -> int:
        return int(self._lib.lsb200_local_dof(self._h))

    # -- system ----------------------------------------------------------------------------
    def set_regularisation(self, weights: dict | None = None):
        ew = effective_stencil_weights(weights or {})
        nat.check(self._lib.lsb200_set_regularisation(self._h, nat.dptr(ew)))

    def set_p1_regularisation(self, w: float):
        """Face gradient-jump regularisation of weight ``w`` (cgw).  The first call assembles the stencil table of
        this support on a mesh of at most 6 cells per axis (p1_tables.py) and uploads it."""
        if not getattr(self, "_p1_tables_set", False):
            from . import p1_tables

            m, table = p1_tables.regularisation_table(self.nsteps - 1, self.step)
            m = np.ascontiguousarray(m, dtype=np.int64)
            nat.check(self._lib.lsb200_set_p1_tables(self._h, nat.iptr(m), nat.dptr(table)))
            self._p1_tables_set = True
        nat.check(self._lib.lsb200_set_p1_regularisation(self._h, float(w)))

    def set_options(self, ridge_factor: float = 1e-8, add_ridge: bool = True, apply_scaling: bool = True,
                    deterministic: bool = False):
        nat.check(self._lib.lsb200_set_options(self._h, float(ridge_factor), int(add_ridge), int(apply_scaling),
                                               int(deterministic)))

    def add_rows(self, A, idx, b, w, normalise: bool = True):
        """add_constraints_to_least_squares (LI/_discrete_interpolator.py:316-389): rows (n,k),
        global column ids (n,k), targets (n,), scalar or (n,) weights."""
        A = nat.as_f64(A)
        if A.ndim != 2:
            raise ValueError("A must be (n_rows, k)")
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        if idx.shape != A.shape:
            raise ValueError(f"idx shape {idx.shape} does not match A {A.shape}")
        n = A.shape[0]
        b = nat.as_f64(np.broadcast_to(np.asarray(b, dtype=np.float64), (n,)))
        w = nat.as_f64(np.broadcast_to(np.asarray(w, dtype=np.float64), (n,)))
        nat.check(self._lib.lsb200_add_rows(self._h, n, A.shape[1], nat.dptr(A), nat.iptr(idx), nat.dptr(b),
                                            nat.dptr(w), int(normalise)))

    def add_point_rows(self, kind: int, data, weight: float) -> int:
        data = nat.as_f64(data)
        if data.ndim != 2:
            raise ValueError("data must be (n, ncols)")
        out = C.c_int64(0)
        nat.check(self._lib.lsb200_add_point_rows(self._h, int(kind), data.shape[0], nat.dptr(data), data.shape[1],
                                                  float(weight), C.byref(out)))
        return int(out.value)

    def add_interface_rows(self, points, weight: float, mode: str = "auto", threshold: int = 200,
                           fallback: str = "star") -> int:
        """FDI.add_interface_constraints (LI/_finite_difference_interpolator.py:297-340).  ``points`` is (n,5):
        x y z id w.  The pairing per interface id is host logic here as it is in the reference
        (_interface_pair_indices, :271-295, over the points that lie inside the grid); the rows themselves are
        generated on the device."""
        pts = nat.as_f64(points)
        if pts.ndim != 2 or pts.shape[1] < 4:
            raise ValueError("interface points must be (n, >=4): x y z id")
        if pts.shape[0] < 2:
            return 0
        # StructuredGrid.inside with its 32 eps tolerance (LC/supports/_3d_base_structured.py:287-293)
        maximum = self.origin + (self.nsteps - 1) * self.step
        tol = np.finfo(float).eps * 32.0 * np.maximum(np.abs(maximum - self.origin), 1.0)
        xyz = pts[:, :3]
        with np.errstate(invalid="ignore"):
            inside = np.all((xyz >= self.origin - tol) & (xyz <= maximum + tol), axis=1)
        ids = pts[inside, 3]
        xin = xyz[inside]
        pi, pj = [], []
        for uid in np.unique(ids[~np.isnan(ids)]):
            sel = np.flatnonzero(ids == uid)
            pair = interface_pair_indices(len(sel), mode, threshold, fallback)
            if pair.shape[0]:
                pi.append(xin[sel[pair[:, 0]]])
                pj.append(xin[sel[pair[:, 1]]])
        if not pi:
            return 0
        xi, xj = nat.as_f64(np.vstack(pi)), nat.as_f64(np.vstack(pj))
        out = C.c_int64(0)
        nat.check(self._lib.lsb200_add_pair_rows(self._h, xi.shape[0], nat.dptr(xi), nat.dptr(xj), float(weight),
                                                 C.byref(out)))
        return int(out.value)

    def add_directional_rows(self, vectors, weight: float) -> int:
        """FDI.minimise_directional_gradient_change (LI/_finite_difference_interpolator.py:1207-1340): ``vectors`` is the
        (n_nodes, 3) direction field in the support's node order."""
        v = nat.as_f64(vectors)
        if v.shape != (int(np.prod(self.nsteps)), 3):
            raise ValueError(f"vectors must be ({int(np.prod(self.nsteps))}, 3), got {v.shape}")
        out = C.c_int64(0)
        nat.check(self._lib.lsb200_add_directional_rows(self._h, nat.dptr(v), float(weight), C.byref(out)))
        return int(out.value)

    def add_face_vector_rows(self, nodes_a, nodes_b, vectors, weight: float) -> int:
        """P1Interpolator.minimise_edge_jumps with a vector per shared face (LI/_p1interpolator.py:115-166): one row
        v . (grad phi_a - grad phi_b) = 0 per pair of face-adjacent tetrahedra, given by their (n,4) global node ids."""
        a = np.ascontiguousarray(nodes_a, dtype=np.int64)
        b = np.ascontiguousarray(nodes_b, dtype=np.int64)
        v = nat.as_f64(vectors)
        if a.ndim != 2 or a.shape[1] != 4 or b.shape != a.shape or v.shape != (a.shape[0], 3):
            raise ValueError("nodes_a, nodes_b must be (n,4) and vectors (n,3)")
        out = C.c_int64(0)
        nat.check(self._lib.lsb200_add_face_vector_rows(self._h, a.shape[0], a.ctypes.data_as(C.POINTER(C.c_int64)),
                                                        b.ctypes.data_as(C.POINTER(C.c_int64)), nat.dptr(v), float(weight),
                                                        C.byref(out)))
        return int(out.value)

    def add_field_orthogonal_rows(self, other: "Engine", step: int = 10, weight: float = 1.0, b: float = 0.0,
                                  other_world_to_local=None) -> int:
        """GeologicalFeatureBuilder.install_gradient_constraint on the device
        (LoopStructural/modelling/features/builders/_geological_feature_builder.py:293-318): rows that make this field's
        gradient orthogonal to the gradient of ``other`` (a solved grid engine) at every ``step``-th cell barycentre."""
        tr = None
        if other_world_to_local is not None:
            from .fault import transform_rows

            tr = np.ascontiguousarray(transform_rows(other_world_to_local))
        out = C.c_int64(0)
        nat.check(self._lib.lsb200_add_field_orthogonal_rows(self._h, other._h, None if tr is None else nat.dptr(tr),
                                                             int(step), float(weight), float(b), C.byref(out)))
        return int(out.value)

    def clear_rows(self):
        nat.check(self._lib.lsb200_clear_rows(self._h))

    @property
    def num_rows(self) -> int:
        return int(self._lib.lsb200_num_rows(self._h))

    def finalize(self):
        nat.check(self._lib.lsb200_finalize_system(self._h))

    def comm_init(self, nccl_id: bytes, rank: int, nranks: int):
        buf = (C.c_ubyte * 128).from_buffer_copy(nccl_id)
        nat.check(self._lib.lsb200_comm_init(self._h, buf, int(rank), int(nranks)))

    @property
    def uses_peer_memory(self) -> bool:
        """True when the scalar reductions of CG go through peer-mapped mailboxes instead of ncclAllReduce."""
        return bool(self._lib.lsb200_comm_uses_peer_memory(self._h))

    def residual_norm(self):
        """(sqrt(r . S^2 r), ||r||_2) of the current solution with r = A^T W b - N c recomputed through the operator;
        collective on slab handles."""
        a, b = C.c_double(0.0), C.c_double(0.0)
        nat.check(self._lib.lsb200_residual_norm(self._h, C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_ubyte * 128)()
        nat.check(nat.load().lsb200_comm_unique_id(buf))
        return bytes(buf)

    def _vec(self):
        return np.empty(self.local_dof, dtype=np.float64)

    def apply_normal(self, p, parts: int = PART_REGULARISATION | PART_ROWS | PART_RIDGE) -> np.ndarray:
        p = nat.as_f64(p, (self.local_dof,))
        q = self._vec()
        nat.check(self._lib.lsb200_apply_normal(self._h, nat.dptr(p), nat.dptr(q), int(parts)))
        return q

    def rhs(self) -> np.ndarray:
        out = self._vec()
        nat.check(self._lib.lsb200_get_rhs(self._h, nat.dptr(out)))
        return out

    def set_rhs(self, rhs):
        """Replace the right-hand side of the normal equations (valid until the rows change)."""
        rhs = nat.as_f64(rhs, (self.local_dof,))
        nat.check(self._lib.lsb200_set_rhs(self._h, nat.dptr(rhs)))

    def diagonal(self) -> np.ndarray:
        out = self._vec()
        nat.check(self._lib.lsb200_get_diagonal(self._h, nat.dptr(out)))
        return out

    def scaling(self) -> np.ndarray:
        out = self._vec()
        nat.check(self._lib.lsb200_get_scaling(self._h, nat.dptr(out)))
        return out

    def pinned_vector(self) -> np.ndarray:
        """A page-locked float64 array of local_dof entries; pass it as ``out=`` to solve_cg / get_c to take the result
        without page faults or a pageable bounce.  The block belongs to the ARRAY (array -> ctypes buffer -> _PinnedBlock),
        not to the engine: it is freed when the last array or view over it dies, so a caller may keep the result past
        ``close()``.  The engine reuses its block only while nobody else holds an array over it; otherwise the next call
        hands out a fresh one and the caller's data stay untouched."""
        n = self.local_dof
        buf = getattr(self, "_pinned_buf", None)
        # references to buf: this attribute, the local name, getrefcount's argument; every live array adds one (its base)
        if buf is None or len(buf) != n or sys.getrefcount(buf) > 3:
            owner = _PinnedBlock(self._lib, 8 * n)
            buf = (C.c_double * n).from_address(owner.ptr.value)
            buf._owner = owner
            self._pinned_buf = buf
        return np.frombuffer(buf, dtype=np.float64)

    def solve_cg(self, rtol: float = 1e-5, atol: float = 0.0, maxiter: int | None = None, x0=None,
                 want_c: bool = True, out: np.ndarray | None = None):
        info = nat.SolveInfo()
        x0p = None
        if x0 is not None:
            x0 = nat.as_f64(x0, (self.local_dof,))
            x0p = nat.dptr(x0)
        if out is not None:
            if out.dtype != np.float64 or out.shape != (self.local_dof,) or not out.flags.c_contiguous:
                raise ValueError("out must be a contiguous float64 array of local_dof entries")
            want_c = True
        c = (out if out is not None else self._vec()) if want_c else None
        nat.check(self._lib.lsb200_solve_cg(self._h, float(rtol), float(atol), int(maxiter or 0), x0p,
                                            nat.dptr(c) if want_c else None, C.byref(info)))
        res = SolveResult(int(info.info), int(info.iterations), float(info.residual_norm), float(info.threshold),
                          float(info.rhs_norm), float(info.setup_seconds), float(info.solve_seconds),
                          int(info.kernel_launches))
        return c, res

    # -- solution and evaluation -----------------------------------------------------------
    def get_c(self) -> np.ndarray:
        out = self._vec()
        nat.check(self._lib.lsb200_get_c(self._h, nat.dptr(out)))
        return out

    def set_c(self, c):
        c = nat.as_f64(c, (self.local_dof,))
        nat.check(self._lib.lsb200_set_c(self._h, nat.dptr(c)))

    def get_c_device(self, out_ptr: int, stream: int = 0):
        """Owned planes of the solution into a device buffer of local_dof doubles (asynchronous on ``stream``)."""
        nat.check(self._lib.lsb200_get_c_dev(self._h, C.c_void_p(out_ptr), C.c_void_p(stream)))

    def set_c_device(self, c_ptr: int, stream: int = 0):
        nat.check(self._lib.lsb200_set_c_dev(self._h, C.c_void_p(c_ptr), C.c_void_p(stream)))

    def evaluate_value(self, xyz) -> np.ndarray:
        xyz = nat.as_f64(xyz)
        if xyz.ndim != 2 or xyz.shape[1] != 3:
            raise ValueError("xyz must be (n,3)")
        out = np.empty(xyz.shape[0], dtype=np.float64)
        nat.check(self._lib.lsb200_eval_value(self._h, xyz.shape[0], nat.dptr(xyz), nat.dptr(out)))
        return out

    def evaluate_gradient(self, xyz) -> np.ndarray:
        xyz = nat.as_f64(xyz)
        if xyz.ndim != 2 or xyz.shape[1] != 3:
            raise ValueError("xyz must be (n,3)")
        out = np.empty((xyz.shape[0], 3), dtype=np.float64)
        nat.check(self._lib.lsb200_eval_gradient(self._h, xyz.shape[0], nat.dptr(xyz), nat.dptr(out)))
        return out

    def evaluate_value_device(self, xyz_ptr: int, out_ptr: int, n: int, stream: int = 0):
        """Device pointers (e.g. ``tensor.data_ptr()``), asynchronous on ``stream``."""
        nat.check(self._lib.lsb200_eval_value_dev(self._h, int(n), C.c_void_p(xyz_ptr), C.c_void_p(out_ptr),
                                                  C.c_void_p(stream)))

    def evaluate_gradient_device(self, xyz_ptr: int, out_ptr: int, n: int, stream: int = 0):
        nat.check(self._lib.lsb200_eval_gradient_dev(self._h, int(n), C.c_void_p(xyz_ptr), C.c_void_p(out_ptr),
                                                     C.c_void_p(stream)))

    def evaluate_lattice(self, start, stop, n, order: str = "F", gradient: bool = False, world_to_local=None,
                         out_ptr: int | None = None, stream: int = 0):
        """Value (N) or world gradient (N,3) of the field on np.linspace(start, stop, n) per axis, flattened in
        ``order`` -- the points of BoundingBox.regular_grid(local=False) (LC/geometry/_bounding_box.py:532-573) --
        without building or copying the coordinates.  ``out_ptr``: device pointer (asynchronous on ``stream``)."""
        start, stop = nat.as_f64(start, (3,)), nat.as_f64(stop, (3,))
        n = np.ascontiguousarray(n, dtype=np.int64).reshape(3)
        if order not in ("F", "C"):
            raise ValueError("order must be 'F' or 'C'")
        tr = None
        if world_to_local is not None:
            from .fault import transform_rows

            tr = np.ascontiguousarray(transform_rows(world_to_local))
        trp = None if tr is None else nat.dptr(tr)
        if out_ptr is not None:
            nat.check(self._lib.lsb200_eval_lattice_dev(self._h, trp, nat.dptr(start), nat.dptr(stop), nat.iptr(n),
                                                        int(order == "F"), int(gradient), C.c_void_p(out_ptr), C.c_void_p(stream)))
            return None
        total = int(np.prod(n))
        out = np.empty((total, 3) if gradient else total, dtype=np.float64)
        nat.check(self._lib.lsb200_eval_lattice(self._h, trp, nat.dptr(start), nat.dptr(stop), nat.iptr(n), int(order == "F"),
                                                int(gradient), nat.dptr(out)))
        return out

    def describe_plan(self) -> str:
        """Launch plan of the regularisation operator (after finalize(): the one chosen by timing the candidates)."""
        buf = C.create_string_buffer(512)
        nat.check(self._lib.lsb200_describe_plan(self._h, buf, 512))
        return buf.value.decode()

    def small_phases(self, iters: int = 64) -> dict:
        """Mean microseconds per iteration of the persistent CG loop by phase, as CTA 0 sees them (diagnostics)."""
        out = np.zeros(8)
        nat.check(self._lib.lsb200_debug_small_phases(self._h, int(iters), nat.dptr(out)))
        names = ("p_update", "barrier_p", "operator", "barrier_q", "rows", "barrier_pq", "xr_update", "barrier_rho")
        return {k: float(v) for k, v in zip(names, out)}

    def cta_trace(self, max_ctas: int = 1 << 18) -> np.ndarray:
        """(n_ctas, 3) uint64: start ns, end ns, SM id of every CTA of one launch of the regularisation operator."""
        out = np.zeros((max_ctas, 3), dtype=np.uint64)
        n = C.c_int64(0)
        nat.check(self._lib.lsb200_debug_cta_trace(self._h, max_ctas, out.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(n)))
        return out[: n.value]

    def time_kernel(self, which: int, reps: int = 20) -> float:
        out = C.c_double(0.0)
        nat.check(self._lib.lsb200_time_kernel(self._h, int(which), int(reps), C.byref(out)))
        return float(out.value)
