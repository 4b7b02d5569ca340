"""Puts the B200 engine behind LoopStructural's own interpolator classes.

``install()`` needs the reference packages (``loop_interpolation`` / ``loop_common``) importable.  It replaces
``interpolator_map[InterpolatorType.FINITE_DIFFERENCE]`` -- the plain module-level dict that
``InterpolatorFactory.create_interpolator`` consults at call time (LI/__init__.py:106-117,
LI/_interpolator_factory.py:57-66) -- by a subclass of ``FiniteDifferenceInterpolator`` which

* lets ``setup_interpolator(..., solver='b200')`` build every data family exactly as the reference does
  (value, gradient, norm, tangent, interface, frame orthogonality, fold rows ... all end up row-normalised in
  ``self.constraints[name] = {"matrix", "b", "w"}``, LI/_discrete_interpolator.py:316-389) but records only the
  WEIGHTS of the twelve stencil families instead of materialising their rows (``assemble_inner`` /
  ``assemble_borders``, LI/_finite_difference_interpolator.py:1052-1164);
* serves ``solve_system(solver='b200')`` (LI/_discrete_interpolator.py:1221-1319) on the device: constraint
  dict -> ELL rows through ``lsb200_add_rows``, ridge / column scaling / CG to the reference's stopping rule on
  the GPU, ``self.c`` back as a host ndarray;
* defers every other solver name to the reference (after re-assembling the stencil rows it had diverted).

``GeologicalModel.create_and_add_foliation('strati', solver='b200')`` then uses the engine unchanged, because
``solver`` rides in the builder's kwargs to both ``setup_interpolator`` and ``solve_system``
(LoopStructural/modelling/features/builders/_geological_feature_builder.py:558-569).

``install()`` also patches ``FaultSegment.apply_to_points`` / ``evaluate_displacement`` (the largest consumers of the
solved fields, LoopStructural/modelling/features/fault/_fault_segment.py:283-423): when the three coordinates of a
fault frame were solved by the engine and carry no faults / regions of their own, the whole integration runs in one
device pass (``fault.FaultFrame``); any other case is left to the stock method.  Likewise
``GeologicalFeatureBuilder.install_gradient_constraint`` (the frame-orthogonality families, :293-318 of
builders/_geological_feature_builder.py): when the source feature was solved by the engine its rows are generated on
the device from the resident field (``lsb200_add_field_orthogonal_rows``) instead of host evaluate -> einsum -> COO.
``BaseFeature.scalar_field`` (regular-grid export) takes its node values from ``lsb200_eval_lattice``, which
generates the lattice coordinates inside the kernel.

There is no CPU fallback: with ``solver='b200'`` and no usable CUDA device ``solve_system`` raises.
"""
from __future__ import annotations

from time import perf_counter

import numpy as np

from .interpolators import B200_SOLVERS

_installed = {}


def constraints_to_rows(constraints: dict):
    """Turn the reference's constraint dict into ELL row blocks: a list of (name, vals (n,k), cols (n,k) int64,
    b (n,), w (n,)).  Rows are already L2-normalised by add_constraints_to_least_squares; duplicate columns were
    merged by its coo->csc conversion (LI/_discrete_interpolator.py:361-366,384-386).  Unused slots hold
    value 0 on column 0."""
    out = []
    for name, c in constraints.items():
        w = np.asarray(c["w"], dtype=np.float64).ravel()
        if w.size == 0:
            continue  # build_matrix skips families with empty w (LI/_discrete_interpolator.py:694-696)
        m = c["matrix"].tocsr()
        m.sort_indices()
        n = m.shape[0]
        counts = np.diff(m.indptr)
        k = int(counts.max()) if n else 0
        if n == 0 or k == 0:
            continue
        vals = np.zeros((n, k))
        cols = np.zeros((n, k), dtype=np.int64)
        slot = np.arange(m.nnz) - np.repeat(m.indptr[:-1], counts)
        rows = np.repeat(np.arange(n), counts)
        vals[rows, slot] = m.data
        cols[rows, slot] = m.indices
        out.append((name, vals, cols, np.asarray(c["b"], dtype=np.float64).ravel(), w))
    return out


def _resolve_tolerances(tol, solver_kwargs):
    kw = dict(solver_kwargs or {})
    rtol = float(kw.get("rtol", 1e-5))
    atol = float(kw.get("atol", 0.0))
    if tol is not None and not ("atol" in kw and "rtol" in kw):
        atol = float(tol)  # LI/_solver_strategy.py:59-60
    return rtol, atol, kw


def _load_rows(it, eng, skip=()) -> int:
    """Every family of ``it.constraints`` except ``skip`` (those travel as stencil weights) -> ELL rows on the device."""
    n_rows = 0
    eng.clear_rows()
    rows = {k: c for k, c in it.constraints.items() if k not in skip}
    for _name, vals, cols, b, w in constraints_to_rows(rows):
        eng.add_rows(vals, cols, b, w, normalise=False)
        n_rows += vals.shape[0]
    for other, w2l, step, w, b, _stock, _pos in getattr(it, "_b200_device_rows", []):
        # frame-orthogonality families diverted by the builder patch: generated on the device from the other field
        n_rows += eng.add_field_orthogonal_rows(other, step=step, weight=w, b=b, other_world_to_local=w2l)
    for fam in getattr(it, "_b200_generated", []):
        # fold rows and direction-weighted face rows: generated on the device from points / faces and their vectors
        if fam["type"] == "element":
            n_rows += eng.add_point_rows(fam["kind"], fam["data"], fam["weight"])
        elif fam["type"] == "directional":
            n_rows += eng.add_directional_rows(fam["vectors"], fam["weight"])
        else:
            el = np.asarray(it.support.elements)
            rel = np.asarray(it.support.shared_element_relationships)
            n_rows += eng.add_face_vector_rows(el[rel[:, 0]], el[rel[:, 1]], fam["vectors"], fam["weight"])
    return n_rows


def _csr_to_ell(m):
    """Rows of a scipy sparse matrix as (vals (n,k), cols (n,k) int64), unused slots value 0 on column 0."""
    m = m.tocsr()
    m.sort_indices()
    n = m.shape[0]
    counts = np.diff(m.indptr)
    k = int(counts.max()) if n else 0
    vals = np.zeros((n, max(k, 1)))
    cols = np.zeros((n, max(k, 1)), dtype=np.int64)
    if k:
        slot = np.arange(m.nnz) - np.repeat(m.indptr[:-1], counts)
        rows = np.repeat(np.arange(n), counts)
        vals[rows, slot] = m.data
        cols[rows, slot] = m.indices
    return vals, cols


def _solve_admm_on_engine(it, eng, skip, stock_solve, tol, kw) -> bool:
    """solver='admm' with solver_kwargs['linsys_solver'] in B200_SOLVERS: the outer ADMM loop is the reference's own
    (LI/loopsolver/admm_solver.py:59-152: projections onto the bounds, dual updates, stopping test); its inner
    least-squares solves  min || [A; rho Q] dx - [b - A m; rho (z - u - Q m)] ||  (:243-316, by default lsmr, optionally
    Jacobi-preconditioned CG on the explicit normal equations) run on the device: the operator
    [A; rho Q]^T [A; rho Q] = S (N + ridge^2) S + rho^2 Q^T Q is applied matrix-free -- N by the engine's stencil / row
    kernels exactly as for solver='b200', rho Q as one more row block -- and only the right-hand side changes between
    the (up to 200) outer iterations (lsb200_set_rhs).  The reference column-scales A but not Q (the model it iterates on
    is y = S^-1 c); in c = S y the system reads (N + ridge^2 + Q'^T Q') c = S^-1 rhs_y with Q' = rho Q S^-1, which is what
    the engine solves, with its own diagonal scaling inside CG."""
    state = {"S": None, "rows_A": None, "qkey": None, "solves": 0, "iterations": 0, "rows": 0}
    stock_scaling = it.compute_column_scaling_matrix

    def capture_scaling(A):
        S = stock_scaling(A)
        state["S"] = np.asarray(S.diagonal(), dtype=np.float64)
        state["rows_A"] = A.shape[0]
        return S

    def inner(system_matrix, rhs, x0=None, **ikw):
        M = system_matrix.tocsr()
        n_a = state["rows_A"] if state["rows_A"] is not None else M.shape[0]
        S = state["S"] if state["S"] is not None else np.ones(M.shape[1])
        Qs = M[n_a:]
        key = (Qs.shape[0], Qs.nnz, float(Qs.data.sum()), float(np.abs(Qs.data).sum()), hash(Qs.indices[:256].tobytes()))
        if key != state["qkey"]:  # first solve, rho changed, or another active set: rebuild the row blocks
            state["rows"] = _load_rows(it, eng, skip)
            if Qs.shape[0]:
                Qc = Qs.copy()
                Qc.data = Qc.data / S[Qc.indices]
                vals, cols = _csr_to_ell(Qc)
                eng.add_rows(vals, cols, np.zeros(Qc.shape[0]), np.ones(Qc.shape[0]), normalise=False)
                state["rows"] += Qc.shape[0]
            state["qkey"] = key
        eng.set_rhs((M.T @ np.asarray(rhs, dtype=np.float64).ravel()) / S)
        rtol = float(ikw.get("rtol", ikw.get("btol", 1e-6)))
        atol = float(ikw["atol"]) if ("atol" in ikw and "rtol" in ikw) else 0.0
        c, res = eng.solve_cg(rtol=rtol, atol=atol, maxiter=ikw.get("maxiter"), x0=None if x0 is None else np.asarray(x0) * S)
        state["solves"] += 1
        state["iterations"] += res.iterations
        return c / S

    eng.set_options(ridge_factor=it.ridge_factor, add_ridge=it.add_ridge_regulatisation, apply_scaling=it.apply_scaling_matrix)
    kw = dict(kw)
    kw["linsys_solver"] = inner
    it.compute_column_scaling_matrix = capture_scaling  # instance attribute: read at call time by _preprocess_main_system
    try:
        ok = stock_solve(solver="admm", tol=tol, solver_kwargs=kw)
    finally:
        del it.compute_column_scaling_matrix
    if ok and state["solves"] > 0:
        eng.set_c(np.asarray(it.c, dtype=np.float64))  # the field lives on the device again: evaluation stays there
        it.latest_solve_timing = dict(it.latest_solve_timing, inner_backend="b200", inner_solves=state["solves"],
                                      inner_iterations=state["iterations"], matrix_rows_device=state["rows"],
                                      b200_field_resident=True)
    return ok


def _solve_on_engine(it, eng, solver, tol, solver_kwargs, skip=()) -> bool:
    """Common tail of solve_system(solver='b200'): constraint dict -> ELL rows -> device CG -> it.c.  ``skip``: keys of
    ``it.constraints`` that reach the engine as stencil weights instead of rows."""
    started = perf_counter()
    rtol, atol, kw = _resolve_tolerances(tol, solver_kwargs)
    eng.set_options(ridge_factor=it.ridge_factor, add_ridge=it.add_ridge_regulatisation,
                    apply_scaling=it.apply_scaling_matrix, deterministic=bool(kw.get("deterministic", False)))
    t_asm = perf_counter()
    n_rows = _load_rows(it, eng, skip)
    assembly = perf_counter() - t_asm
    c, res = eng.solve_cg(rtol=rtol, atol=atol, maxiter=kw.get("maxiter"), x0=kw.get("x0"))
    it.c = c  # an ndarray that owns its memory: callers may keep it across solves and past the interpolator's life
    it.up_to_date = True  # non-convergence is a warning, never a failure (LI/_solver_strategy.py:66-71)
    it.latest_solve_timing = {
        "solver": solver, "backend": "b200", "assembly_seconds": assembly, "preprocess_seconds": res.setup_seconds,
        "solve_seconds": res.solve_seconds, "total_seconds": perf_counter() - started, "iterations": res.iterations,
        "info": res.info, "matrix_rows": n_rows, "matrix_cols": int(it.dof), "up_to_date": True,
    }
    return True


class _B200Mixin:
    """Shared by the FDI and P1 subclasses: set-up bookkeeping, solver dispatch, device evaluation.

    Two set-up states serve ``solver='b200'``:
      * *diverted* (``setup_interpolator(..., solver='b200')``, the GeologicalModel flow): the O(N) regularisation
        families were recorded as weights and never assembled;
      * *explicit* (any other set-up, e.g. loop_interpolation's InterpolatorBuilder, which pops ``solver`` from its kwargs
        before it calls ``setup``): the families are in ``self.constraints``; those recognised as default stencil
        families at assembly time are still applied matrix-free, everything else goes to the engine as rows.
    Neither direction re-runs the set-up, so constraints added after it (orthogonal rows, equality constraints, the
    builder's gradient constraints) are kept, exactly as with the stock class."""

    def _b200_init(self):
        self._b200_divert = False
        self._b200_reg = {}        # diverted families: name -> effective weight
        self._b200_reg_ops = {}    # diverted families: name -> what the stock assembly needs to materialise them
        self._b200_explicit = {}   # explicit set-up: constraints key -> (stencil name, weight) of a default family
        self._b200_setup_kwargs = {}
        self._b200_engine = None
        # (other engine, its world_to_local, element step, weight, b, stock, position) per diverted orthogonality
        # family; stock() assembles the same family on the host (used when a stock solver is asked for after all)
        self._b200_device_rows = []
        # families generated on the device from host vectors (fold rows, direction-weighted face rows): dicts with
        # type 'element' (kind, data (n,7), weight) or 'face' (vectors (n_faces,3), weight), name, pos, stock
        self._b200_generated = []

    def setup_interpolator(self, **kwargs):
        self._b200_setup_kwargs = dict(kwargs)
        self._b200_reg, self._b200_reg_ops, self._b200_explicit = {}, {}, {}
        self._b200_device_rows = []
        self._b200_generated = []
        self._b200_divert = (self._b200_supported(kwargs) and kwargs.get("solver") in B200_SOLVERS
                             and not kwargs.get("use_regularisation_weight_scale", False))
        return super().setup_interpolator(**kwargs)

    def reset(self):
        super().reset()
        if hasattr(self, "_b200_generated"):
            # the stock reset empties self.constraints (LI/_discrete_interpolator.py:300-314): what a diverted set-up recorded
            # instead of assembling goes with it
            self._b200_reg, self._b200_reg_ops, self._b200_explicit = {}, {}, {}
            self._b200_device_rows, self._b200_generated = [], []
        if getattr(self, "_b200_divert", False):
            # DiscreteInterpolator.reset allocates np.ones(dof) (LI/_discrete_interpolator.py:313); only the explicit
            # stencil assembly with use_regularisation_weight_scale reads it, and that case is never diverted
            ones = np.broadcast_to(np.float64(1.0), (int(self.dof),))
            self.regularisation_scale = ones

    def _b200_record(self, name, payload):
        """A family that is not assembled now: remember where it would have stood in self.constraints."""
        self._b200_reg_ops[name] = (len(self.constraints), payload)

    def _b200_assemble_recorded(self, name, payload):
        raise NotImplementedError

    def _b200_materialise(self):
        """Assemble on the host what a diverted set-up only recorded (a stock solver was asked for after all), each
        family at the place the stock set-up would have put it, so that the stock solver sees the stock matrix."""
        self._b200_divert = False  # from here on the assembly methods take the stock path
        self.regularisation_scale = np.ones(self.dof)
        todo = [(pos, lambda n=name, p=payload: self._b200_assemble_recorded(n, p)) for name, (pos, payload) in self._b200_reg_ops.items()]
        todo += [(entry[6], entry[5]) for entry in self._b200_device_rows]
        todo += [(fam["pos"], fam["stock"]) for fam in self._b200_generated]
        placed = []
        for pos, run in todo:
            before = set(self.constraints)
            run()
            placed.append((pos, [k for k in self.constraints if k not in before]))
        moved = {k for _pos, keys in placed for k in keys}
        order = [k for k in self.constraints if k not in moved]
        shift = 0
        for pos, keys in sorted(placed, key=lambda t: t[0]):  # stable: equal positions keep their recording order
            order[pos + shift:pos + shift] = keys
            shift += len(keys)
        self.constraints = {k: self.constraints[k] for k in order}
        self._b200_reg, self._b200_reg_ops, self._b200_device_rows, self._b200_generated = {}, {}, [], []

    def _b200_get_engine(self):
        """One engine per interpolator, kept across solves (7 node vectors of device memory)."""
        sup = self.support
        key = (tuple(int(v) for v in sup.nsteps), tuple(float(v) for v in sup.origin), tuple(float(v) for v in sup.step_vector))
        if self._b200_engine is not None and getattr(self, "_b200_engine_key", None) != key:
            self._b200_engine.close()
            self._b200_engine = None
        if self._b200_engine is None:
            self._b200_engine = self._b200_make_engine()
            self._b200_engine_key = key
        return self._b200_engine

    def solve_system(self, solver=None, tol=None, solver_kwargs=None) -> bool:
        if solver == "admm" and (solver_kwargs or {}).get("linsys_solver") in B200_SOLVERS:
            # inequality-constrained solve: the reference's ADMM loop with its inner solves on the device
            if not self._b200_supported(self._b200_setup_kwargs):
                raise ValueError(self._b200_unsupported)
            if self._b200_divert:
                self._b200_materialise()  # the stock pipeline assembles the explicit A for its own bookkeeping
            eng = self._b200_get_engine()
            skip = self._b200_set_regularisation(eng)
            return _solve_admm_on_engine(self, eng, skip, super().solve_system, tol, solver_kwargs)
        if solver not in B200_SOLVERS:
            if self._b200_divert:
                self._b200_materialise()
            return super().solve_system(solver=solver, tol=tol, solver_kwargs=solver_kwargs)
        if not self._b200_supported(self._b200_setup_kwargs):
            raise ValueError(self._b200_unsupported)
        if not self._pre_solve():
            raise ValueError("Pre solve failed")
        eng = self._b200_get_engine()
        skip = self._b200_set_regularisation(eng)
        return _solve_on_engine(self, eng, solver, tol, solver_kwargs, skip=skip)

    # -- evaluation: the device holds c after a b200 solve -------------------------------------------------
    def _b200_engine_current(self):
        t = self.latest_solve_timing
        if self._b200_engine is None or not self.up_to_date or not (t.get("backend") == "b200" or t.get("b200_field_resident")):
            return None
        return self._b200_engine

    def _b200_world_to_local(self):
        bb = getattr(self, "bounding_box", None)
        return None if bb is None else np.asarray(bb.world_to_local_matrix, dtype=float)

    def evaluate_value(self, locations):
        """GeologicalInterpolator.evaluate_value (LI/_geological_interpolator.py:723-730) with the world -> local
        projection on the device: the reference builds an (N,4) homogeneous copy of the points and multiplies it by
        the 4 x 4 matrix on the host (BoundingBox._apply_affine, LC/geometry/_bounding_box.py:154-176) before every
        evaluation; lsb200_faulted_eval with no faults takes the world points and the matrix."""
        self.update()
        eng = self._b200_engine_current()
        locations = np.asarray(locations, dtype=np.float64)
        if eng is None or locations.ndim != 2 or locations.shape[1] != 3:
            return super().evaluate_value(locations)
        from . import fault as fl

        return fl.faulted_evaluate([], [], [], [], eng, locations, gradient=False, world_to_local=self._b200_world_to_local())

    def evaluate_gradient(self, locations):
        """GeologicalInterpolator.evaluate_gradient (:732-741): projection, gather and the rotation of the gradients back
        to world axes (reproject_vectors) in one device call."""
        self.update()
        eng = self._b200_engine_current()
        locations = np.asarray(locations, dtype=np.float64)
        if eng is None or locations.ndim != 2 or locations.shape[1] != 3:
            return super().evaluate_gradient(locations)
        if locations.shape[0] == 0:
            return np.zeros((0, 3))
        from . import fault as fl

        return fl.faulted_evaluate([], [], [], [], eng, locations, gradient=True, world_to_local=self._b200_world_to_local())

    def _evaluate_value_local(self, locations):
        self.update()
        eng = self._b200_engine_current()
        if eng is None:
            return super()._evaluate_value_local(locations)
        return eng.evaluate_value(np.asarray(locations, dtype=np.float64))

    def _evaluate_gradient_local(self, locations):
        self.update()
        eng = self._b200_engine_current()
        if eng is None:
            return super()._evaluate_gradient_local(locations)
        locations = np.asarray(locations, dtype=np.float64)
        if locations.shape[0] == 0:
            return np.zeros((0, 3))
        return eng.evaluate_gradient(locations)


def _uniform_weight(c, w) -> bool:
    ww = np.asarray(c["w"], dtype=np.float64).ravel()
    return ww.size > 0 and bool(np.all(ww == w))


def _make_fdi_class():
    from loop_common.supports import StructuredGrid
    from loop_interpolation import FiniteDifferenceInterpolator
    from loop_interpolation._operator import Operator

    default_masks = {"dxx": Operator.Dxx_mask, "dyy": Operator.Dyy_mask, "dzz": Operator.Dzz_mask,
                     "dxy": Operator.Dxy_mask, "dxz": Operator.Dxz_mask, "dyz": Operator.Dyz_mask}
    border_axis = {"dx_lower": "dx", "dx_upper": "dx", "dy_lower": "dy", "dy_upper": "dy", "dz_lower": "dz", "dz_upper": "dz"}

    class B200FDI(_B200Mixin, FiniteDifferenceInterpolator):
        """FiniteDifferenceInterpolator whose ``solver='b200'`` runs on the GPU engine."""

        _b200_unsupported = ("solver='b200' needs a 3-D regular StructuredGrid (RectilinearGrid and 2-D grids stay on "
                             "the stock solvers)")

        def __init__(self, grid, data=None):
            super().__init__(grid, data=data)
            self._b200_init()

        @property
        def _b200_stencil(self):
            return self._b200_reg

        def _b200_supported(self, kwargs) -> bool:
            # the exact class: RectilinearGrid subclasses StructuredGrid (LC/supports/_3d_rectilinear_grid.py:18) but its
            # node spacing is not uniform, and the engine gathers with uniform-cell trilinear weights
            sup = self.support
            return type(sup) is StructuredGrid and getattr(sup, "dimension", 3) == 3

        def _b200_matrix_free_family(self, operator, w, name) -> bool:
            """A default second-derivative family with one scalar weight: what the engine applies matrix-free."""
            return (name in default_masks and operator is not None and np.ndim(w) == 0
                    and np.shape(operator) == np.shape(default_masks[name]) and np.array_equal(operator, default_masks[name])
                    and self._b200_supported(self._b200_setup_kwargs) and not self.use_regularisation_weight_scale
                    and not getattr(self, "regularisation_matrix_free", False))

        def assemble_inner(self, operator, w, name="regularisation"):
            ok = self._b200_matrix_free_family(operator, w, name)
            if self._b200_divert and ok and name not in self._b200_reg:
                # w is the effective row weight (mixed terms already / 4, LC/supports/_3d_structured_grid.py:502-509)
                self._b200_reg[name] = float(w)
                self._b200_record(name, (operator, w))
                return
            before = set(self.constraints)
            out = super().assemble_inner(operator, w, name=name)
            new = set(self.constraints) - before
            if ok and len(new) == 1 and name not in {v[0] for v in self._b200_explicit.values()}:
                self._b200_explicit[new.pop()] = (name, float(w))
            return out

        def assemble_borders(self):
            ok = self._b200_supported(self._b200_setup_kwargs) and not self.use_regularisation_weight_scale and \
                not getattr(self, "regularisation_matrix_free", False)
            if self._b200_divert and ok:
                for k in ("dx", "dy", "dz"):
                    self._b200_reg[k] = float(self.interpolation_weights[k])
                self._b200_record("borders", None)
                return
            before = set(self.constraints)
            out = super().assemble_borders()
            if ok:
                for key in set(self.constraints) - before:
                    if key in border_axis:
                        self._b200_explicit[key] = (border_axis[key], float(self.interpolation_weights[border_axis[key]]))
            return out

        def add_norm_constraints(self, w=1.0):
            """FDI.add_norm_constraints (LI/_finite_difference_interpolator.py:403-476) without its KD-tree pass over ALL
            nodes (compute_weighting over support.nodes, :436-444: O(N log n) time and a 24 B/node table) -- that pass only
            feeds regularisation_scale, which nothing reads unless use_regularisation_weight_scale is set."""
            if not self._b200_divert or self.use_regularisation_weight_scale:
                return super().add_norm_constraints(w)
            pts = self.get_norm_constraints()
            if pts.shape[0] == 0:
                return
            dim = self.support.dimension
            corners, inside = self.support.position_to_cell_corners(pts[:, :dim])
            corners = np.asarray(corners, dtype=int)[inside, :]
            T = self.support.get_element_gradient_for_location(pts[inside, :dim])[1]
            per_point = pts[inside, 2 * dim] if pts.shape[1] > 2 * dim else np.ones(int(inside.sum()))
            share = w / dim  # one row per axis: the point's weight is shared between them (:455-460)
            weights = (share[inside] if isinstance(share, np.ndarray) else float(share)) * per_point
            for d in range(dim):
                self.add_constraints_to_least_squares(T[:, d, :], pts[inside, dim + d], corners, w=weights, name=f"norm_{d}")
            self.up_to_date = False

        def minimise_directional_gradient_change(self, w, vector, name="directional regularisation"):
            """FDI.minimise_directional_gradient_change (LI/_finite_difference_interpolator.py:1207-1340) for a diverted
            set-up: six families of one row per interior node (index tables, tiled masks, six sparse matrices on the
            host) become one record -- the direction field -- and the rows are generated on the device
            (lsb200_add_directional_rows)."""
            vec = None if vector is None else np.asarray(vector)
            if (self._b200_divert and vec is not None and vec.ndim == 2 and vec.shape == (self.support.n_nodes, 3)
                    and np.ndim(w) == 0 and self._b200_supported(self._b200_setup_kwargs)
                    and not self.use_regularisation_weight_scale):
                stock = super().minimise_directional_gradient_change
                self._b200_generated.append(dict(type="directional", vectors=np.ascontiguousarray(vec, dtype=np.float64),
                                                 weight=float(w), name=name, pos=len(self.constraints),
                                                 stock=lambda: stock(w, vector, name=name)))
                self.up_to_date = False
                return
            return super().minimise_directional_gradient_change(w, vector, name=name)

        def _b200_assemble_recorded(self, name, payload):
            if name == "borders":
                self.assemble_borders()
            else:
                self.assemble_inner(payload[0], payload[1], name=name)

        def _b200_make_engine(self):
            from .engine import Engine

            sup = self.support
            return Engine(np.asarray(sup.nsteps), np.asarray(sup.origin, float), np.asarray(sup.step_vector, float))

        def _b200_set_regularisation(self, eng):
            """Stencil weights of this solve; returns the keys of self.constraints they replace."""
            st, skip = dict(self._b200_reg), set()
            if not self._b200_divert:
                st = {}
                borders = {}
                for key, (name, w) in self._b200_explicit.items():
                    c = self.constraints.get(key)
                    if c is None or not _uniform_weight(c, w):
                        continue
                    if name in ("dx", "dy", "dz"):
                        borders.setdefault(name, []).append(key)
                    elif name not in st:
                        st[name] = w
                        skip.add(key)
                for name, keys in borders.items():
                    if len(keys) == 2:  # the lower and the upper plane: the engine applies both or neither
                        st[name] = self._b200_explicit[keys[0]][1]
                        skip.update(keys)
            # effective weights go in directly: the engine's dict API divides the mixed terms by 4 itself
            eng.set_regularisation({"dxx": st.get("dxx", 0.0), "dyy": st.get("dyy", 0.0), "dzz": st.get("dzz", 0.0),
                                    "dxy": 4.0 * st.get("dxy", 0.0), "dxz": 4.0 * st.get("dxz", 0.0),
                                    "dyz": 4.0 * st.get("dyz", 0.0), "dx": st.get("dx", 0.0), "dy": st.get("dy", 0.0),
                                    "dz": st.get("dz", 0.0)})
            return skip

    B200FDI.__name__ = "B200FiniteDifferenceInterpolator"
    return B200FDI


def _make_tet_class(base, name, doc):
    """B200 subclass of a TetMesh interpolator (P1Interpolator or a class derived from it)."""
    from loop_common.supports import TetMesh

    class B200Tet(_B200Mixin, base):
        _b200_unsupported = "solver='b200' needs a 3-D structured TetMesh"

        def __init__(self, *args, **kwargs):
            super().__init__(*args, **kwargs)
            self._b200_init()

        def _b200_supported(self, kwargs) -> bool:
            return type(self.support) is TetMesh

        def minimise_edge_jumps(self, w=0.1, vector_func=None, vector=None, name="edge jump"):
            plain = (vector_func is None and vector is None and name == "edge jump" and np.ndim(w) == 0
                     and self._b200_supported(self._b200_setup_kwargs)
                     and not self._b200_setup_kwargs.get("use_regularisation_weight_scale", False))
            if self._b200_divert and plain and "cgw" not in self._b200_reg:
                self._b200_reg["cgw"] = float(w)
                self._b200_record("cgw", w)
                return
            if (self._b200_divert and vector_func is None and vector is not None and np.ndim(w) == 0
                    and self._b200_supported(self._b200_setup_kwargs)
                    and not self._b200_setup_kwargs.get("use_regularisation_weight_scale", False)):
                # a direction per shared face (add_directional_regularisation, the fold regularisation families): the rows
                # v . (grad phi_T1 - grad phi_T2) = 0 are generated on the device from the two tetrahedra of a face.  A
                # vector of any other length is ignored by the reference (LI/_p1interpolator.py:130-131) and a NaN drops
                # the family (LI/_discrete_interpolator.py:357-359): both stay with the stock method.
                vec = np.asarray(vector, dtype=np.float64)
                cells = [int(v) - 1 for v in self.support.nsteps]
                n_faces = 10 * cells[0] * cells[1] * cells[2] - 2 * (cells[0] * cells[1] + cells[1] * cells[2] + cells[2] * cells[0])
                if vec.shape == (n_faces, 3) and not np.isnan(vec).any():
                    stock = super().minimise_edge_jumps
                    self._b200_generated.append(dict(
                        type="face", vectors=vec, weight=float(w), name=name, pos=len(self.constraints),
                        stock=lambda: stock(w=w, vector_func=None, vector=vector, name=name)))
                    self.up_to_date = False
                    return
            before = set(self.constraints)
            out = super().minimise_edge_jumps(w=w, vector_func=vector_func, vector=vector, name=name)
            new = set(self.constraints) - before
            if plain and len(new) == 1 and not self._b200_explicit:
                self._b200_explicit[new.pop()] = ("cgw", float(w))
            return out

        def _b200_assemble_recorded(self, name, payload):
            self.minimise_edge_jumps(w=payload)

        def _b200_make_engine(self):
            from .engine import Engine

            sup = self.support
            return Engine(np.asarray(sup.nsteps), np.asarray(sup.origin, float), np.asarray(sup.step_vector, float),
                          kind="tetmesh")

        def _b200_set_regularisation(self, eng):
            skip = set()
            cgw = self._b200_reg.get("cgw", 0.0)
            if not self._b200_divert:
                cgw = 0.0
                for key, (_n, w) in self._b200_explicit.items():
                    c = self.constraints.get(key)
                    if c is not None and _uniform_weight(c, w):
                        cgw = w
                        skip.add(key)
            eng.set_p1_regularisation(cgw)
            return skip

    B200Tet.__name__ = B200Tet.__qualname__ = name
    B200Tet.__doc__ = doc
    return B200Tet


def _make_p1_class():
    from loop_interpolation import P1Interpolator

    return _make_tet_class(P1Interpolator, "B200P1Interpolator",
                           "P1Interpolator whose ``solver='b200'`` runs on the GPU engine: the face gradient-jump rows of "
                           "minimise_edge_jumps (LI/_p1interpolator.py:115-166; 99 % of A) are applied matrix-free.")


def _make_dfi_class():
    """DiscreteFoldInterpolator (LI/_discrete_fold_interpolator.py:21-242; every folded foliation / folded fold frame is
    built with it, LoopStructural/modelling/core/_model_feature_factory.py:248-250) on the same engine path as P1: its
    fold orientation / axis / normalisation rows and its direction-weighted face rows ('fold regularisation n',
    minimise_edge_jumps with a vector per shared face) are generated on the device from the fold vectors the reference's
    FoldEvent computes on the host; a plain 'edge jump' family, when the builder asks for one, is applied matrix-free
    like P1's."""
    import inspect

    from loop_common.math import rng  # the module-level generator the reference shuffles the elements with
    from loop_interpolation._discrete_fold_interpolator import DiscreteFoldInterpolator

    from . import engine as en

    base = _make_tet_class(DiscreteFoldInterpolator, "B200DiscreteFoldInterpolator",
                           "DiscreteFoldInterpolator whose ``solver='b200'`` runs on the GPU engine.")
    stock_sig = inspect.signature(DiscreteFoldInterpolator.add_fold_constraints)

    class B200DFI(base):
        def add_fold_constraints(self, *args, **kwargs):
            """DiscreteFoldInterpolator.add_fold_constraints (LI/_discrete_fold_interpolator.py:66-242) for a diverted
            set-up: the same elements (the reference's generator is shuffled exactly as the stock method shuffles it), the
            same fold vectors (FoldEvent.get_deformed_orientation, host code of the reference), but no element-gradient
            table, no einsum and no sparse matrix on the host: a family is its barycentres, vectors and target, and its
            rows are generated on the device at solve time (LSB200_ROWS_ELEMENT_VECTOR); the three direction-weighted
            face families go through minimise_edge_jumps above (lsb200_add_face_vector_rows)."""
            a = stock_sig.bind(self, *args, **kwargs)
            a.apply_defaults()
            a = a.arguments
            if not self._b200_divert or "dgz_alignment" not in a:
                return super().add_fold_constraints(*args, **kwargs)
            from loop_interpolation._fold_norm_alignment import resolve_fold_norm_target

            sup, step, mask_fn = self.support, a["step"], a["mask_fn"]
            bary = np.asarray(sup.barycentre, dtype=np.float64)
            deformed, axis, dgz = self.fold.get_deformed_orientation(bary)
            element_idx = np.arange(sup.n_elements)
            rng.shuffle(element_idx)
            weight = np.ones(sup.n_elements, dtype=float)
            if mask_fn is not None:
                weight[mask_fn(bary)] = 0
            for name, w, vec in (("fold orientation", a["fold_orientation"], deformed), ("fold axis", a["fold_axis_w"], axis),
                                 ("fold normalisation", a["fold_normalisation"], dgz)):
                if w is None:
                    continue
                rng.shuffle(element_idx)
                sel = element_idx[::step].copy()
                target = 0.0
                if name == "fold normalisation":
                    target = w * resolve_fold_norm_target(
                        fold=self.fold, normal_constraints=self.get_norm_constraints(), fold_norm=a["fold_norm"],
                        dgz_alignment=a["dgz_alignment"], logger=_fold_logger(), default_fold_norm=-1.0)
                self._b200_add_element_family(name, sel, np.asarray(vec, dtype=np.float64), float(w), float(target), weight, bary)
            reg = a["fold_regularisation"]
            if reg is not None:
                # add_directional_regularisation (LI/_discrete_interpolator.py:413-451) with the three providers of the stock
                # method (normal, deformed orientation, axis at the shared-face centres), one fold evaluation instead of three
                pts = np.asarray(self.get_regularisation_sample_points(), dtype=float)
                d_f, a_f, n_f = self.fold.get_deformed_orientation(pts)
                for i, vec in enumerate((n_f, d_f, a_f)):
                    if reg[i] == 0:
                        continue
                    vec = np.asarray(vec, dtype=float)
                    if mask_fn is not None:
                        vec = vec.copy()
                        vec[mask_fn(pts)] = 0.0
                    if vec.shape != pts.shape:
                        continue
                    self._add_directional_regularisation(reg[i], vec, name=f"fold regularisation {i + 1}")

        def _b200_add_element_family(self, name, sel, vec, w, target, weight, bary):
            v = vec[sel]
            sup = self.support

            def stock():
                # the same family on the host (a stock solver was asked for after all): vol (v . eg) x = vol target
                eg = sup.get_element_gradients(sel)
                idc = sup.get_elements()[sel]
                nodes = sup.nodes[idc]
                vol = np.abs(np.linalg.det(nodes[:, 1:, :] - nodes[:, 0, None, :])) / 6
                self.add_constraints_to_least_squares(np.einsum("ij,ijk->ik", v, eg) * vol[:, None], np.ones(len(sel)) * target * vol,
                                                      idc, w=weight[sel] * w, name=name)

            if np.isnan(v).any() or not np.isfinite(target):
                return stock()  # the reference drops such a family with a warning (LI/_discrete_interpolator.py:357-359)
            keep = weight[sel] > 0
            data = np.column_stack([bary[sel][keep], v[keep], np.full(int(keep.sum()), target)])
            self._b200_generated.append(dict(type="element", kind=en.ROWS_ELEMENT_VECTOR, data=data, weight=w, name=name,
                                             pos=len(self.constraints), stock=stock))
            self.up_to_date = False

    B200DFI.__name__ = B200DFI.__qualname__ = "B200DiscreteFoldInterpolator"
    B200DFI.__doc__ = base.__doc__
    return B200DFI


def _fold_logger():
    import logging

    return logging.getLogger("loop_interpolation._discrete_fold_interpolator")


# -- fault frames: FaultSegment.apply_to_points / evaluate_displacement on the device -------------------------------
def _frame_of(segment):
    """FaultFrame over the three coordinate features of a FaultSegment, or None when the reference has to do the
    work: a coordinate that is not a plain interpolated GeologicalFeature, that carries faults or regions of its
    own, or whose interpolator was not solved by the grid engine."""
    from . import fault as fl

    engines, transforms = [], []
    for i in range(3):
        feat = segment[i]
        if type(feat).__name__ != "GeologicalFeature":
            return None
        if len(getattr(feat, "regions", [])) > 0 or (getattr(feat, "faults_enabled", True) and len(getattr(feat, "faults", [])) > 0):
            return None
        if getattr(feat, "builder", None) is not None:
            feat.builder.up_to_date()  # lazy solve, as GeologicalFeature.evaluate_value does (_geological_feature.py:139-140)
        it = getattr(feat, "interpolator", None)
        eng = it._b200_engine_current() if hasattr(it, "_b200_engine_current") else None
        if eng is None or getattr(eng, "kind", "grid") not in ("grid", "tetmesh"):
            return None
        engines.append(eng)
        bb = getattr(it, "bounding_box", None)
        transforms.append(np.eye(4) if bb is None else np.asarray(bb.world_to_local_matrix, dtype=float))
    key = tuple(id(e) for e in engines)
    cached = getattr(segment, "_b200_frame", None)
    if cached is None or cached[0] != key:
        cached = (key, fl.FaultFrame(engines, transforms))
        segment._b200_frame = cached
    return cached[1]


def _make_fault_methods(stock_apply, stock_displacement):
    from . import fault as fl

    def apply_to_points(self, points, reverse=False, precomputed_gx=None):
        """FaultSegment.apply_to_points (LS/modelling/features/fault/_fault_segment.py:315-423) in one device pass when
        the three frame coordinates were solved by the B200 engine; the stock method otherwise."""
        frame = prof = None
        if precomputed_gx is None:
            prof = fl.profile_from_reference(self.faultfunction)
            if prof is not None:
                frame = _frame_of(self)
        if frame is None:
            return stock_apply(self, points, reverse=reverse, precomputed_gx=precomputed_gx)
        return frame.apply_to_points(np.array(points, dtype=float), prof, float(self.displacement), int(self.steps))

    def evaluate_displacement(self, points):
        prof = fl.profile_from_reference(self.faultfunction)
        frame = _frame_of(self) if prof is not None else None
        if frame is None:
            return stock_displacement(self, points)
        return frame.evaluate_displacement(np.array(points, dtype=float), prof, float(self.displacement),
                                           float(self.fault_offset))

    return apply_to_points, evaluate_displacement


def _solved_grid_field(feature, kinds=("grid",), allow_faults=False):
    """(engine, world_to_local) of a plain interpolated feature solved by the engine on one of ``kinds``, else None.
    (Outside a tet mesh the gradient is NaN and the reference then drops the WHOLE orthogonality family,
    LI/_discrete_interpolator.py:357-359: lsb200_add_field_orthogonal_rows does the same.)"""
    if type(feature).__name__ != "GeologicalFeature":
        return None
    if not allow_faults and getattr(feature, "faults_enabled", True) and len(getattr(feature, "faults", [])) > 0:
        return None
    if getattr(feature, "builder", None) is not None:
        feature.builder.up_to_date()
    it = getattr(feature, "interpolator", None)
    eng = it._b200_engine_current() if hasattr(it, "_b200_engine_current") else None
    if eng is None or getattr(eng, "kind", "grid") not in kinds:
        return None
    bb = getattr(it, "bounding_box", None)
    return eng, (np.eye(4) if bb is None else np.asarray(bb.world_to_local_matrix, dtype=float))


def _make_install_gradient_constraint(stock):
    def install_gradient_constraint(self):
        """GeologicalFeatureBuilder.install_gradient_constraint
        (LS/modelling/features/builders/_geological_feature_builder.py:293-318): families whose source feature was
        solved by the engine (grid or tet mesh; the target may be either, too) are recorded and generated on the device at solve time
        (lsb200_add_field_orthogonal_rows); the others go through the stock method."""
        it = self.interpolator
        if not (getattr(it, "_b200_divert", False) and hasattr(it, "_b200_device_rows")):
            return stock(self)
        rest = {}

        def only(key, g):
            def run():  # the same family through the stock method (a stock solver was asked for after a diverted set-up)
                keep = self._orthogonal_features
                self._orthogonal_features = {key: g}
                try:
                    stock(self)
                finally:
                    self._orthogonal_features = keep
            return run

        for key, g in self._orthogonal_features.items():
            feature, w, _region, step, b = g
            if w == 0:
                continue
            src = _solved_grid_field(feature, kinds=("grid", "tetmesh"))
            if src is None:
                rest[key] = g
            else:
                it._b200_device_rows.append((src[0], src[1], int(step), float(w), float(b), only(key, g), len(it.constraints)))
        if rest:
            keep = self._orthogonal_features
            self._orthogonal_features = rest
            try:
                stock(self)
            finally:
                self._orthogonal_features = keep

    return install_gradient_constraint


def _install_builder_patch():
    try:
        from LoopStructural.modelling.features.builders import GeologicalFeatureBuilder
    except Exception:
        return
    if "builder_stock" not in _installed:
        _installed["builder_stock"] = (GeologicalFeatureBuilder, GeologicalFeatureBuilder.install_gradient_constraint)
    cls, stock = _installed["builder_stock"]
    cls.install_gradient_constraint = _make_install_gradient_constraint(stock)


def _make_scalar_field(stock):
    def scalar_field(self, bounding_box=None):
        """BaseFeature.scalar_field (LS/modelling/features/_base_geological_feature.py:359-388): node values on
        BoundingBox.regular_grid(local=False, order='F') from lsb200_eval_lattice (coordinates generated in the
        kernel) when the feature is a plain one solved by the grid engine; the stock method otherwise."""
        src = None
        if type(self).__name__ == "GeologicalFeature" and len(getattr(self, "regions", [])) == 0:
            src = _solved_grid_field(self, kinds=("grid", "tetmesh"))
        if src is None:
            return stock(self, bounding_box)
        from LoopStructural.geometry import StructuredGrid

        if bounding_box is None:
            if self.model is None:
                raise ValueError("Must specify bounding box")
            bounding_box = self.model.bounding_box
        grid = StructuredGrid(origin=bounding_box.origin, step_vector=bounding_box.step_vector, nsteps=bounding_box.nsteps,
                              name=self.name)
        eng, w2l = src
        grid.properties[self.name] = eng.evaluate_lattice(np.asarray(bounding_box.origin, dtype=float),
                                                          np.asarray(bounding_box.maximum, dtype=float),
                                                          np.asarray(bounding_box.nsteps), order="F", world_to_local=w2l)
        # the cell-centre cloud is not a lattice to the bit (projected, shifted, reprojected): ordinary evaluation
        grid.cell_properties[self.name] = self.evaluate_value(bounding_box.reproject(bounding_box.cell_centres(order="F")))
        return grid

    return scalar_field


def _make_gradient_norm_scalar_field(stock):
    def gradient_norm_scalar_field(self, bounding_box=None):
        """BaseFeature.gradient_norm_scalar_field (LS/modelling/features/_base_geological_feature.py:390-426): the node
        values are |grad f| on BoundingBox.regular_grid(local=False, order='F') -- the world gradient on that lattice from
        lsb200_eval_lattice (coordinates generated in the kernel, gradients rotated back to world axes there) when the
        feature is a plain one solved by the engine; the stock method otherwise."""
        src = None
        if type(self).__name__ == "GeologicalFeature" and len(getattr(self, "regions", [])) == 0:
            src = _solved_grid_field(self, kinds=("grid", "tetmesh"))
        if src is None:
            return stock(self, bounding_box)
        from LoopStructural.geometry import StructuredGrid

        if bounding_box is None:
            if self.model is None:
                raise ValueError("Must specify bounding box")
            bounding_box = self.model.bounding_box
        grid = StructuredGrid(origin=bounding_box.origin, step_vector=bounding_box.step_vector, nsteps=bounding_box.nsteps,
                              name=self.name)
        eng, w2l = src
        g = eng.evaluate_lattice(np.asarray(bounding_box.origin, dtype=float), np.asarray(bounding_box.maximum, dtype=float),
                                 np.asarray(bounding_box.nsteps), order="F", gradient=True, world_to_local=w2l)
        grid.properties[self.name] = np.linalg.norm(g, axis=1)
        centres = bounding_box.reproject(bounding_box.cell_centres(order="F"))
        grid.cell_properties[self.name] = np.linalg.norm(self.evaluate_gradient(centres), axis=1)
        return grid

    return gradient_norm_scalar_field


def _fault_chain_of(feature):
    """(frames, profiles, displacements, steps) of the faults that cut ``feature`` when every one of them can run on the
    engine, else None."""
    from . import fault as fl

    faults = list(getattr(feature, "faults", [])) if getattr(feature, "faults_enabled", True) else []
    frames, profiles = [], []
    for f in faults:
        prof = fl.profile_from_reference(getattr(f, "faultfunction", None)) if type(f).__name__ == "FaultSegment" else None
        frame = _frame_of(f) if prof is not None else None
        if frame is None:
            return None
        frames.append(frame)
        profiles.append(prof)
    return frames, profiles, [float(f.displacement) for f in faults], [int(f.steps) for f in faults]


def _make_evaluate_value(stock):
    def evaluate_value(self, pos, ignore_regions=False, fillnan=None):
        """GeologicalFeature.evaluate_value (LS/modelling/features/_geological_feature.py:116-157) of a FAULTED feature
        in one device call (lsb200_faulted_eval: fault chain, then the field at the moved points, which never come back
        to the host) when the feature and the frames of all its faults were solved by the engine; the stock method in
        every other case (no faults, fillnan, other feature types)."""
        if fillnan is not None or type(self).__name__ != "GeologicalFeature" or np.ndim(pos) != 2 or np.shape(pos)[1] != 3:
            return stock(self, pos, ignore_regions=ignore_regions, fillnan=fillnan)
        if not (getattr(self, "faults_enabled", True) and len(getattr(self, "faults", [])) > 0):
            return stock(self, pos, ignore_regions=ignore_regions, fillnan=fillnan)
        src = _solved_grid_field(self, kinds=("grid", "tetmesh"), allow_faults=True)
        chain = _fault_chain_of(self) if src is not None else None
        if chain is None:
            return stock(self, pos, ignore_regions=ignore_regions, fillnan=fillnan)
        from . import fault as fl

        pts = np.array(pos, dtype=float)
        mask = self._calculate_mask(pos, ignore_regions=ignore_regions)
        v = np.full(pts.shape[0], np.nan)
        if mask.dtype not in [int, bool]:
            return v  # the reference logs an error and returns NaN (:146-147)
        val = fl.faulted_evaluate(*chain, src[0], pts, gradient=False, world_to_local=src[1])
        keep = mask.astype(bool)
        v[keep] = val[keep]
        return v

    return evaluate_value


def _make_apply_faults(stock):
    def _apply_faults(self, evaluation_points, reverse=False):
        """BaseFeature._apply_faults (LS/modelling/features/_base_geological_feature.py:252-270): two or more faults
        whose frames were all solved by the engine are applied in one call (lsb200_fault_displace_chain: the points
        cross the bus once); anything else goes fault by fault through FaultSegment.apply_to_points as before."""
        from . import fault as fl

        if not getattr(self, "faults_enabled", True) or len(getattr(self, "faults", [])) < 2:
            return stock(self, evaluation_points, reverse=reverse)
        chain = _fault_chain_of(self)
        if chain is None:
            return stock(self, evaluation_points, reverse=reverse)
        return fl.apply_fault_chain(*chain, np.array(evaluation_points, dtype=float))

    return _apply_faults


def _install_feature_patch():
    try:
        from LoopStructural.modelling.features._base_geological_feature import BaseFeature
    except Exception:
        return
    if "feature_stock" not in _installed:
        _installed["feature_stock"] = (BaseFeature, BaseFeature.scalar_field)
        _installed["apply_faults_stock"] = BaseFeature._apply_faults
        _installed["gradient_norm_stock"] = BaseFeature.gradient_norm_scalar_field
    cls, stock = _installed["feature_stock"]
    cls.scalar_field = _make_scalar_field(stock)
    cls.gradient_norm_scalar_field = _make_gradient_norm_scalar_field(_installed["gradient_norm_stock"])
    cls._apply_faults = _make_apply_faults(_installed["apply_faults_stock"])
    try:
        from LoopStructural.modelling.features._geological_feature import GeologicalFeature
    except Exception:
        return
    if "evaluate_value_stock" not in _installed:
        _installed["evaluate_value_stock"] = (GeologicalFeature, GeologicalFeature.evaluate_value)
    gcls, gstock = _installed["evaluate_value_stock"]
    gcls.evaluate_value = _make_evaluate_value(gstock)


def _install_isosurface_patch():
    """LoopIsosurfacer.fit (LoopStructural/utils/_surface.py:139-147) calls the module-level name ``marching_cubes``
    (skimage.measure.marching_cubes): it is pointed at the device isosurface (loopstructural_b200.isosurface), which takes
    the same arguments and returns the same four arrays.  BaseFeature.surfaces then runs: evaluate_value on the regular
    grid (device, through the patched evaluate_value) -> device isosurface -> Surface objects."""
    try:
        import LoopStructural.utils._surface as mod
    except Exception:
        return
    from .isosurface import marching_tetrahedra

    if "marching_cubes_stock" not in _installed:
        _installed["marching_cubes_stock"] = (mod, mod.marching_cubes)

    def marching_cubes(volume, level=None, *, spacing=(1.0, 1.0, 1.0), mask=None, **kwargs):
        return marching_tetrahedra(volume, level, spacing=spacing, mask=mask)

    mod.marching_cubes = marching_cubes


def _install_fault_patch():
    """Patch FaultSegment when LoopStructural itself is importable (the interpolation package alone is enough for
    everything else in this module)."""
    try:
        from LoopStructural.modelling.features.fault import FaultSegment
    except Exception:
        return
    if "fault_stock" in _installed:
        stock = _installed["fault_stock"]
    else:
        stock = _installed["fault_stock"] = (FaultSegment, FaultSegment.apply_to_points, FaultSegment.evaluate_displacement)
    FaultSegment.apply_to_points, FaultSegment.evaluate_displacement = _make_fault_methods(stock[1], stock[2])


def install():
    """Register the B200 subclasses in the reference's ``interpolator_map`` (idempotent).  Returns the classes."""
    import loop_interpolation as li

    T = li.InterpolatorType
    if "fdi" not in _installed:
        _installed["fdi"] = _make_fdi_class()
        _installed["p1"] = _make_p1_class()
        _installed["dfi"] = _make_dfi_class()
        _installed["stock"] = {T.FINITE_DIFFERENCE: li.interpolator_map[T.FINITE_DIFFERENCE],
                               T.PIECEWISE_LINEAR: li.interpolator_map[T.PIECEWISE_LINEAR],
                               T.DISCRETE_FOLD: li.interpolator_map[T.DISCRETE_FOLD]}
    li.interpolator_map[T.FINITE_DIFFERENCE] = _installed["fdi"]
    li.interpolator_map[T.PIECEWISE_LINEAR] = _installed["p1"]
    li.interpolator_map[T.DISCRETE_FOLD] = _installed["dfi"]
    _install_fault_patch()
    _install_builder_patch()
    _install_feature_patch()
    _install_isosurface_patch()
    return {"FDI": _installed["fdi"], "P1": _installed["p1"], "DFI": _installed["dfi"]}


def uninstall():
    """Put the stock classes back."""
    import loop_interpolation as li

    for key, cls in _installed.get("stock", {}).items():
        li.interpolator_map[key] = cls
    if "fault_stock" in _installed:
        cls, stock_apply, stock_displacement = _installed["fault_stock"]
        cls.apply_to_points, cls.evaluate_displacement = stock_apply, stock_displacement
    if "builder_stock" in _installed:
        cls, stock = _installed["builder_stock"]
        cls.install_gradient_constraint = stock
    if "feature_stock" in _installed:
        cls, stock = _installed["feature_stock"]
        cls.scalar_field = stock
        cls._apply_faults = _installed["apply_faults_stock"]
        cls.gradient_norm_scalar_field = _installed["gradient_norm_stock"]
    if "evaluate_value_stock" in _installed:
        gcls, gstock = _installed["evaluate_value_stock"]
        gcls.evaluate_value = gstock
    if "marching_cubes_stock" in _installed:
        mod, stock = _installed["marching_cubes_stock"]
        mod.marching_cubes = stock
