"""Host-side mirror of the reference's DiscreteInterpolator API for the FDI / P1 hot path.

``B200FiniteDifferenceInterpolator`` and ``B200P1Interpolator`` have the method names, argument meaning and
error behaviour of ``loop_interpolation.FiniteDifferenceInterpolator`` / ``P1Interpolator``
(LI/_finite_difference_interpolator.py:91-224, LI/_p1interpolator.py:23-113, LI/_discrete_interpolator.py:
1221-1379, LI/_geological_interpolator.py:249-433,723-741) but hold no numerics: every call ends in the C-ABI of
``include/lsb200.h``.  They need neither the reference nor scipy, so they are what the GPU parity tests, the
benchmark's end-to-end arm and ``__graft_entry__.smoke()`` drive; ``plugin.py`` puts the same engine behind the
reference's own classes.

Coordinates are the support's LOCAL coordinates, as they are inside the reference once
``GeologicalInterpolator.set_*_constraints`` has projected them (LI/_geological_interpolator.py:273-286).
"""
from __future__ import annotations

import time

import numpy as np

from . import engine as en
from .engine import Engine, FDI_DEFAULT_WEIGHTS

B200_SOLVERS = ("b200", "b200-cg")
# P1 weight defaults, LI/_p1interpolator.py:40-47
P1_DEFAULT_WEIGHTS = dict(cgw=0.1, cpw=1.0, npw=1.0, gpw=1.0, tpw=1.0, ipw=1.0)


class Support:
    """The part of StructuredGrid / TetMesh the hot path needs: origin, step vector and CELL counts
    (``nsteps_cells``; node counts are one more, LC/supports/_3d_base_structured.py:66,121-123)."""

    def __init__(self, origin, nsteps_cells, step_vector):
        self.origin = np.asarray(origin, dtype=np.float64).reshape(3).copy()
        self.nsteps_cells = np.asarray(nsteps_cells, dtype=np.int64).reshape(3).copy()
        self.step_vector = np.asarray(step_vector, dtype=np.float64).reshape(3).copy()
        if np.any(self.nsteps_cells < 1):
            raise ValueError("nsteps_cells must be at least 1 on every axis")
        if np.any(self.step_vector <= 0):
            raise ValueError("step_vector must be positive")

    @property
    def nsteps(self) -> np.ndarray:
        return self.nsteps_cells + 1

    @property
    def n_nodes(self) -> int:
        return int(np.prod(self.nsteps))

    @property
    def maximum(self) -> np.ndarray:
        return self.origin + self.nsteps_cells * self.step_vector


class _B200DiscreteInterpolator:
    """Shared driver: data setters, solve_system, evaluate_*.  Subclasses say which support kind they
    need and how setup_interpolator turns the data into rows."""

    kind = "grid"
    default_weights: dict = {}

    def __init__(self, support: Support, data: dict | None = None, device: int = 0):
        self.support = support
        self.dof = support.n_nodes
        self.device = int(device)
        self.interpolation_weights = dict(self.default_weights)
        # DiscreteInterpolator.__init__ attributes, LI/_discrete_interpolator.py:59,76-78
        self.ridge_factor = 1e-8
        self.add_ridge_regulatisation = True  # (sic) the reference's spelling
        self.apply_scaling_matrix = True
        self.solver = "b200"
        self.data = {k: np.zeros((0, n)) for k, n in
                     (("value", 5), ("gradient", 7), ("normal", 7), ("tangent", 7), ("interface", 5))}
        if data:
            for k, v in data.items():
                self._set(k, v)
        self.c = np.full(self.dof, np.nan)
        self.up_to_date = False
        self.latest_solve_timing: dict = {}
        self.row_counts: dict = {}
        self._engine: Engine | None = None
        self._c_on_device = False

    # -- data (GeologicalInterpolator.set_*_constraints, LI/_geological_interpolator.py:325-433) ----------
    def _set(self, key, points):
        points = np.asarray(points, dtype=np.float64)
        need = 5 if key in ("value", "interface") else 7
        if points.ndim != 2 or points.shape[1] < need:
            raise ValueError(f"{key} points array too small, needs {need} columns")
        self.data[key] = np.ascontiguousarray(points[:, :need])
        self.up_to_date = False

    def set_value_constraints(self, points):
        self._set("value", points)

    def set_gradient_constraints(self, points):
        self._set("gradient", points)

    def set_normal_constraints(self, points):
        self._set("normal", points)

    def set_tangent_constraints(self, points):
        self._set("tangent", points)

    def set_interface_constraints(self, points):
        self._set("interface", points)

    # -- engine -------------------------------------------------------------------------------------------
    @property
    def engine(self) -> Engine:
        if self._engine is None:
            self._engine = Engine(self.support.nsteps, self.support.origin, self.support.step_vector, kind=self.kind,
                                  device=self.device)
        return self._engine

    def close(self):
        if self._engine is not None:
            self._c_on_device = False  # self.c stays valid: it owns its (page-locked) block
            self._engine.close()
            self._engine = None

    def _inside(self, xyz: np.ndarray) -> np.ndarray:
        """StructuredGrid.inside with its 32-eps tolerance (LC/supports/_3d_base_structured.py:287-293); TetMesh.inside is
        strict on top of it (LC/supports/_3d_structured_tetra.py:264-272)."""
        lo, hi = self.support.origin, self.support.maximum
        with np.errstate(invalid="ignore"):
            if self.kind == "tetmesh":
                return np.all((xyz > lo) & (xyz < hi), axis=1)
            tol = np.finfo(float).eps * 32.0 * np.maximum(np.abs(hi - lo), 1.0)
            return np.all((xyz >= lo - tol) & (xyz <= hi + tol), axis=1)

    def _family(self, key: str, used_cols) -> np.ndarray | None:
        """Constraint points of one family, or None when the reference would add nothing: it drops the WHOLE family when
        any row that reaches add_constraints_to_least_squares holds a NaN (LI/_discrete_interpolator.py:355-359); points
        outside the support never get that far.  (The C-ABI itself drops single rows.)"""
        pts = self.data[key]
        if len(pts) == 0:
            return None
        inside = self._inside(pts[:, :3])
        if inside.any() and np.isnan(pts[inside][:, used_cols]).any():
            import warnings

            warnings.warn(f"Constraints contain nan not adding constraints: {key}", stacklevel=3)
            return None
        return pts

    def _apply_weight_kwargs(self, kwargs):
        # every unknown key lands in interpolation_weights (LI/_discrete_interpolator.py:172-177)
        for k, v in kwargs.items():
            if k in ("regularisation", "solver", "tol", "solver_kwargs") or v is None:
                continue
            self.interpolation_weights[k] = v

    # -- solve (DiscreteInterpolator.solve_system, LI/_discrete_interpolator.py:1221-1319) -----------------
    def solve_system(self, solver=None, tol: float | None = None, solver_kwargs: dict | None = None) -> bool:
        """``solver`` must be one of B200_SOLVERS (or None = the instance default): this class has no CPU
        path to fall back to.  ``tol`` becomes ``atol`` unless ``solver_kwargs`` holds both ``atol`` and
        ``rtol`` (LI/_solver_strategy.py:59-60); ``maxiter`` and ``x0`` pass through as for scipy's cg."""
        solver = self.solver if solver is None else solver
        if solver not in B200_SOLVERS:
            raise ValueError(f"solver {solver!r} is not served by the B200 engine (use one of {B200_SOLVERS})")
        if self._engine is None or not self.row_counts:
            raise ValueError("Pre solve failed")  # the reference's message when nothing was set up (:1245-1246)
        kw = dict(solver_kwargs or {})
        rtol = float(kw.get("rtol", 1e-5))
        atol = float(kw.get("atol", 0.0))
        if tol is not None and not ("atol" in kw and "rtol" in kw):
            atol = float(tol)
        t0 = time.perf_counter()
        e = self.engine
        e.set_options(ridge_factor=self.ridge_factor, add_ridge=self.add_ridge_regulatisation,
                      apply_scaling=self.apply_scaling_matrix, deterministic=bool(kw.get("deterministic", False)))
        # The result lands in a page-locked array that owns its block (Engine.pinned_vector).  Like the reference's
        # solve_system this rebinds self.c; an array the caller still holds from an earlier solve keeps its data (the
        # engine reuses a block only when nobody else refers to it) and outlives close().
        x0 = kw.get("x0")
        x0 = None if x0 is None else np.array(x0, dtype=np.float64)
        self.c = None  # drop our reference to the previous result so that its block can be reused
        try:
            c, res = e.solve_cg(rtol=rtol, atol=atol, maxiter=kw.get("maxiter"), x0=x0, out=e.pinned_vector())
        except Exception:
            self.c = np.full(self.dof, np.nan)
            self._c_on_device = False
            raise
        self.c = c
        self._c_on_device = True
        self.up_to_date = True  # non-convergence is a warning, never a failure (LI/_solver_strategy.py:66-71)
        self.latest_solve_timing = {
            "backend": "b200", "solver": solver, "iterations": res.iterations, "info": res.info,
            "residual_norm": res.residual_norm, "threshold": res.threshold,
            "preprocess_seconds": res.setup_seconds, "solve_seconds": res.solve_seconds,
            "total_seconds": time.perf_counter() - t0, "kernel_launches": res.kernel_launches,
        }
        return True

    def get_last_solve_timing(self) -> dict:
        return dict(self.latest_solve_timing)

    def update(self) -> bool:
        """DiscreteInterpolator.update (LI/_discrete_interpolator.py:1321-1340)."""
        if self.up_to_date:
            return True
        self.setup_interpolator()
        return self.solve_system(self.solver)

    # -- evaluation (LI/_discrete_interpolator.py:1342-1379) -------------------------------------------------
    def _sync_c(self):
        if not self._c_on_device:
            c = np.asarray(self.c, dtype=np.float64)
            if c.shape != (self.dof,):
                raise ValueError(f"c must have shape ({self.dof},)")
            self.engine.set_c(c)
            self._c_on_device = True

    def set_c(self, c):
        """Assign node values computed elsewhere (the reference lets callers write ``interpolator.c``)."""
        self.c = np.asarray(c, dtype=np.float64).copy()
        self._c_on_device = False

    def evaluate_value(self, locations) -> np.ndarray:
        locations = np.asarray(locations, dtype=np.float64)
        if locations.ndim != 2 or locations.shape[1] != 3:
            raise ValueError("locations must be (n,3)")
        self._sync_c()
        return self.engine.evaluate_value(locations)

    def evaluate_gradient(self, locations) -> np.ndarray:
        locations = np.asarray(locations, dtype=np.float64)
        if locations.ndim != 2 or locations.shape[1] != 3:
            raise ValueError("locations must be (n,3)")
        self._sync_c()
        return self.engine.evaluate_gradient(locations)


class B200FiniteDifferenceInterpolator(_B200DiscreteInterpolator):
    """FiniteDifferenceInterpolator on a StructuredGrid (LI/_finite_difference_interpolator.py:91-224)."""

    kind = "grid"
    default_weights = FDI_DEFAULT_WEIGHTS

    def setup_interpolator(self, **kwargs):
        """Weights as keyword arguments, ``regularisation=`` sets the six second-derivative weights at once
        (LI/_finite_difference_interpolator.py:158-170).  Families are added in the reference's order
        (:205-210): normal, gradient, value, tangent, interface; then the stencil families, which the engine
        applies matrix-free and never stores."""
        reg = kwargs.get("regularisation")
        if reg is not None:
            for k in ("dxy", "dyz", "dxz", "dxx", "dyy", "dzz"):
                self.interpolation_weights[k] = float(reg)
        self._apply_weight_kwargs(kwargs)
        w = self.interpolation_weights
        e = self.engine
        e.clear_rows()
        e.set_regularisation(w)
        rc = {}
        vec = [3, 4, 5]
        normal, gradient, value, tangent = (self._family("normal", vec), self._family("gradient", vec),
                                            self._family("value", [3]), self._family("tangent", vec))
        rc["normal"] = e.add_point_rows(en.ROWS_NORMAL, normal, w["npw"]) if normal is not None else 0
        rc["gradient"] = e.add_point_rows(en.ROWS_GRADIENT, gradient, w["gpw"]) if gradient is not None else 0
        rc["value"] = e.add_point_rows(en.ROWS_VALUE, value, w["cpw"]) if value is not None else 0
        if tangent is not None:
            # add_tangent_constraints -> add_gradient_orthogonal_constraints(points, vectors, w, b=0)
            # (LI/_discrete_interpolator.py:668-682): column 6 of the point array is the row target b
            t = tangent.copy()
            t[:, 6] = 0.0
            rc["tangent"] = e.add_point_rows(en.ROWS_TANGENT, t, w["tpw"])
        else:
            rc["tangent"] = 0
        if len(self.data["interface"]):
            rc["interface"] = e.add_interface_rows(self.data["interface"], w["ipw"])
        else:
            rc["interface"] = 0
        rc["regularisation"] = "matrix-free"
        self.row_counts = rc
        self.up_to_date = False
        return rc

    def add_gradient_orthogonal_constraints(self, points, vectors, w=1.0, b=0.0, name="gradient orthogonal"):
        """FDI.add_gradient_orthogonal_constraints (LI/_finite_difference_interpolator.py:478-539): one row
        v_hat . grad phi = b per point, used for tangents and structural-frame orthogonality."""
        points = np.asarray(points, dtype=np.float64)
        vectors = np.asarray(vectors, dtype=np.float64)
        if points.shape[0] == 0:
            return 0
        data = np.column_stack([points[:, :3], vectors[:, :3], np.broadcast_to(np.asarray(b, float), (points.shape[0],))])
        n = self.engine.add_point_rows(en.ROWS_TANGENT, data, float(w))
        self.row_counts[name] = self.row_counts.get(name, 0) + n
        self.up_to_date = False
        return n

    def minimise_directional_gradient_change(self, w, vector, name="directional regularisation"):
        """FDI.minimise_directional_gradient_change (LI/_finite_difference_interpolator.py:1207-1340): the six
        direction-weighted second-derivative families from the (n_nodes, 3) direction field at the nodes; a field of
        another shape is skipped with the reference's warning semantics (returns 0)."""
        vector = None if vector is None else np.asarray(vector, dtype=np.float64)
        if vector is None or vector.ndim != 2 or vector.shape != (int(np.prod(self.engine.nsteps)), 3):
            return 0
        n = self.engine.add_directional_rows(vector, float(w))
        self.row_counts[name] = self.row_counts.get(name, 0) + n
        self.up_to_date = False
        return n


class B200P1Interpolator(_B200DiscreteInterpolator):
    """P1Interpolator on a TetMesh (LI/_p1interpolator.py:23-113)."""

    kind = "tetmesh"
    default_weights = P1_DEFAULT_WEIGHTS

    def setup_interpolator(self, **kwargs):
        """LI/_p1interpolator.py:198-246: ``regularisation=`` overrides cgw; families in the reference's order:
        face gradient jumps (matrix-free here), normal, value, tangent.  P1 ignores gradient constraints
        (add_gradient_constraints is a no-op, :70-71)."""
        reg = kwargs.get("regularisation")
        if reg is not None:
            self.interpolation_weights["cgw"] = float(reg)
        self._apply_weight_kwargs(kwargs)
        w = self.interpolation_weights
        e = self.engine
        e.clear_rows()
        e.set_p1_regularisation(w["cgw"])
        rc = {"regularisation": "matrix-free"}
        normal, value, tangent = self._family("normal", [3, 4, 5]), self._family("value", [3]), self._family("tangent", [3, 4, 5])
        rc["normal"] = e.add_point_rows(en.ROWS_NORMAL, normal, w["npw"]) if normal is not None else 0
        rc["value"] = e.add_point_rows(en.ROWS_VALUE, value, w["cpw"]) if value is not None else 0
        if tangent is not None:
            t = tangent.copy()
            t[:, 6] = 0.0
            rc["tangent"] = e.add_point_rows(en.ROWS_TANGENT, t, w["tpw"])
        else:
            rc["tangent"] = 0
        self.row_counts = rc
        self.up_to_date = False
        return rc
