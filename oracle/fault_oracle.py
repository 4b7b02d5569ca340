"""CPU restatement of the fault-frame evaluation that consumes the solved fields (SURVEY.md section 8 f, rank 1).

TEST INFRASTRUCTURE ONLY: imported by tests/ (and the bench's cpu legs); never by the product path.

Restates, in numpy and in the reference's operation order,
  * FaultSegment.apply_to_points       LS/modelling/features/fault/_fault_segment.py:315-423
  * FaultSegment.evaluate_displacement LS/modelling/features/fault/_fault_segment.py:283-313
  * the fault profile functions        LS/modelling/features/fault/_fault_function.py:12-16 (smooth_peak),
                                       :40-140 (CubicFunction), :198-226 (Composite), :256-305 (Ones / Zeros),
                                       :308-369 (FaultDisplacement)
  * the world <-> local transforms     LC/geometry/_bounding_box.py:118-176,702-761 behind
                                       GeologicalInterpolator.evaluate_value / evaluate_gradient
                                       (LI/_geological_interpolator.py:723-741)
for fault-frame features that carry no faults and no regions of their own (GeologicalFeature.evaluate_value /
evaluate_gradient then reduce to the interpolator calls, LS/modelling/features/_geological_feature.py:116-237).

Parity is pinned: tests/golden/fault_*.npz hold inputs and outputs of the unmodified reference
(tests/golden/make_golden_fault.py), tests/test_oracle_golden_fault.py checks this module against them.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import fdi_oracle as fo

ZEROS, ONES, CUBIC, SMOOTH_PEAK = 0, 1, 2, 3


@dataclass
class Profile1D:
    """One 1-D profile.  kind CUBIC carries CubicFunction's w, lim, min_v, max_v."""

    kind: int = ONES
    w: tuple = (0.0, 0.0, 0.0, 0.0)
    lim: tuple = (-1.0, 1.0)
    min_v: float = -99999.0
    max_v: float = 999999.0

    def __call__(self, v: np.ndarray) -> np.ndarray:
        v = np.asarray(v, dtype=float)
        if self.kind == ZEROS:
            return np.zeros(v.shape)
        if self.kind == ONES:
            return np.ones(v.shape)
        if self.kind == SMOOTH_PEAK:  # _fault_function.py:12-16
            out = np.zeros(v.shape)
            m = np.logical_and(v >= -1, v <= 1)
            out[m] = v[m] ** 4 - 2 * v[m] ** 2 + 1
            return out
        w, lim = self.w, self.lim  # CubicFunction.__call__, _fault_function.py:122-140
        eva = w[0] * v**3 + w[1] * v**2 + w[2] * v + w[3]
        eva[v > lim[1]] = w[0] * lim[1] ** 3 + w[1] * lim[1] ** 2 + w[2] * lim[1] + w[3]
        eva[v < lim[0]] = w[0] * lim[0] ** 3 + w[1] * lim[0] ** 2 + w[2] * lim[0] + w[3]
        eva[eva > self.max_v] = self.max_v
        eva[eva < self.min_v] = self.min_v
        return eva


@dataclass
class AxisProfile:
    """Profile along one fault-frame coordinate: Composite(positive, negative) (0 at v == 0) or one function."""

    pos: Profile1D = field(default_factory=Profile1D)
    neg: Profile1D | None = None  # None: `pos` for every v

    def __call__(self, v: np.ndarray) -> np.ndarray:
        v = np.asarray(v, dtype=float)
        if self.neg is None:
            return self.pos(v)
        r = np.zeros(v.shape)  # Composite.__call__, _fault_function.py:211-226
        r[v > 0] = self.pos(v[v > 0])
        r[v < 0] = self.neg(v[v < 0])
        return r


@dataclass
class FaultProfile:
    """FaultDisplacement (scale * gx(gx) * gy(gy) * gz(gz), _fault_function.py:366-368); `axes` None = the
    segment has no fault function (d = [gx > 0])."""

    axes: tuple | None = None
    scale: float = 0.5

    def __call__(self, gx, gy, gz):
        return self.scale * self.axes[0](gx) * self.axes[1](gy) * self.axes[2](gz)


@dataclass
class FrameField:
    """One solved scalar field of the fault frame: its support, coefficients and world -> local transform."""

    grid: fo.Grid
    c: np.ndarray
    world_to_local: np.ndarray | None = None  # 4 x 4; None = points are already local
    kind: str = "grid"  # "grid": StructuredGrid (FDI); "tetmesh": TetMesh on the same lattice (P1 / fold interpolators)

    def project(self, pts):
        if self.world_to_local is None:
            return np.asarray(pts, dtype=float)
        hom = np.ones((pts.shape[0], 4))  # BoundingBox._apply_affine, _bounding_box.py:154-176
        hom[:, :3] = pts
        return (self.world_to_local @ hom.T).T[:, :3]

    def value(self, pts):
        loc = self.project(pts)
        if self.kind == "tetmesh":
            from . import p1_oracle as po

            out = np.zeros(loc.shape[0])  # rows with NaN coordinates: 0.0 (LI/_discrete_interpolator.py:1357-1361)
            good = ~np.any(np.isnan(loc), axis=1)
            if good.any():
                out[good] = po.evaluate_value(self.grid, loc[good], self.c)
            return out
        return fo.evaluate_value(self.grid, loc, self.c)

    def gradient(self, pts):
        if self.kind == "tetmesh":
            from . import p1_oracle as po

            g = po.evaluate_gradient(self.grid, self.project(pts), self.c)  # NaN outside the mesh
        else:
            g = fo.evaluate_gradient(self.grid, self.project(pts), self.c)
        if self.world_to_local is None:
            return g
        rot = np.linalg.inv(self.world_to_local)[:3, :3]  # reproject_vectors, _bounding_box.py:750-761
        return (rot @ g.T).T


def _step_displacement(fields, profile, gx, gy, gz):
    """d of one integration step (before `* displacement`), _fault_segment.py:389-403."""
    d = np.zeros(gx.shape)
    ok = ~np.isnan(gx) & ~np.isnan(gy) & ~np.isnan(gz)
    pos = np.zeros_like(ok)
    pos[ok] = gx[ok] > 0
    d[pos] = 1.0
    if profile.axes is not None:
        d[ok] = profile(gx[ok], gy[ok], gz[ok])
    return d, ok


def apply_to_points(fields, profile: FaultProfile, displacement: float, steps: int, points: np.ndarray) -> np.ndarray:
    """FaultSegment.apply_to_points (_fault_segment.py:315-423).  `reverse` only scales a value the reference
    discards before the loop (:360-363 vs :403), so it does not appear here."""
    newp = np.copy(points).astype(float)
    gx, gy, gz = (f.value(newp) for f in fields)
    d, _ = _step_displacement(fields, profile, gx, gy, gz)
    mask = np.abs(d) > 0.0
    for i in range(steps):
        if i == 0:
            gx_m, gy_m, gz_m = gx[mask], gy[mask], gz[mask]
        else:
            gx_m, gy_m, gz_m = (f.value(newp[mask, :]) for f in fields)
        g = fields[1].gradient(newp[mask, :])
        d, ok = _step_displacement(fields, profile, gx_m, gy_m, gz_m)
        d *= displacement
        g_mag = np.zeros(g.shape[0])
        g_mag[ok] = np.linalg.norm(g[ok], axis=1)
        g[g_mag > 0.0] /= g_mag[g_mag > 0, None]
        g *= (1.0 / steps) * d[:, None]
        newp[mask, :] += g
    return newp


def evaluate_displacement(fields, profile: FaultProfile, displacement: float, fault_offset: float,
                          points: np.ndarray) -> np.ndarray:
    """FaultSegment.evaluate_displacement (_fault_segment.py:283-313)."""
    newp = np.copy(points).astype(float)
    gx, gy, gz = (f.value(newp) for f in fields)
    d = np.zeros(gx.shape)
    ok = ~np.isnan(gx) & ~np.isnan(gy) & ~np.isnan(gz)
    pos = np.zeros_like(ok)
    pos[ok] = gx[ok] > 0
    d[pos] = 1.0
    if profile.axes is not None:
        d[ok] = profile(gx[ok] + fault_offset, gy[ok], gz[ok])
    return d * displacement


def frame_eval(fields, points: np.ndarray, gradient_of: int = 1):
    """The gathers one step needs: the three values and the gradient of one field at the same points."""
    vals = np.stack([f.value(points) for f in fields], axis=1)
    return vals, fields[gradient_of].gradient(points)


# ---------------------------------------------------------------------------------------------------------------
# Orthogonality rows of a structural frame (SURVEY.md section 8 f, rank 2)
# ---------------------------------------------------------------------------------------------------------------


def barycentres(grid: fo.Grid, elements: np.ndarray) -> np.ndarray:
    """StructuredGrid.cell_centres (LC/supports/_3d_structured_grid.py:70-88): origin + step * cell + step * 0.5, cell
    index = cx + ncx * (cy + ncy * cz) (global_index_to_cell_index, LC/supports/_3d_base_structured.py)."""
    nc = grid.ncells
    cx = elements % nc[0]
    cy = (elements // nc[0]) % nc[1]
    cz = elements // (nc[0] * nc[1])
    cell = np.stack([cx, cy, cz], axis=1)
    return grid.origin[None, :] + grid.step[None, :] * cell + grid.step[None, :] * 0.5


def field_orthogonal_rows(grid: fo.Grid, other: FrameField, element_step: int, w: float, b: float = 0.0,
                          name: str = "orthogonal"):
    """GeologicalFeatureBuilder.install_gradient_constraint
    (LS/modelling/features/builders/_geological_feature_builder.py:293-318) for a discrete interpolator on a
    StructuredGrid: gradient of the OTHER feature at every barycentre of this support, normalised, every
    `element_step`-th element -> FDI.add_gradient_orthogonal_constraints (LI/_finite_difference_interpolator.py:478-539).
    The barycentres are this support's LOCAL coordinates and the reference hands them to the other feature's
    evaluate_gradient, which projects them as if they were world coordinates: restated as is."""
    n_el = int(np.prod(grid.ncells))
    bary = barycentres(grid, np.arange(n_el))
    vec = other.gradient(bary)
    norm = np.linalg.norm(vec, axis=1)
    vec[norm > 0] /= norm[norm > 0, None]
    sel = np.arange(n_el)[::element_step]
    if np.isnan(vec[sel]).any():  # a tet-mesh source, barycentre outside it: the reference drops the whole family
        return None
    return fo.orthogonal_rows(grid, bary[sel], vec[sel], w=w, b=b, name=name)


def p1_field_orthogonal_rows(grid: fo.Grid, other: FrameField, element_step: int, w: float, b: float = 0.0,
                             name: str = "orthogonal"):
    """The same builder step for a P1 interpolator on a TetMesh: TetMesh.barycentre (mean of the four vertices,
    LC/supports/_3d_structured_tetra.py:106-118) of every `element_step`-th tet and the other feature's normalised gradient
    there -> P1Interpolator.add_gradient_orthogonal_constraints (LI/_p1interpolator.py:248-296; weight w * tet volume).
    Returns None when a selected vector holds a NaN (a barycentre outside a tet-mesh source): the reference then drops the
    whole family (LI/_discrete_interpolator.py:357-359)."""
    from . import p1_oracle as po

    el = po.elements(grid)
    bary = np.sum(po.node_xyz(grid, el), axis=1) / 4.0
    vec = other.gradient(bary)
    norm = np.linalg.norm(vec, axis=1)
    vec[norm > 0] /= norm[norm > 0, None]
    sel = np.arange(el.shape[0])[::element_step]
    if np.isnan(vec[sel]).any():
        return None
    return po.orthogonal_rows(grid, bary[sel], vec[sel], w=w, b=b, name=name)


# ---------------------------------------------------------------------------------------------------------------
# Regular-grid export (SURVEY.md section 8 f, rank 4)
# ---------------------------------------------------------------------------------------------------------------


def lattice_points(start, stop, n, order: str = "F") -> np.ndarray:
    """BoundingBox.regular_grid(local=False) (LC/geometry/_bounding_box.py:532-573): np.linspace per axis, meshgrid
    'ij', flattened in `order`."""
    coords = [np.linspace(start[i], stop[i], int(n[i])) for i in range(3)]
    grid = np.meshgrid(*coords, indexing="ij")
    return np.array([g.flatten(order=order) for g in grid]).T


def evaluate_lattice(fld: FrameField, start, stop, n, order: str = "F", gradient: bool = False) -> np.ndarray:
    """What BaseFeature.scalar_field stores in grid.properties (LS/modelling/features/_base_geological_feature.py:359-388)
    for a feature without faults and regions: evaluate_value / evaluate_gradient on the lattice points."""
    pts = lattice_points(start, stop, n, order)
    return fld.gradient(pts) if gradient else fld.value(pts)
