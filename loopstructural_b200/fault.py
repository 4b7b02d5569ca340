"""Host side of the fault-frame evaluation (csrc/frame.cuh; SURVEY.md section 8 f, rank 1).

``FaultFrame`` binds three solved ``Engine`` handles -- the fields of a structural frame, coordinates 0 / 1 / 2 --
and offers the calls FaultSegment makes on them (LoopStructural/modelling/features/fault/_fault_segment.py):
``apply_to_points`` (:315-423), ``evaluate_displacement`` (:283-313) and the fused gather ``evaluate`` behind both.
It holds no numerics: every method is one call through the C-ABI (``lsb200_fault_displace`` ...).

``profile_from_reference`` translates the reference's profile objects
(LoopStructural/modelling/features/fault/_fault_function.py) into the plain ``lsb200_fault_profile`` struct by
duck-typing, without importing the reference; anything it does not recognise returns ``None`` and the caller must
leave the work to the reference.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nat

PROFILE_ZEROS, PROFILE_ONES, PROFILE_CUBIC, PROFILE_SMOOTH_PEAK = 0, 1, 2, 3


def profile1d(kind: int = PROFILE_ONES, w=(0.0, 0.0, 0.0, 0.0), lim=(-1.0, 1.0), min_v: float = -99999.0,
              max_v: float = 999999.0) -> nat.Profile1D:
    """One 1-D profile; the defaults of min_v / max_v / lim are CubicFunction's (_fault_function.py:21,52-53)."""
    p = nat.Profile1D()
    p.kind = int(kind)
    p.w[:] = [float(x) for x in w]
    p.lim[:] = [float(lim[0]), float(lim[1])]
    p.min_v, p.max_v = float(min_v), float(max_v)
    return p


def fault_profile(axes=None, scale: float = 0.5) -> nat.FaultProfile:
    """``axes``: three entries, each a Profile1D (used for every v) or a (positive, negative) pair (Composite);
    ``None`` = the segment has no fault function (d = [gx > 0])."""
    fp = nat.FaultProfile()
    fp.has_function = 0 if axes is None else 1
    fp.scale = float(scale)
    if axes is not None:
        if len(axes) != 3:
            raise ValueError("a fault profile has three axes (gx, gy, gz)")
        for a, ax in enumerate(axes):
            if isinstance(ax, (tuple, list)):
                fp.composite[a] = 1
                fp.pos[a], fp.neg[a] = ax[0], ax[1]
            else:
                fp.composite[a] = 0
                fp.pos[a] = fp.neg[a] = ax
    return fp


def _one_from_reference(fn):
    name = type(fn).__name__
    if name == "function" and getattr(fn, "__name__", "") == "smooth_peak" and \
            getattr(fn, "__module__", "").endswith("_fault_function"):
        return profile1d(PROFILE_SMOOTH_PEAK)
    if name == "Ones":
        return profile1d(PROFILE_ONES)
    if name == "Zeros":
        return profile1d(PROFILE_ZEROS)
    if name == "CubicFunction":
        fn.solve()  # CubicFunction.__call__ solves lazily (_fault_function.py:122-123)
        return profile1d(PROFILE_CUBIC, np.asarray(fn.w, dtype=float).ravel()[:4], fn.lim, fn.min_v, fn.max_v)
    return None


def profile_from_reference(faultfunction):
    """lsb200_fault_profile of a reference FaultDisplacement (or None = no function); returns None when the
    object contains anything outside the closed family Zeros / Ones / CubicFunction / smooth_peak / Composite."""
    if faultfunction is None:
        return fault_profile(None, 0.0)
    if type(faultfunction).__name__ != "FaultDisplacement":
        return None
    axes = []
    for fn in (faultfunction.gx, faultfunction.gy, faultfunction.gz):
        if type(fn).__name__ == "Composite":
            pair = (_one_from_reference(fn.positive), _one_from_reference(fn.negative))
            if pair[0] is None or pair[1] is None:
                return None
            axes.append(pair)
        else:
            one = _one_from_reference(fn)
            if one is None:
                return None
            axes.append(one)
    return fault_profile(axes, float(faultfunction.scale))


def transform_rows(world_to_local=None, local_to_world=None) -> np.ndarray:
    """The 21 doubles per field the C-ABI takes: upper 3 x 4 of world_to_local, then the 3 x 3 rotation of
    local_to_world (BoundingBox.world_to_local_matrix / local_to_world_matrix, LC/geometry/_bounding_box.py:118-152)."""
    if world_to_local is None:
        w2l = np.eye(4)
    else:
        w2l = np.asarray(world_to_local, dtype=np.float64).reshape(4, 4)
    l2w = np.linalg.inv(w2l) if local_to_world is None else np.asarray(local_to_world, dtype=np.float64).reshape(4, 4)
    return np.concatenate([w2l[:3, :].ravel(), l2w[:3, :3].ravel()])


class FaultFrame:
    """Three solved grid engines (same device) + their world -> local transforms."""

    def __init__(self, engines, world_to_local=None):
        engines = list(engines)
        if len(engines) != 3:
            raise ValueError("a fault frame has three fields")
        self._lib = nat.load()
        self.engines = engines
        self._handles = (nat.Handle * 3)(*[e._h for e in engines])
        if world_to_local is None:
            self._transforms = None
        else:
            if len(world_to_local) != 3:
                raise ValueError("one world_to_local matrix per field")
            self._transforms = np.ascontiguousarray(np.concatenate([transform_rows(m) for m in world_to_local]))

    def _tp(self):
        return None if self._transforms is None else nat.dptr(self._transforms)

    @staticmethod
    def _points(xyz):
        xyz = nat.as_f64(xyz)
        if xyz.ndim != 2 or xyz.shape[1] != 3:
            raise ValueError("xyz must be (n,3)")
        return xyz

    def evaluate(self, xyz, gradient_of: int | None = 1):
        """values (n,3) of the three fields and (unless gradient_of is None) the world gradient (n,3) of one."""
        xyz = self._points(xyz)
        n = xyz.shape[0]
        vals = np.empty((n, 3))
        grad = None if gradient_of is None else np.empty((n, 3))
        nat.check(self._lib.lsb200_frame_eval(self._handles, self._tp(), -1 if gradient_of is None else int(gradient_of), n,
                                              nat.dptr(xyz), nat.dptr(vals), None if grad is None else nat.dptr(grad)))
        return vals, grad

    def apply_to_points(self, xyz, profile: nat.FaultProfile, displacement: float, steps: int) -> np.ndarray:
        xyz = self._points(xyz)
        out = np.empty_like(xyz)
        nat.check(self._lib.lsb200_fault_displace(self._handles, self._tp(), C.byref(profile), float(displacement),
                                                  int(steps), xyz.shape[0], nat.dptr(xyz), nat.dptr(out)))
        return out

    def evaluate_displacement(self, xyz, profile: nat.FaultProfile, displacement: float, fault_offset: float = 0.0):
        xyz = self._points(xyz)
        out = np.empty(xyz.shape[0])
        nat.check(self._lib.lsb200_fault_displacement(self._handles, self._tp(), C.byref(profile), float(displacement),
                                                      float(fault_offset), xyz.shape[0], nat.dptr(xyz), nat.dptr(out)))
        return out

    # device pointers (e.g. tensor.data_ptr()), asynchronous on `stream`
    def apply_to_points_device(self, xyz_ptr: int, out_ptr: int, n: int, profile: nat.FaultProfile, displacement: float,
                               steps: int, stream: int = 0):
        nat.check(self._lib.lsb200_fault_displace_dev(self._handles, self._tp(), C.byref(profile), float(displacement),
                                                      int(steps), int(n), C.c_void_p(xyz_ptr), C.c_void_p(out_ptr),
                                                      C.c_void_p(stream)))

    def evaluate_device(self, xyz_ptr: int, values_ptr: int, gradient_ptr: int | None, n: int, gradient_of: int = 1,
                        stream: int = 0):
        nat.check(self._lib.lsb200_frame_eval_dev(self._handles, self._tp(), int(gradient_of), int(n), C.c_void_p(xyz_ptr),
                                                  C.c_void_p(values_ptr), C.c_void_p(gradient_ptr or 0), C.c_void_p(stream)))

    def evaluate_displacement_device(self, xyz_ptr: int, out_ptr: int, n: int, profile: nat.FaultProfile,
                                     displacement: float, fault_offset: float = 0.0, stream: int = 0):
        nat.check(self._lib.lsb200_fault_displacement_dev(self._handles, self._tp(), C.byref(profile), float(displacement),
                                                          float(fault_offset), int(n), C.c_void_p(xyz_ptr),
                                                          C.c_void_p(out_ptr), C.c_void_p(stream)))


def _chain_args(frames, profiles, displacements, steps):
    nf = len(frames)
    if not (len(profiles) == len(displacements) == len(steps) == nf):
        raise ValueError("one profile, displacement and step count per fault")
    handles = (nat.Handle * max(3 * nf, 1))(*[h for fr in frames for h in fr._handles])
    if any(fr._transforms is not None for fr in frames):
        tr = np.ascontiguousarray(np.concatenate([fr._transforms if fr._transforms is not None else
                                                  np.concatenate([transform_rows(None)] * 3) for fr in frames]))
    else:
        tr = None
    profs = (nat.FaultProfile * max(nf, 1))(*profiles)
    disp = np.ascontiguousarray(list(displacements) or [0.0], dtype=np.float64)
    st = (C.c_int * max(nf, 1))(*[int(s) for s in steps])
    return nf, handles, tr, profs, disp, st


def faulted_evaluate(frames, profiles, displacements, steps, engine, xyz, gradient: bool = False, world_to_local=None):
    """GeologicalFeature.evaluate_value / evaluate_gradient of a faulted feature in one call: the fault chain, then
    ``engine``'s field at the moved points, which stay on the device (lsb200_faulted_eval)."""
    frames = list(frames)
    xyz = FaultFrame._points(xyz)
    nf, handles, tr, profs, disp, st = _chain_args(frames, profiles, displacements, steps)
    ftr = None if world_to_local is None else np.ascontiguousarray(transform_rows(world_to_local))
    out = np.empty((xyz.shape[0], 3) if gradient else xyz.shape[0])
    nat.check(nat.load().lsb200_faulted_eval(nf, handles, None if tr is None else nat.dptr(tr), profs, nat.dptr(disp), st, engine._h,
                                             None if ftr is None else nat.dptr(ftr), int(gradient), xyz.shape[0], nat.dptr(xyz),
                                             nat.dptr(out)))
    return out


def apply_fault_chain(frames, profiles, displacements, steps, xyz) -> np.ndarray:
    """Several faults one after the other on the same points (BaseFeature._apply_faults,
    LoopStructural/modelling/features/_base_geological_feature.py:252-270) with one transfer of the points in each
    direction: ``frames`` are FaultFrame objects on one device, in the order the faults are applied."""
    frames = list(frames)
    xyz = FaultFrame._points(xyz)
    if not frames:
        return xyz.copy()
    nf, handles, tr, profs, disp, st = _chain_args(frames, profiles, displacements, steps)
    out = np.empty_like(xyz)
    nat.check(nat.load().lsb200_fault_displace_chain(nf, handles, None if tr is None else nat.dptr(tr), profs, nat.dptr(disp), st,
                                                     xyz.shape[0], nat.dptr(xyz), nat.dptr(out)))
    return out
