"""ctypes binding of the C-ABI in ``include/lsb200.h`` (``liblsb200.so``).

This is the only way the Python host code reaches the device: there is no CPU implementation
behind it.  Loading fails loudly when the library has not been built, and every compute entry
point fails with ``Lsb200Error`` when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblsb200.so")

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)
Handle = C.c_void_p


class Lsb200Error(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"lsb200 error {code}: {message}")
        self.code = code


class SolveInfo(C.Structure):
    _fields_ = [
        ("info", C.c_int32),
        ("iterations", C.c_int64),
        ("residual_norm", C.c_double),
        ("threshold", C.c_double),
        ("rhs_norm", C.c_double),
        ("setup_seconds", C.c_double),
        ("solve_seconds", C.c_double),
        ("matvec_seconds", C.c_double),
        ("kernel_launches", C.c_int64),
    ]


class Profile1D(C.Structure):
    """lsb200_profile1d"""

    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("w", C.c_double * 4), ("lim", C.c_double * 2),
                ("min_v", C.c_double), ("max_v", C.c_double)]


class FaultProfile(C.Structure):
    """lsb200_fault_profile"""

    _fields_ = [("has_function", C.c_int32), ("composite", C.c_int32 * 3), ("scale", C.c_double),
                ("pos", Profile1D * 3), ("neg", Profile1D * 3)]


_Handle3 = C.POINTER(Handle)
_ProfileP = C.POINTER(FaultProfile)

# every symbol include/lsb200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "lsb200_abi_version": (C.c_int, []),
    "lsb200_last_error": (C.c_char_p, []),
    "lsb200_device_count": (C.c_int, []),
    "lsb200_create_grid": (C.c_int, [c_int64_p, c_double_p, c_double_p, C.c_int64, C.c_int64, C.c_int, C.POINTER(Handle)]),
    "lsb200_create_tetmesh": (C.c_int, [c_int64_p, c_double_p, c_double_p, C.c_int, C.POINTER(Handle)]),
    "lsb200_create_tetmesh_slab": (C.c_int, [c_int64_p, c_double_p, c_double_p, C.c_int64, C.c_int64, C.c_int, C.POINTER(Handle)]),
    "lsb200_destroy": (C.c_int, [Handle]),
    "lsb200_local_dof": (C.c_int64, [Handle]),
    "lsb200_support_kind": (C.c_int, [Handle]),
    "lsb200_set_regularisation": (C.c_int, [Handle, c_double_p]),
    "lsb200_set_p1_regularisation": (C.c_int, [Handle, C.c_double]),
    "lsb200_set_p1_tables": (C.c_int, [Handle, c_int64_p, c_double_p]),
    "lsb200_add_rows": (C.c_int, [Handle, C.c_int64, C.c_int, c_double_p, c_int64_p, c_double_p, c_double_p, C.c_int]),
    "lsb200_add_point_rows": (C.c_int, [Handle, C.c_int, C.c_int64, c_double_p, C.c_int, C.c_double, c_int64_p]),
    "lsb200_add_pair_rows": (C.c_int, [Handle, C.c_int64, c_double_p, c_double_p, C.c_double, c_int64_p]),
    "lsb200_add_directional_rows": (C.c_int, [Handle, c_double_p, C.c_double, c_int64_p]),
    "lsb200_add_face_vector_rows": (C.c_int, [Handle, C.c_int64, c_int64_p, c_int64_p, c_double_p, C.c_double, c_int64_p]),
    "lsb200_clear_rows": (C.c_int, [Handle]),
    "lsb200_num_rows": (C.c_int64, [Handle]),
    "lsb200_set_options": (C.c_int, [Handle, C.c_double, C.c_int, C.c_int, C.c_int]),
    "lsb200_finalize_system": (C.c_int, [Handle]),
    "lsb200_comm_unique_id": (C.c_int, [C.POINTER(C.c_ubyte)]),
    "lsb200_comm_init": (C.c_int, [Handle, C.POINTER(C.c_ubyte), C.c_int, C.c_int]),
    "lsb200_solve_cg": (C.c_int, [Handle, C.c_double, C.c_double, C.c_int64, c_double_p, c_double_p, C.POINTER(SolveInfo)]),
    "lsb200_apply_normal": (C.c_int, [Handle, c_double_p, c_double_p, C.c_int]),
    "lsb200_get_rhs": (C.c_int, [Handle, c_double_p]),
    "lsb200_get_diagonal": (C.c_int, [Handle, c_double_p]),
    "lsb200_get_scaling": (C.c_int, [Handle, c_double_p]),
    "lsb200_get_c": (C.c_int, [Handle, c_double_p]),
    "lsb200_set_rhs": (C.c_int, [Handle, c_double_p]),
    "lsb200_comm_uses_peer_memory": (C.c_int, [Handle]),
    "lsb200_residual_norm": (C.c_int, [Handle, c_double_p, c_double_p]),
    "lsb200_set_c": (C.c_int, [Handle, c_double_p]),
    "lsb200_get_c_dev": (C.c_int, [Handle, C.c_void_p, C.c_void_p]),
    "lsb200_set_c_dev": (C.c_int, [Handle, C.c_void_p, C.c_void_p]),
    "lsb200_eval_value": (C.c_int, [Handle, C.c_int64, c_double_p, c_double_p]),
    "lsb200_eval_gradient": (C.c_int, [Handle, C.c_int64, c_double_p, c_double_p]),
    "lsb200_eval_value_dev": (C.c_int, [Handle, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lsb200_eval_gradient_dev": (C.c_int, [Handle, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lsb200_frame_eval": (C.c_int, [_Handle3, c_double_p, C.c_int, C.c_int64, c_double_p, c_double_p, c_double_p]),
    "lsb200_fault_displace": (C.c_int, [_Handle3, c_double_p, _ProfileP, C.c_double, C.c_int, C.c_int64, c_double_p, c_double_p]),
    "lsb200_fault_displace_chain": (C.c_int, [C.c_int, _Handle3, c_double_p, _ProfileP, c_double_p, C.POINTER(C.c_int), C.c_int64,
                                             c_double_p, c_double_p]),
    "lsb200_faulted_eval": (C.c_int, [C.c_int, _Handle3, c_double_p, _ProfileP, c_double_p, C.POINTER(C.c_int), Handle, c_double_p,
                                     C.c_int, C.c_int64, c_double_p, c_double_p]),
    "lsb200_fault_displacement": (C.c_int, [_Handle3, c_double_p, _ProfileP, C.c_double, C.c_double, C.c_int64, c_double_p, c_double_p]),
    "lsb200_frame_eval_dev": (C.c_int, [_Handle3, c_double_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lsb200_fault_displace_dev": (C.c_int, [_Handle3, c_double_p, _ProfileP, C.c_double, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lsb200_fault_displacement_dev": (C.c_int, [_Handle3, c_double_p, _ProfileP, C.c_double, C.c_double, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lsb200_add_field_orthogonal_rows": (C.c_int, [Handle, Handle, c_double_p, C.c_int64, C.c_double, C.c_double, c_int64_p]),
    "lsb200_eval_lattice": (C.c_int, [Handle, c_double_p, c_double_p, c_double_p, c_int64_p, C.c_int, C.c_int, c_double_p]),
    "lsb200_eval_lattice_dev": (C.c_int, [Handle, c_double_p, c_double_p, c_double_p, c_int64_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "lsb200_isosurface": (C.c_int, [C.c_int, c_int64_p, c_double_p, C.c_double, c_double_p, C.POINTER(C.c_void_p), c_int64_p, c_int64_p]),
    "lsb200_isosurface_fetch": (C.c_int, [C.c_void_p, c_double_p, C.POINTER(C.c_int32), c_double_p]),
    "lsb200_isosurface_free": (C.c_int, [C.c_void_p]),
    "lsb200_host_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p)]),
    "lsb200_host_free": (C.c_int, [C.c_void_p]),
    "lsb200_time_kernel": (C.c_int, [Handle, C.c_int, C.c_int, c_double_p]),
    "lsb200_describe_plan": (C.c_int, [Handle, C.c_char_p, C.c_int64]),
    "lsb200_debug_small_phases": (C.c_int, [Handle, C.c_int, c_double_p]),
    "lsb200_debug_cta_trace": (C.c_int, [Handle, C.c_int64, C.POINTER(C.c_uint64), c_int64_p]),
}

_lib = None


def load() -> C.CDLL:
    """Load liblsb200.so and bind every declared symbol; raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("LSB200_LIBRARY", LIB_PATH)  # another build of the same ABI (A/B measurements)
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -m loopstructural_b200.build` "
            "(the engine has no CPU fallback)"
        )
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.lsb200_abi_version() != 1:
        raise ImportError("liblsb200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise Lsb200Error(rc, load().lsb200_last_error().decode("utf-8", "replace"))


def as_f64(a, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != shape:
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


def dptr(a: np.ndarray):
    return a.ctypes.data_as(c_double_p)


def iptr(a: np.ndarray):
    return a.ctypes.data_as(c_int64_p)
