"""ctypes mirror of include/mbm_b200.h.

The struct layouts below must match the header field by field; tests/test_abi.py checks the
sizes against a C probe and that every declared symbol is exported by the built libraries.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

ABI_VERSION = 4
LAW_NONE, LAW_OLIVINE_MELT_D1 = 0, 1      # mbm_couple_params.law (examples/D1.jl:126-132)

# interface conditions / dt rules / flags (include/mbm_b200.h)
BC_NONE, BC_FLUX, BC_MB, BC_STEFAN, BC_DIRICHLET_HOLD, BC_LASAGA = 0, 1, 2, 3, 4, 5
DT_FIND_DT, DT_ADVECTIVE, DT_STEFAN, DT_SIMPLE = 0, 1, 2, 3
FLAG_NEWTON_WARM = 0x0001
FLAG_RECORD = 0x0002
FLAG_SOLVER_FAITHFUL = 0x0004
FLAG_SUM_JULIA = 0x0008
FLAG_DEVICE_POINTERS = 0x0010
FLAG_FORCE_RESIDENT = 0x0020
FLAG_FORCE_STREAMING = 0x0040
FLAG_MASS_FINAL = 0x0080

STATUS = {
    0: "ok", 1: "iface_fast_pos", 2: "iface_fast_neg", 3: "newton", 4: "kd_range", 5: "t_range",
    6: "interp_bounds", 7: "ndim", 8: "pchip_index", 9: "density_bounds", 10: "capacity",
    11: "nonfinite", 12: "mrefin",
}

_dp = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u64p = C.POINTER(C.c_uint64)


class Common(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("n_members", C.c_int32), ("ndim", C.c_int32),
        ("hist_len", C.c_int32), ("cap_left", C.c_int32), ("cap_right", C.c_int32),
        ("inner_bc", C.c_int32), ("dt_mode", C.c_int32), ("flags", C.c_uint32),
        ("n_devices", C.c_int32), ("max_steps", C.c_int64), ("device_ids", _i32p),
        ("rec_cap", C.c_int64),
    ]


class CoupleParams(C.Structure):
    _fields_ = [
        ("Di", _dp), ("D0", _dp), ("Ea", _dp), ("rho", _dp), ("V_ip", _dp), ("t_tot", _dp),
        ("CFL", _dp), ("MRefin", _dp), ("BCout", _i32p), ("resmin", _i64p),
        ("t_ar", _dp), ("T_ar", _dp), ("KD_ar", _dp),
        ("t_stride", C.c_int64), ("T_stride", C.c_int64), ("KD_stride", C.c_int64),
        ("R", C.c_double),
        ("coeff_up", _dp), ("coeff_do", _dp), ("coeff_stride", C.c_int64),
        ("C_leftB", _dp), ("C_rightB", _dp), ("nd", C.c_int32), ("law", C.c_int32),
        ("dens_X", _dp), ("dens_T", _dp), ("dens_left", _dp), ("dens_right", _dp), ("law_par", _dp),
    ]


class CoupleIO(C.Structure):
    _fields_ = [
        ("x_left", _dp), ("C_left", _dp), ("x_right", _dp), ("C_right", _dp),
        ("res", _i64p), ("Ri", _dp), ("t", _dp), ("dt", _dp), ("it", _i64p),
        ("dx1", _dp), ("dx2", _dp), ("Mass0", _dp), ("Mass", _dp), ("newton_R", _dp),
        ("last", _dp), ("status", _i32p), ("warn", _i32p), ("idx_hash", _u64p),
        ("node_steps", _i64p), ("rec", _dp),
    ]


class SimpleParams(C.Structure):
    _fields_ = [
        ("Di", _dp), ("D0", _dp), ("Ea1", _dp), ("rho", _dp), ("dx", _dp), ("t_tot", _dp),
        ("CFL", _dp), ("BCout", _i32p), ("t_ar", _dp), ("T_ar", _dp),
        ("t_stride", C.c_int64), ("T_stride", C.c_int64), ("R", C.c_double),
    ]


class SimpleIO(C.Structure):
    _fields_ = [
        ("x", _dp), ("C", _dp), ("res", _i64p), ("t", _dp), ("dt", _dp), ("it", _i64p),
        ("D", _dp), ("Mass", _dp), ("status", _i32p), ("node_steps", _i64p),
    ]


# every symbol include/mbm_b200.h declares
VAL_MASS_ERR, VAL_RAYLEIGH, VAL_SMITH, VAL_SPHERE = 0, 1, 2, 3   # mbm_validate_args.kind


class ValidateArgs(C.Structure):
    """mbm_validate_args: the known-answer gates of the reference's tests, per member (include/mbm_b200.h)"""
    _fields_ = [
        ("abi_version", C.c_int32), ("n_members", C.c_int32), ("ndim", C.c_int32), ("cap_left", C.c_int32),
        ("cap_right", C.c_int32), ("kind", C.c_int32), ("flags", C.c_uint32), ("device", C.c_int32),
        ("x_left", _dp), ("C_left", _dp), ("x_right", _dp), ("C_right", _dp), ("res", _i64p), ("Ri", _dp),
        ("rho", _dp), ("Mass0", _dp), ("Ri0", _dp), ("C0_r", _dp), ("KD0", _dp), ("Di_r", _dp), ("V_ip", _dp),
        ("t_tot", _dp), ("out", _dp), ("relerr", _dp), ("zeta", _dp), ("C0_left", _dp), ("C0_right", _dp),
    ]


ABI_SYMBOLS = [
    "mbm_version", "mbm_device_count", "mbm_backend", "mbm_status_string",
    "mbm_simple_diff_batch", "mbm_couple_flux_batch", "mbm_couple_mb_batch",
    "mbm_stefan_xt_batch", "mbm_couple_batch", "mbm_create_grid", "mbm_calc_mass_vol",
    "mbm_last_kernel_seconds", "mbm_last_kernel_launches", "mbm_set_stream", "mbm_measure_dfma_tflops",
    "mbm_validate_batch", "mbm_stream_pass_counts", "mbm_last_tuned_capacity",
]

PKG_DIR = Path(__file__).resolve().parent
CUDA_LIB_PATH = PKG_DIR / "csrc" / "libmbm_b200.so"


def bind(lib: C.CDLL) -> C.CDLL:
    """Attach argtypes/restypes for every ABI symbol (raises AttributeError if one is missing)."""
    lib.mbm_version.restype = C.c_int
    lib.mbm_device_count.restype = C.c_int
    lib.mbm_backend.restype = C.c_char_p
    lib.mbm_status_string.restype = C.c_char_p
    lib.mbm_status_string.argtypes = [C.c_int32]
    for name, P, IO in (
        ("mbm_simple_diff_batch", SimpleParams, SimpleIO),
        ("mbm_couple_flux_batch", CoupleParams, CoupleIO),
        ("mbm_couple_mb_batch", CoupleParams, CoupleIO),
        ("mbm_stefan_xt_batch", CoupleParams, CoupleIO),
        ("mbm_couple_batch", CoupleParams, CoupleIO),
    ):
        f = getattr(lib, name)
        f.restype = C.c_int
        f.argtypes = [C.POINTER(Common), C.POINTER(P), C.POINTER(IO)]
    lib.mbm_create_grid.restype = C.c_int
    lib.mbm_create_grid.argtypes = [C.c_int32, C.c_int32, C.c_int32, _dp, _i64p, _dp, _dp, _dp, _dp,
                                    _dp, _dp, _i32p]
    lib.mbm_calc_mass_vol.restype = C.c_int
    lib.mbm_calc_mass_vol.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, _dp, _dp, _dp, _dp,
                                      _i64p, _dp, _dp]
    lib.mbm_last_kernel_seconds.restype = C.c_double
    lib.mbm_last_kernel_launches.restype = C.c_int64
    lib.mbm_set_stream.restype = C.c_int
    lib.mbm_set_stream.argtypes = [C.c_void_p]
    lib.mbm_measure_dfma_tflops.restype = C.c_double
    lib.mbm_measure_dfma_tflops.argtypes = [C.c_int]
    lib.mbm_validate_batch.restype = C.c_int
    lib.mbm_validate_batch.argtypes = [C.POINTER(ValidateArgs)]
    lib.mbm_last_tuned_capacity.restype = C.c_int
    lib.mbm_stream_pass_counts.restype = C.c_int
    lib.mbm_stream_pass_counts.argtypes = [C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    return lib


_cuda_lib = None


def load_cuda_library() -> C.CDLL:
    """Load libmbm_b200.so (the product).  There is NO fallback: a missing library is an error."""
    global _cuda_lib
    if _cuda_lib is None:
        path = Path(os.environ.get("MBM_B200_LIB", CUDA_LIB_PATH))
        if not path.exists():
            raise RuntimeError(
                f"libmbm_b200.so not found at {path}: build it with `python __graft_entry__.py build` "
                "(nvcc, sm_100a).  There is no CPU fallback."
            )
        _cuda_lib = bind(C.CDLL(str(path)))
        if _cuda_lib.mbm_version() != ABI_VERSION:
            raise RuntimeError("libmbm_b200.so ABI version mismatch")
    return _cuda_lib


# ------------------------------------------------------------------------------------------------
# numpy helpers
# ------------------------------------------------------------------------------------------------
def f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = np.ascontiguousarray(np.broadcast_to(a, shape), dtype=np.float64)
    return a


def dptr(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def i32ptr(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_i32p)


def i64ptr(a):
    if a is None:
        return None
    assert a.dtype == np.int64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_i64p)


def u64ptr(a):
    if a is None:
        return None
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u64p)
