"""Builds (if needed) and loads the CPU oracle.  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the product."""
import ctypes as C
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import mbm_b200  # noqa: E402
from mbm_b200 import abi  # noqa: E402

LIB = HERE / "_build" / "libmbm_oracle.so"


def build(force=False):
    src = HERE / "mbm_oracle.c"
    hdr = HERE.parent / "include" / "mbm_b200.h"
    if force or not LIB.exists() or LIB.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["make", "-C", str(HERE), "-s"], check=True)
    return LIB


_lib = None


def load():
    global _lib
    if _lib is None:
        build()
        lib = abi.bind(C.CDLL(str(LIB)))
        dp, i64p, i32p = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int32)
        lib.mbm_oracle_create_grid_flags.restype = C.c_int
        lib.mbm_oracle_create_grid_flags.argtypes = [C.c_int32, C.c_int32, C.c_int32, dp, i64p, dp, C.c_uint32, dp, dp, i32p]
        lib.mbm_oracle_pchip.restype = C.c_int
        lib.mbm_oracle_pchip.argtypes = [dp, dp, C.c_int64, dp, C.c_int64, dp]
        lib.mbm_oracle_newton.restype = C.c_int
        lib.mbm_oracle_newton.argtypes = [C.c_double, C.c_double, C.c_int64, C.c_double, dp, C.POINTER(C.c_int)]
        lib.mbm_oracle_pow_int.restype = C.c_double
        lib.mbm_oracle_pow_int.argtypes = [C.c_double, C.c_int64]
        lib.mbm_oracle_scan.restype = None
        lib.mbm_oracle_scan.argtypes = [dp, C.c_int64, dp, C.c_int, C.c_int]
        lib.mbm_oracle_capture_next_system.restype = None
        lib.mbm_oracle_capture_next_system.argtypes = [dp, dp, dp, C.c_int64, i64p]
        lib.mbm_oracle_set_threads.restype = C.c_int
        lib.mbm_oracle_set_threads.argtypes = [C.c_int]
        _lib = lib
    return _lib
