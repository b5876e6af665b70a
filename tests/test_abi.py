"""CPU-side checks of the drop-in boundary: struct layouts match the header, both libraries load and export
every symbol include/mbm_b200.h declares (no compute calls without a GPU)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

from mbm_b200 import abi

ROOT = Path(__file__).resolve().parent.parent


def test_header_symbols_match_binding_list():
    hdr = (ROOT / "include" / "mbm_b200.h").read_text()
    declared = set(re.findall(r"\b(mbm_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(abi.ABI_SYMBOLS), declared ^ set(abi.ABI_SYMBOLS)


def test_struct_sizes_match_c(tmp_path):
    src = tmp_path / "probe.c"
    src.write_text('#include <stdio.h>\n#include "mbm_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
                   "sizeof(mbm_common),sizeof(mbm_couple_params),sizeof(mbm_couple_io),sizeof(mbm_simple_params),"
                   "sizeof(mbm_simple_io));return 0;}\n")
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), "-o", str(exe), str(src)], check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(abi.Common), C.sizeof(abi.CoupleParams), C.sizeof(abi.CoupleIO),
                     C.sizeof(abi.SimpleParams), C.sizeof(abi.SimpleIO)]


def test_oracle_exports_every_symbol(oracle):
    for s in abi.ABI_SYMBOLS:
        assert hasattr(oracle, s), s
    assert oracle.mbm_backend() == b"oracle-cpu" and oracle.mbm_version() == abi.ABI_VERSION


def test_cuda_library_builds_and_exports_every_symbol():
    import __graft_entry__ as ge
    ge.build()
    lib = abi.load_cuda_library()
    for s in abi.ABI_SYMBOLS:
        assert hasattr(lib, s), s
    assert lib.mbm_version() == abi.ABI_VERSION
    assert lib.mbm_backend() == b"cuda-sm_100a"


def test_cuda_library_fails_loudly_without_device():
    """no silent CPU fallback: without a GPU the product returns MBM_RC_NO_DEVICE"""
    import numpy as np
    from mbm_b200 import drivers as d
    lib = abi.load_cuda_library()
    if lib.mbm_device_count() > 0:
        pytest.skip("a GPU is present")
    ens = d.build_simple_ensemble([d.pars_A1()])
    assert ens.run(lib) == -5
    with pytest.raises(RuntimeError):
        d.run_couple(d.pars_B1(), lib)


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    monkeypatch.setenv("MBM_B200_LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(abi, "_cuda_lib", None)
    with pytest.raises(RuntimeError):
        abi.load_cuda_library()
    monkeypatch.setattr(abi, "_cuda_lib", None)


def test_bad_arguments_are_reported_before_the_device_check():
    """argument checks come first and need no GPU: a malformed call is MBM_RC_BAD_ARG (-1) on any machine; only a
    well-formed call reaches the device check (MBM_RC_NO_DEVICE here, a launch on a GPU box).  The oracle exports
    the same ABI and rejects the same calls."""
    from mbm_b200 import drivers as d
    from oracle import load
    cuda = abi.load_cuda_library()
    ora = load.load()
    for lib in (cuda, ora):
        ens, _ = d.build_couple_ensemble([d.pars_DCGMB()], ora)
        c, p, io = ens.structs(max_steps=1)
        c.dt_mode = 9
        assert lib.mbm_couple_batch(C.byref(c), C.byref(p), C.byref(io)) == -1
        c.dt_mode = abi.DT_STEFAN; c.inner_bc = abi.BC_STEFAN              # Stefan without its coefficient arrays
        assert lib.mbm_couple_batch(C.byref(c), C.byref(p), C.byref(io)) == -1
        assert lib.mbm_stefan_xt_batch(C.byref(c), C.byref(p), C.byref(io)) == -1
        c.dt_mode = abi.DT_FIND_DT; c.inner_bc = abi.BC_MB
        p.nd = 3                                                           # density table announced, pointers NULL
        assert lib.mbm_couple_batch(C.byref(c), C.byref(p), C.byref(io)) == -1
        p.nd = 0; p.law = abi.LAW_OLIVINE_MELT_D1                          # material law without parameters
        assert lib.mbm_couple_batch(C.byref(c), C.byref(p), C.byref(io)) == -1
        p.law = 0; c.abi_version = abi.ABI_VERSION + 1
        assert lib.mbm_couple_batch(C.byref(c), C.byref(p), C.byref(io)) == -1
        c.abi_version = abi.ABI_VERSION
        io.x_left = None
        assert lib.mbm_couple_batch(C.byref(c), C.byref(p), C.byref(io)) == -1
    if cuda.mbm_device_count() == 0:
        ens, _ = d.build_couple_ensemble([d.pars_DCGMB()], ora)
        c, p, io = ens.structs(max_steps=1)
        assert cuda.mbm_couple_batch(C.byref(c), C.byref(p), C.byref(io)) == -5
