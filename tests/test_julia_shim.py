"""The Julia shim cannot be executed here (no Julia in the image): what CAN be checked is that its struct declarations are the
C structs of include/mbm_b200.h field for field - through the ctypes mirror, whose sizes tests/test_abi.py checks against the
compiler - and that its constants and the symbols it binds are the header's."""
import ctypes as C
import re
from pathlib import Path

import mbm_b200  # noqa: F401
from mbm_b200 import abi

ROOT = Path(__file__).resolve().parent.parent
SRC = (ROOT / "julia" / "MovingBoundaryMineralsB200.jl").read_text()
HDR = (ROOT / "include" / "mbm_b200.h").read_text()

JL2C = {"Int32": C.c_int32, "UInt32": C.c_uint32, "Int64": C.c_int64, "Float64": C.c_double,
        "Ptr{Float64}": C.POINTER(C.c_double), "Ptr{Int32}": C.POINTER(C.c_int32), "Ptr{Int64}": C.POINTER(C.c_int64),
        "Ptr{UInt64}": C.POINTER(C.c_uint64)}


def julia_struct(name):
    body = re.search(r"^struct %s\b[^\n]*\n(.*?)^end" % name, SRC, re.S | re.M).group(1)
    body = re.sub(r"#[^\n]*", "", body)
    return [(f.strip(), t.strip()) for f, t in re.findall(r"(\w+)::([\w{}]+)", body)]


def test_struct_declarations_match_the_ctypes_mirror_field_for_field():
    for name in ("Common", "CoupleParams", "CoupleIO", "SimpleParams", "SimpleIO", "ValidateArgs"):
        jl = julia_struct(name)
        py = getattr(abi, name)._fields_
        assert [f for f, _ in jl] == [f for f, _ in py], name
        for (f, jt), (_, ct) in zip(jl, py):
            assert JL2C[jt] is ct or (issubclass(ct, C._Pointer) and JL2C[jt]._type_ is ct._type_), (name, f, jt, ct)


def test_constants_are_the_headers():
    def hdr_enum(sym):                          # enumerator (NAME = v) or macro (#define NAME v)
        m = re.search(r"\b%s\s*=\s*(0x[0-9a-fA-F]+|\d+)" % sym, HDR) or \
            re.search(r"#define\s+%s\s+(0x[0-9a-fA-F]+|\d+)" % sym, HDR)
        assert m, sym
        return int(m.group(1), 0)

    assert "const ABI_VERSION = Int32(%d)" % hdr_enum("MBM_ABI_VERSION") in SRC
    names = re.search(r"const (BC_FLUX[^=]*)=([^\n]*)", SRC)
    vals = [int(v) for v in re.findall(r"Int32\((\d+)\)", names.group(2))]
    hdr = [hdr_enum(s) for s in ("MBM_BC_FLUX", "MBM_BC_MB", "MBM_BC_STEFAN", "MBM_BC_DIRICHLET_HOLD", "MBM_BC_LASAGA")]
    assert vals == hdr
    vals = [int(v) for v in re.findall(r"Int32\((\d+)\)", re.search(r"const DT_FIND_DT[^=]*=([^\n]*)", SRC).group(1))]
    assert vals == [hdr_enum(s) for s in ("MBM_DT_FIND_DT", "MBM_DT_ADVECTIVE", "MBM_DT_STEFAN", "MBM_DT_SIMPLE")]
    vals = [int(v, 0) for v in re.findall(r"UInt32\((0x[0-9a-f]+)\)", re.search(r"const FLAG_RECORD[^=]*=([^\n]*)", SRC).group(1))]
    assert vals == [hdr_enum("MBM_FLAG_RECORD"), hdr_enum("MBM_FLAG_MASS_FINAL")]
    vals = [int(v) for v in re.findall(r"VAL_\w+ = Int32\((\d+)\)", SRC)]
    assert vals == [hdr_enum(s) for s in ("MBM_VAL_MASS_ERR", "MBM_VAL_RAYLEIGH", "MBM_VAL_SMITH", "MBM_VAL_SPHERE")]


def test_every_ccall_names_a_declared_entry_point_with_the_right_arity():
    calls = re.findall(r"ccall\(\(:(\w+), LIB\), (\w+),\s*\(([^)]*)\)", SRC)
    assert {c[0] for c in calls} >= {"mbm_couple_batch", "mbm_simple_diff_batch", "mbm_stefan_xt_batch", "mbm_create_grid",
                                     "mbm_calc_mass_vol", "mbm_validate_batch", "mbm_status_string"}
    for sym, _ret, args in calls:
        decl = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % sym, re.sub(r"/\*.*?\*/", "", HDR, flags=re.S), re.S)
        assert decl, sym
        n_c = len([a for a in decl.group(1).split(",") if a.strip() and a.strip() != "void"])
        n_jl = len([a for a in args.split(",") if a.strip()])
        assert n_c == n_jl, (sym, n_c, n_jl)


def _top_level_args(text, start):
    """arguments of the call whose '(' is at text[start]"""
    depth, args, cur = 0, [], []
    for ch in text[start:]:
        if ch in "([{":
            depth += 1
            if depth == 1:
                continue
        elif ch in ")]}":
            depth -= 1
            if depth == 0:
                args.append("".join(cur).strip())
                return [a for a in args if a]
        if ch == "," and depth == 1:
            args.append("".join(cur).strip()); cur = []
        else:
            cur.append(ch)
    raise AssertionError("unbalanced call")


def test_every_struct_constructor_call_passes_one_value_per_field():
    seen = 0
    for name in ("Common", "CoupleParams", "CoupleIO", "SimpleParams", "SimpleIO", "ValidateArgs"):
        nfields = len(getattr(abi, name)._fields_)
        for m in re.finditer(r"Ref\(%s\(" % name, SRC):
            args = _top_level_args(SRC, m.end() - 1)
            assert len(args) == nfields, (name, len(args), nfields, SRC[m.start(): m.start() + 80])
            seen += 1
    assert seen >= 8
