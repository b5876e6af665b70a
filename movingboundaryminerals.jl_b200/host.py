"""Host-side mirror of the reference's pre/post-loop code and the ensemble containers.

Nothing here is on the hot path: these functions run once per run, before and after the batched
time loop that the C-ABI library executes (`scaling`, `rescale`: reference
src/functions.jl:988-1008, :941-955; LinRange semantics: SURVEY.md §A.4).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import abi


# ------------------------------------------------------------------------------------------------
# Julia semantics used by the drivers
# ------------------------------------------------------------------------------------------------
def linrange(a: float, b: float, n: int) -> np.ndarray:
    """Julia LinRange(a,b,n) collected: element i is (1-t)*a + t*b with t = i/(n-1)."""
    d = max(n - 1, 1)
    t = np.arange(n, dtype=np.float64) / float(d)
    return (1.0 - t) * float(a) + t * float(b)


def steprange(dx: float, L: float) -> np.ndarray:
    """Julia [0:dx:L;] — a StepRangeLen with twice-precision arithmetic: elements are the
    correctly rounded (i)*dx (examples/A1.jl:21); length floor(L/dx)+1 (with the usual
    range-length guard against rounding)."""
    from fractions import Fraction

    n = int(np.floor(L / dx + 1e-9)) + 1
    fdx = Fraction(dx)
    return np.array([float(i * fdx) for i in range(n)], dtype=np.float64)


def scaling(Ri, Di, D0, V_ip, t_tot, t_ar_ends):
    """reference src/functions.jl:988-1008.  t_ar_ends = (start, stop) of the LinRange t_ar;
    `LinRange * scalar` stays a LinRange with scaled endpoints."""
    Ri = np.asarray(Ri, dtype=np.float64)
    Di = np.asarray(Di, dtype=np.float64)
    D0 = np.asarray(D0, dtype=np.float64)
    Lsc = 1e-3
    if Di[0] == -1 and Di[1] == -1:
        Dsc = (D0[0] + D0[1]) * (1.0 / 2.0)
    else:
        Dsc = (Di[0] + Di[1]) * (1.0 / 2.0)
    tsc = Lsc**2 * (1.0 / Dsc)
    Vsc = Dsc * (1.0 / Lsc)
    V = V_ip * (1.0 / Vsc)
    tt = t_tot * (1.0 / tsc)
    tar = (t_ar_ends[0] * (1.0 / tsc), t_ar_ends[1] * (1.0 / tsc))
    return V, tt, tar, Di * (1.0 / Dsc), D0 * (1.0 / Dsc), Ri * (1.0 / Lsc), Lsc, Dsc, Vsc, tsc


def rescale(arrs, Lsc):
    """reference src/functions.jl:941-955 (lengths only; the drivers discard the rest)."""
    return [np.asarray(a) * Lsc for a in arrs]


def sinusoid_profile(C0, n, L, D, t, G, x):
    """reference src/functions.jl:1266-1274 (used to build initial conditions of A1 / DCNI)."""
    x = np.asarray(x, dtype=np.float64)
    s = np.zeros_like(x)
    for ni, gi in zip(n, G):
        s = s + np.exp(-((ni * np.pi * (1.0 / L)) ** 2) * D * t) * gi * np.sin(ni * np.pi * x * (1.0 / L))
    return C0 + s


# ------------------------------------------------------------------------------------------------
# Ensemble containers (member index is the FIRST numpy axis == last column-major dimension)
# ------------------------------------------------------------------------------------------------
@dataclass
class CoupleEnsemble:
    """Argument block of mbm_couple_*_batch for M members."""

    M: int
    ndim: int
    inner_bc: int
    dt_mode: int
    cap_left: int
    cap_right: int
    hist_len: int
    R: float = 8.314472
    arrays: dict = field(default_factory=dict)

    def __post_init__(self):
        M, a = self.M, self.arrays
        z = lambda *s: np.zeros(s, dtype=np.float64)
        a.setdefault("Di", z(M, 2)); a.setdefault("D0", np.full((M, 2), np.nan)); a.setdefault("Ea", np.full((M, 2), np.nan))
        a.setdefault("rho", np.ones((M, 2))); a.setdefault("V_ip", z(M)); a.setdefault("t_tot", z(M))
        a.setdefault("CFL", z(M)); a.setdefault("MRefin", np.ones(M))
        a.setdefault("BCout", np.zeros((M, 2), dtype=np.int32)); a.setdefault("resmin", np.zeros((M, 2), dtype=np.int64))
        a.setdefault("t_ar", z(1, self.hist_len)); a.setdefault("T_ar", z(1, self.hist_len)); a.setdefault("KD_ar", z(1, self.hist_len))
        a.setdefault("x_left", z(M, self.cap_left)); a.setdefault("C_left", z(M, self.cap_left))
        a.setdefault("x_right", z(M, self.cap_right)); a.setdefault("C_right", z(M, self.cap_right))
        a.setdefault("res", np.zeros((M, 2), dtype=np.int64)); a.setdefault("Ri", z(M, 2))
        a.setdefault("t", z(M)); a.setdefault("dt", z(M)); a.setdefault("it", np.zeros(M, dtype=np.int64))
        a.setdefault("dx1", z(M)); a.setdefault("dx2", z(M)); a.setdefault("Mass0", z(M)); a.setdefault("Mass", z(M))
        a.setdefault("newton_R", z(M)); a.setdefault("last", z(M, 8))
        a.setdefault("status", np.zeros(M, dtype=np.int32)); a.setdefault("warn", np.zeros(M, dtype=np.int32))
        a.setdefault("idx_hash", np.zeros(M, dtype=np.uint64)); a.setdefault("node_steps", np.zeros(M, dtype=np.int64))
        # Stefan extras
        a.setdefault("coeff_up", None); a.setdefault("coeff_do", None)
        a.setdefault("C_leftB", None); a.setdefault("C_rightB", None)
        a.setdefault("dens", None)  # (X[nd], T[nd], left[nd,nd], right[nd,nd]) or None
        a.setdefault("law_par", None)  # [M,8] parameters of the in-loop material law (examples/D1.jl), or None
        a.setdefault("rec", None)

    def __getattr__(self, name):
        arrays = object.__getattribute__(self, "arrays")
        if name in arrays:
            return arrays[name]
        raise AttributeError(name)

    def copy(self) -> "CoupleEnsemble":
        arrs = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in self.arrays.items()}
        return CoupleEnsemble(self.M, self.ndim, self.inner_bc, self.dt_mode, self.cap_left, self.cap_right,
                              self.hist_len, self.R, arrs)

    def _stride(self, name):
        return 0 if self.arrays[name].shape[0] == 1 else self.hist_len

    def structs(self, flags=0, max_steps=0, rec_cap=0, n_devices=1, device_ids=None):
        a = self.arrays
        for k, v in list(a.items()):
            if isinstance(v, np.ndarray) and not v.flags["C_CONTIGUOUS"]:
                a[k] = np.ascontiguousarray(v)
        c = abi.Common(abi.ABI_VERSION, self.M, self.ndim, self.hist_len, self.cap_left, self.cap_right,
                       self.inner_bc, self.dt_mode, flags, n_devices, max_steps, None, rec_cap)
        if device_ids is not None:
            self._dev_ids = np.ascontiguousarray(device_ids, dtype=np.int32)
            c.device_ids = abi.i32ptr(self._dev_ids)
        p = abi.CoupleParams()
        for n in ("Di", "D0", "Ea", "rho", "V_ip", "t_tot", "CFL", "MRefin", "t_ar", "T_ar", "KD_ar"):
            setattr(p, n, abi.dptr(a[n]))
        p.BCout = abi.i32ptr(a["BCout"]); p.resmin = abi.i64ptr(a["resmin"])
        p.t_stride, p.T_stride, p.KD_stride = self._stride("t_ar"), self._stride("T_ar"), self._stride("KD_ar")
        p.R = self.R
        if a["coeff_up"] is not None:
            p.coeff_up = abi.dptr(a["coeff_up"]); p.coeff_do = abi.dptr(a["coeff_do"])
            p.coeff_stride = 0 if a["coeff_up"].shape[0] == 1 else 3
            p.C_leftB = abi.dptr(a["C_leftB"]); p.C_rightB = abi.dptr(a["C_rightB"])
        if a["dens"] is not None:
            X, T, L, Rr = a["dens"]
            p.nd = len(X)
            p.dens_X, p.dens_T, p.dens_left, p.dens_right = abi.dptr(X), abi.dptr(T), abi.dptr(L), abi.dptr(Rr)
        if a.get("law_par") is not None:
            p.law = int(a.get("law", abi.LAW_OLIVINE_MELT_D1)); p.law_par = abi.dptr(a["law_par"])
        io = abi.CoupleIO()
        for n in ("x_left", "C_left", "x_right", "C_right", "Ri", "t", "dt", "dx1", "dx2", "Mass0", "Mass",
                  "newton_R", "last"):
            setattr(io, n, abi.dptr(a[n]))
        io.res = abi.i64ptr(a["res"]); io.it = abi.i64ptr(a["it"]); io.node_steps = abi.i64ptr(a["node_steps"])
        io.status = abi.i32ptr(a["status"]); io.warn = abi.i32ptr(a["warn"])
        if a["idx_hash"] is not None:                     # optional output: None = no regrid checksum (NULL in the ABI)
            io.idx_hash = abi.u64ptr(a["idx_hash"])
        if rec_cap > 0:
            if a["rec"] is None or a["rec"].shape != (self.M, rec_cap, 4):
                a["rec"] = np.zeros((self.M, rec_cap, 4), dtype=np.float64)
            io.rec = abi.dptr(a["rec"])
        return c, p, io

    def run(self, lib, flags=0, max_steps=0, rec_cap=0, n_devices=1, device_ids=None) -> int:
        """One call of the batched time loop.  Returns the library return code."""
        if rec_cap > 0:
            flags |= abi.FLAG_RECORD
        c, p, io = self.structs(flags, max_steps, rec_cap, n_devices, device_ids)
        return lib.mbm_couple_batch(C.byref(c), C.byref(p), C.byref(io))

    # pre-loop helpers -------------------------------------------------------------------------
    def create_grid(self, lib):
        a = self.arrays
        st = np.zeros(self.M, dtype=np.int32)
        rc = lib.mbm_create_grid(self.M, self.cap_left, self.cap_right, abi.dptr(a["Ri"]), abi.i64ptr(a["res"]),
                                 abi.dptr(a["MRefin"]), abi.dptr(a["x_left"]), abi.dptr(a["x_right"]),
                                 abi.dptr(a["dx1"]), abi.dptr(a["dx2"]), abi.dptr(a["newton_R"]), abi.i32ptr(st))
        if rc != 0:
            raise RuntimeError(f"mbm_create_grid failed rc={rc}")
        a["status"][:] = st
        return st

    def calc_mass(self, lib) -> np.ndarray:
        a = self.arrays
        out = np.zeros(self.M, dtype=np.float64)
        rc = lib.mbm_calc_mass_vol(self.M, self.ndim, self.cap_left, self.cap_right, abi.dptr(a["x_left"]),
                                   abi.dptr(a["x_right"]), abi.dptr(a["C_left"]), abi.dptr(a["C_right"]),
                                   abi.i64ptr(a["res"]), abi.dptr(a["rho"]), abi.dptr(out))
        if rc != 0:
            raise RuntimeError(f"mbm_calc_mass_vol failed rc={rc}")
        return out

    def validate(self, lib, kind, Ri0=None, C0_r=None, KD0=None, Di_r=None, V_ip=None, t_tot=None, device=0, zeta=None,
                 C0_left=None, C0_right=None):
        """mbm_validate_batch on the ensemble's state: (out[M,2], relerr[M]).  kind = abi.VAL_MASS_ERR (calc_mass_err of
        calc_mass_vol against Mass0), abi.VAL_RAYLEIGH (test_C1.jl's gate; needs Ri0[M,2], C0_r[M], KD0[M]) or
        abi.VAL_SMITH (test_B5.jl's gate, n = 1; needs KD0, Di_r, V_ip, t_tot [M], in the units of the state) or
        abi.VAL_SPHERE (test_B1.jl / test_B2.jl: sphere series; needs zeta [M] = time-integrated diffusivity and the initial
        profiles C0_left [M, cap_left], C0_right [M, cap_right])."""
        a = self.arrays
        keep = []
        def dp(v):
            if v is None:
                return None
            v = np.ascontiguousarray(v, dtype=np.float64); keep.append(v)
            return abi.dptr(v)
        out = np.zeros((self.M, 2), dtype=np.float64); err = np.zeros(self.M, dtype=np.float64)
        va = abi.ValidateArgs(abi.ABI_VERSION, self.M, self.ndim, self.cap_left, self.cap_right, kind, 0, device)
        va.x_left, va.C_left, va.x_right, va.C_right = (abi.dptr(a[k]) for k in ("x_left", "C_left", "x_right", "C_right"))
        va.res = abi.i64ptr(a["res"]); va.Ri = abi.dptr(a["Ri"]); va.rho = abi.dptr(a["rho"]); va.Mass0 = abi.dptr(a["Mass0"])
        va.Ri0, va.C0_r, va.KD0, va.Di_r, va.V_ip, va.t_tot = dp(Ri0), dp(C0_r), dp(KD0), dp(Di_r), dp(V_ip), dp(t_tot)
        va.out, va.relerr = abi.dptr(out), abi.dptr(err)
        va.zeta, va.C0_left, va.C0_right = dp(zeta), dp(C0_left), dp(C0_right)
        rc = lib.mbm_validate_batch(C.byref(va))
        if rc != 0:
            raise RuntimeError(f"mbm_validate_batch failed rc={rc}")
        return out, err

    def member(self, m):
        """(x_left, C_left, x_right, C_right) views trimmed to the member's node counts."""
        a = self.arrays
        n1, n2 = int(a["res"][m, 0]), int(a["res"][m, 1])
        return a["x_left"][m, :n1], a["C_left"][m, :n1], a["x_right"][m, :n2], a["C_right"][m, :n2]


@dataclass
class SimpleEnsemble:
    """Argument block of mbm_simple_diff_batch for M members."""

    M: int
    ndim: int
    cap: int
    hist_len: int
    R: float = 8.314472
    arrays: dict = field(default_factory=dict)

    def __post_init__(self):
        M, a = self.M, self.arrays
        z = lambda *s: np.zeros(s, dtype=np.float64)
        for n in ("Di", "D0", "Ea1", "rho", "dx", "t_tot", "CFL", "t", "dt", "D", "Mass"):
            a.setdefault(n, z(M))
        a.setdefault("BCout", np.zeros((M, 2), dtype=np.int32))
        a.setdefault("t_ar", z(1, self.hist_len)); a.setdefault("T_ar", z(1, self.hist_len))
        a.setdefault("x", z(M, self.cap)); a.setdefault("C", z(M, self.cap))
        a.setdefault("res", np.zeros(M, dtype=np.int64)); a.setdefault("it", np.zeros(M, dtype=np.int64))
        a.setdefault("status", np.zeros(M, dtype=np.int32)); a.setdefault("node_steps", np.zeros(M, dtype=np.int64))

    def __getattr__(self, name):
        arrays = object.__getattribute__(self, "arrays")
        if name in arrays:
            return arrays[name]
        raise AttributeError(name)

    def copy(self):
        return SimpleEnsemble(self.M, self.ndim, self.cap, self.hist_len, self.R,
                              {k: v.copy() for k, v in self.arrays.items()})

    def structs(self, flags=0, max_steps=0, n_devices=1):
        a = self.arrays
        c = abi.Common(abi.ABI_VERSION, self.M, self.ndim, self.hist_len, self.cap, 0, abi.BC_NONE,
                       abi.DT_SIMPLE, flags, n_devices, max_steps, None, 0)
        p = abi.SimpleParams()
        for n in ("Di", "D0", "Ea1", "rho", "dx", "t_tot", "CFL", "t_ar", "T_ar"):
            setattr(p, n, abi.dptr(a[n]))
        p.BCout = abi.i32ptr(a["BCout"])
        p.t_stride = 0 if a["t_ar"].shape[0] == 1 else self.hist_len
        p.T_stride = 0 if a["T_ar"].shape[0] == 1 else self.hist_len
        p.R = self.R
        io = abi.SimpleIO()
        for n in ("x", "C", "t", "dt", "D", "Mass"):
            setattr(io, n, abi.dptr(a[n]))
        io.res = abi.i64ptr(a["res"]); io.it = abi.i64ptr(a["it"]); io.status = abi.i32ptr(a["status"])
        io.node_steps = abi.i64ptr(a["node_steps"])
        return c, p, io

    def run(self, lib, flags=0, max_steps=0, n_devices=1) -> int:
        c, p, io = self.structs(flags, max_steps, n_devices)
        return lib.mbm_simple_diff_batch(C.byref(c), C.byref(p), C.byref(io))
