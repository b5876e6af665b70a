"""Drop-in drivers with the reference's entry-point names and parameter conventions.

Each function mirrors one reference driver: same parameter names, same defaults, same return
tuple (reference main_codes/*.jl and examples/*.jl); the `while t < t_tot` loop runs inside
the C-ABI library (`lib`), everything else (parameter block, `scaling`, initial condition,
`rescale`) is host code as in the reference.  `lib=None` loads libmbm_b200.so (CUDA); the
tests pass the CPU oracle instead to pin the host logic without a GPU.

The only addition is the batch dimension: `build_couple_ensemble([pars, pars, ...])`.
"""
from __future__ import annotations

import numpy as np

from . import abi
from .host import CoupleEnsemble, SimpleEnsemble, linrange, scaling, sinusoid_profile, steprange

Myr2Sec = 60 * 60 * 24 * 365.25 * 1e6


def _lib(lib):
    return abi.load_cuda_library() if lib is None else lib


# ------------------------------------------------------------------------------------------------
# parameter blocks exactly as shipped (reference file:line in each docstring)
# ------------------------------------------------------------------------------------------------
def pars_DCGMB():
    """main_codes/Diff_couple_1D_general_MB.jl:7-33"""
    Cl_i, Cr_i = 0.5, 0.7
    return dict(kind="mb", Di=[4e-10, 5e-8], D0=[1e-4, 5e-4], rho=[1.0, 1.0], Ri=[0.2, 0.8], Cl_i=Cl_i, Cr_i=Cr_i,
                V_ip=-1e-7, R=8.314472, Ea1=292879.6767, Ea2=300660.4018, t_tot=1e4, n=2,
                KD_ar=(Cl_i / Cr_i, Cl_i / Cr_i * 0.9), T_ar=(973.15, 723.15), CFL=0.5, res=[100, 150],
                MRefin=2.0, BCout=[0, 0], scale=True, dt_mode=abi.DT_FIND_DT, init="uniform")


def pars_DCGF():
    """main_codes/Diff_couple_1D_general_Flux.jl:7-33"""
    Cl_i = 0.5
    Cr_i = Cl_i / 100
    return dict(kind="flux", Di=[4e-8, 5e-8], D0=[1e-4, 5e-4], rho=[1.0, 1.0], Ri=[0.5, 1.0], Cl_i=Cl_i, Cr_i=Cr_i,
                V_ip=1e-7, R=8.314472, Ea1=292879.6767, Ea2=300660.4018, t_tot=1e4, n=3,
                KD_ar=(Cl_i / Cr_i, Cl_i / Cr_i), T_ar=(1273.15, 923.15), CFL=0.3, res=[100, 150],
                MRefin=2.0, BCout=[0, 0], scale=True, dt_mode=abi.DT_FIND_DT, init="uniform")


def pars_B1():
    """examples/B1.jl:7-33 (unscaled SI)"""
    return dict(kind="flux", Di=[2.65e-18, 2.65e-18], D0=[np.nan, np.nan], rho=[1.0, 1.0], Ri=[0.0005, 0.001],
                Cl_i=0.0, Cr_i=1.0, V_ip=0.0, R=np.nan, Ea1=np.nan, Ea2=np.nan, t_tot=1e-3 * Myr2Sec, n=3,
                KD_ar=(1.0, 1.0), T_ar=(1273.15, 1273.15), CFL=0.3, res=[50, 75], MRefin=15.0, BCout=[0, 1],
                scale=False, dt_mode=abi.DT_FIND_DT, init="b1")


def pars_B2():
    """examples/B2.jl:7-33"""
    p = pars_B1()
    p.update(Di=[-1.0, -1.0], D0=[2.75e-6, 2.75e-6], R=8.314472, Ea1=292879.6767, Ea2=292879.6767,
             t_tot=2e-3 * Myr2Sec, T_ar=(1273.0, 1073.0))
    return p


def pars_B4():
    """examples/B4.jl:7-33, dt rule :93"""
    Cl_i = 0.5
    Cr_i = Cl_i / 100
    return dict(kind="flux", Di=[1e-4, 1e4], D0=[np.nan, np.nan], rho=[1.0, 1.0], Ri=[1e-2, 1.0], Cl_i=Cl_i, Cr_i=Cr_i,
                V_ip=1.0, R=np.nan, Ea1=np.nan, Ea2=np.nan, t_tot=0.35, n=3, KD_ar=(Cl_i / Cr_i, Cl_i / Cr_i),
                T_ar=(1273.15, 1273.15), CFL=0.4, res=[80, 120], MRefin=50.0, BCout=[0, 0], scale=True,
                dt_mode=abi.DT_ADVECTIVE, init="uniform")


def pars_B5():
    """examples/B5.jl:7-33, dt rule :93"""
    Cl_i = 0.1
    Cr_i = Cl_i / 1e-3
    return dict(kind="flux", Di=[1e-5, 1e1], D0=[np.nan, np.nan], rho=[1.0, 1.0], Ri=[1e-2, 10.0], Cl_i=Cl_i, Cr_i=Cr_i,
                V_ip=1.0, R=np.nan, Ea1=np.nan, Ea2=np.nan, t_tot=0.4, n=1, KD_ar=(Cl_i / Cr_i, Cl_i / Cr_i),
                T_ar=(1273.15, 1273.15), CFL=0.80, res=[100, 160], MRefin=10.0, BCout=[0, 0], scale=True,
                dt_mode=abi.DT_ADVECTIVE, init="uniform")


def pars_C1():
    """examples/C1.jl:8-34, dt rule :93"""
    p = pars_B4()
    p.update(kind="mb")
    return p


def pars_MB_growth():
    """examples/Diff_couple_MB_growth.jl:8-34, initial condition :55-56"""
    Cl_i = 0.5
    Cr_i = Cl_i / 0.1
    return dict(kind="mb", Di=[4e-7, 5e-8], D0=[np.nan, np.nan], rho=[1.0, 1.0], Ri=[0.5, 1.0], Cl_i=Cl_i, Cr_i=Cr_i,
                V_ip=-1e-7, R=np.nan, Ea1=np.nan, Ea2=np.nan, t_tot=2e4, n=2, KD_ar=(Cl_i / Cr_i, Cl_i / Cr_i),
                T_ar=(1273.15, 923.15), CFL=0.3, res=[200, 300], MRefin=20.0, BCout=[0, 0], scale=True,
                dt_mode=abi.DT_FIND_DT, init="mb_growth")


def pars_B6():
    """examples/B6.jl:7-33: flux couple, growing sphere, KD(t) falling to half its initial value"""
    Cl_i = 0.1
    Cr_i = Cl_i / 50.0
    return dict(kind="flux", Di=[0.001, 0.005], D0=[np.nan, np.nan], rho=[1.0, 1.0], Ri=[1.0, 10.0], Cl_i=Cl_i, Cr_i=Cr_i,
                V_ip=5e-3, R=np.nan, Ea1=np.nan, Ea2=np.nan, t_tot=3e2, n=3, KD_ar=(Cl_i / Cr_i, Cl_i / Cr_i * 0.5),
                T_ar=(1273.0, 1073.0), CFL=0.4, res=[100, 100], MRefin=5.0, BCout=[0, 0], scale=True,
                dt_mode=abi.DT_FIND_DT, init="uniform")


def pars_B7():
    """examples/B7.jl:7-33: flux couple, resorption (V_ip < 0), KD(t); ~3e5 steps at CFL 0.005"""
    Cl_i = 0.1
    Cr_i = Cl_i / 20.0
    return dict(kind="flux", Di=[0.001, 0.005], D0=[np.nan, np.nan], rho=[1.0, 1.0], Ri=[5.0, 10.0], Cl_i=Cl_i, Cr_i=Cr_i,
                V_ip=-5e-3, R=np.nan, Ea1=np.nan, Ea2=np.nan, t_tot=3e2, n=3, KD_ar=(Cl_i / Cr_i, Cl_i / Cr_i * 0.5),
                T_ar=(1273.0, 1073.0), CFL=0.005, res=[200, 250], MRefin=10.0, BCout=[0, 0], scale=True,
                dt_mode=abi.DT_FIND_DT, init="uniform")


def pars_C2():
    """examples/C2.jl:8-34: B7's couple with the mass-balance row"""
    p = pars_B7()
    p.update(kind="mb", T_ar=(1273.15, 1073.15))
    return p


def pars_Flux_growth():
    """examples/Diff_couple_Flux_growth.jl:8-34: flux couple, cylinder, resorption, KD(t) falling to 0.4 KD0"""
    Cl_i = 0.5
    Cr_i = Cl_i / 0.1
    return dict(kind="flux", Di=[4e-7, 5e-8], D0=[np.nan, np.nan], rho=[1.0, 1.0], Ri=[0.5, 1.0], Cl_i=Cl_i, Cr_i=Cr_i,
                V_ip=-1e-7, R=np.nan, Ea1=np.nan, Ea2=np.nan, t_tot=1e6, n=2, KD_ar=(Cl_i / Cr_i, Cl_i / Cr_i * 0.4),
                T_ar=(1273.15, 923.15), CFL=0.3, res=[100, 150], MRefin=10.0, BCout=[0, 0], scale=True,
                dt_mode=abi.DT_FIND_DT, init="uniform")


def pars_DC_Flux():
    """examples/Diff_couple_Flux.jl:7-33 (unscaled SI): fixed interface, Arrhenius diffusivities on a cooling path, KD(t);
    initial condition :54-55"""
    return dict(kind="flux", Di=[-1.0, -1.0], D0=[2.75e-6, 3.9e-7], rho=[1.0, 1.0], Ri=[0.0002, 0.0005], Cl_i=0.6,
                Cr_i=0.3, V_ip=0.0, R=8.314472, Ea1=292879.6767, Ea2=360660.4018, t_tot=1e-3 * Myr2Sec, n=3,
                KD_ar=(1.0, 0.7), T_ar=(1273.15, 973.15), CFL=0.5, res=[100, 150], MRefin=2.0, BCout=[0, 0],
                scale=False, dt_mode=abi.DT_FIND_DT, init="mb_growth")


def pars_DC_MB():
    """examples/Diff_couple_MB.jl:7-33: the same couple with the mass-balance row"""
    p = pars_DC_Flux()
    p.update(kind="mb")
    return p


def pars_DCNI():
    """examples/Diff_couple_no_interaction.jl:7-34"""
    return dict(kind="hold", Di=[2.65e-18, 2.65e-18], D0=[np.nan, np.nan], rho=[1.0, 1.0], Ri=[0.0002, 0.0004],
                Cl_i=0.6, Cr_i=0.6, V_ip=0.0, R=np.nan, Ea1=np.nan, Ea2=np.nan, t_tot=1e-5 * Myr2Sec, n=1,
                KD_ar=(1.0, 1.0), T_ar=(1273.15, 1273.15), CFL=0.5, res=[500, 500], MRefin=1.0, BCout=[1, 1],
                scale=False, dt_mode=abi.DT_FIND_DT, init="dcni")


_KIND = {"flux": abi.BC_FLUX, "mb": abi.BC_MB, "hold": abi.BC_DIRICHLET_HOLD, "stefan": abi.BC_STEFAN,
         "lasaga": abi.BC_LASAGA}

DCNI_MODES_L = (np.array([1, 2, 5, 7, 12.0]), np.array([12, 0.5, 3, -2, 1.0]))
DCNI_MODES_R = (np.array([1, 2, 3.0]), np.array([1, 8.0, 1 * (1.0 / 3.0)]))


def build_couple_ensemble(pars_list, lib=None, cap_left=None, cap_right=None, hist_len=1000):
    """Pre-loop part of DCGMB()/DCGF() (e.g. ..._MB.jl:35-85) for a list of members.

    All members must share kind / n / dt_mode.  Returns (ensemble, aux) where aux holds the
    per-member scales and initial data the drivers return (x0, C0, Ri0, Lsc ...)."""
    lib = _lib(lib)
    M = len(pars_list)
    p0 = pars_list[0]
    res_max = np.max([p["res"] for p in pars_list], axis=0)
    if cap_left is None:
        cap_left = int(2 * res_max[0] + 8)
    if cap_right is None:
        cap_right = int(2 * res_max[1] + 8)
    ens = CoupleEnsemble(M, int(p0["n"]), _KIND[p0["kind"]], int(p0["dt_mode"]), cap_left, cap_right, hist_len,
                         R=float(p0["R"]),
                         arrays=dict(t_ar=np.zeros((M, hist_len)), T_ar=np.zeros((M, hist_len)),
                                     KD_ar=np.zeros((M, hist_len))))
    a = ens.arrays
    aux = dict(Lsc=np.ones(M), Dsc=np.ones(M), Vsc=np.ones(M), tsc=np.ones(M), Ri0=np.zeros((M, 2)))
    for m, p in enumerate(pars_list):
        assert p["kind"] == p0["kind"] and p["n"] == p0["n"] and p["dt_mode"] == p0["dt_mode"]
        Ri, Di, D0, V, tt = np.array(p["Ri"], float), np.array(p["Di"], float), np.array(p["D0"], float), p["V_ip"], p["t_tot"]
        tar = (0.0, p["t_tot"])
        if p["scale"]:
            V, tt, tar, Di, D0, Ri, Lsc, Dsc, Vsc, tsc = scaling(Ri, Di, D0, V, tt, tar)
            aux["Lsc"][m], aux["Dsc"][m], aux["Vsc"][m], aux["tsc"][m] = Lsc, Dsc, Vsc, tsc
        if Ri[0] >= Ri[1]:
            raise ValueError("Please change the size of the system. Increase Ri[2].")
        if p["res"][0] > p["res"][1]:
            raise ValueError("Please change the resolution of the system. res[2] >= res[1].")
        a["Di"][m], a["D0"][m], a["Ea"][m], a["rho"][m] = Di, D0, [p["Ea1"], p["Ea2"]], p["rho"]
        a["V_ip"][m], a["t_tot"][m], a["CFL"][m], a["MRefin"][m] = V, tt, p["CFL"], p["MRefin"]
        a["BCout"][m] = p["BCout"]
        a["resmin"][m] = p.get("resmin", p["res"])
        a["res"][m] = p["res"]
        a["Ri"][m] = Ri
        aux["Ri0"][m] = Ri
        a["t_ar"][m] = linrange(tar[0], tar[1], hist_len)
        a["T_ar"][m] = linrange(p["T_ar"][0], p["T_ar"][1], hist_len)
        a["KD_ar"][m] = linrange(p["KD_ar"][0], p["KD_ar"][1], hist_len)
        if np.any(np.diff(a["t_ar"][m]) <= 0.0) or np.any(a["t_ar"][m] < 0.0) or np.any(a["t_ar"][m] > tt):
            raise ValueError("The time array is not valid. Please check your inputs.")
    st = ens.create_grid(lib)                       # create_grid!, functions.jl:364-370
    if np.any(st != 0):
        raise RuntimeError(f"create_grid failed: status {st[st != 0][:4]}")
    for m, p in enumerate(pars_list):
        n1, n2 = p["res"]
        if p["init"] == "uniform":
            a["C_left"][m, :n1] = p["Cl_i"]; a["C_right"][m, :n2] = p["Cr_i"]
        elif p["init"] == "b1":                     # examples/B1.jl:54-56
            a["C_left"][m, :n1] = p["Cl_i"]; a["C_right"][m, :n2] = p["Cl_i"]; a["C_right"][m, n2 - 1] = p["Cr_i"]
        elif p["init"] == "mb_growth":              # examples/Diff_couple_MB_growth.jl:55-56
            a["C_left"][m, :n1] = p["Cl_i"]
            a["C_right"][m, :n2] = p["Cr_i"] * p["Cl_i"] * 1.0 * (1.0 / a["KD_ar"][m, 0])
        elif p["init"] == "dcni":                   # examples/Diff_couple_no_interaction.jl:52-59
            Ri = a["Ri"][m]
            a["C_left"][m, :n1] = sinusoid_profile(0.0, DCNI_MODES_L[0], Ri[0], a["Di"][m, 0], 0.0, DCNI_MODES_L[1], a["x_left"][m, :n1])
            a["C_right"][m, :n2] = sinusoid_profile(0.0, DCNI_MODES_R[0], Ri[1] - Ri[0], a["Di"][m, 1], 0.0, DCNI_MODES_R[1], a["x_right"][m, :n2])
        else:
            raise ValueError(p["init"])
        kind = p["kind"]
        if kind == "mb" and (p["BCout"][0] != 0 or p["BCout"][1] != 0):
            raise ValueError("The code is only valid for a closed system. Please set outer BC to Neumann conditions (0).")
        if kind != "mb" and p["BCout"][0] != 0 and p["n"] in (2, 3):
            raise ValueError("The code is only valid for cylindrical/spherical geometry, where the left outer BC has Neumann conditions (0).")
    a["Mass0"][:] = ens.calc_mass(lib)              # calc_mass_vol, functions.jl:205-212
    aux["x0"] = [np.concatenate([ens.member(m)[0], ens.member(m)[2]]).copy() for m in range(M)]
    aux["C0"] = [np.concatenate([ens.member(m)[1], ens.member(m)[3]]).copy() for m in range(M)]
    aux["C0_r"] = [ens.member(m)[3].copy() for m in range(M)]
    aux["KD0"] = a["KD_ar"][:, 0].copy()
    return ens, aux


def run_couple(pars, lib=None, flags=0, max_steps=0, record=False, rec_cap=0):
    """Single-member convenience: returns a dict with everything the reference drivers return."""
    lib = _lib(lib)
    ens, aux = build_couple_ensemble([pars], lib)
    if record and rec_cap == 0:
        rec_cap = 400000
    rc = ens.run(lib, flags=flags | abi.FLAG_MASS_FINAL, max_steps=max_steps, rec_cap=rec_cap if record else 0)
    if rc != 0:
        raise RuntimeError(f"library call failed rc={rc}")
    st = int(ens.status[0])
    if st != 0:
        raise RuntimeError(lib.mbm_status_string(st).decode())
    xl, cl, xr, cr = ens.member(0)
    Lsc = aux["Lsc"][0] if pars["scale"] else 1.0
    out = dict(x_left=xl * Lsc, x_right=xr * Lsc, dx1=float(ens.dx1[0]), dx2=float(ens.dx2[0]), x0=aux["x0"][0] * Lsc,
               res=ens.res[0].copy(), Ri=ens.Ri[0] * Lsc, Ri0=aux["Ri0"][0] * Lsc, C_left=cl.copy(), C_right=cr.copy(),
               C0=aux["C0"][0], C0_r=aux["C0_r"][0], KD0=float(aux["KD0"][0]), n=pars["n"], it=int(ens.it[0]),
               t=float(ens.t[0]) * (aux["tsc"][0] if pars["scale"] else 1.0), Mass0=float(ens.Mass0[0]),
               Mass_final=float(ens.Mass[0]), maxC=max(cl.max(), cr.max()), idx_hash=int(ens.idx_hash[0]),
               node_steps=int(ens.node_steps[0]), last=ens.last[0].copy(), ens=ens, aux=aux)
    if record:
        out["Mass"] = ens.rec[0, : out["it"], 0].copy()
    return out


def DCGMB(lib=None, **over):
    """main_codes/Diff_couple_1D_general_MB.jl:4-140 -> (x_left, x_right, dx1, dx2, x0, res, Ri, C_left, C_right, C0)"""
    p = pars_DCGMB(); p.update(over)
    o = run_couple(p, lib)
    return o["x_left"], o["x_right"], o["dx1"], o["dx2"], o["x0"], o["res"], o["Ri"], o["C_left"], o["C_right"], o["C0"]


def DCGF(lib=None, **over):
    """main_codes/Diff_couple_1D_general_Flux.jl:4-138 -> (..., C0, Mass, Mass0)"""
    p = pars_DCGF(); p.update(over)
    o = run_couple(p, lib, record=True)
    return (o["x_left"], o["x_right"], o["dx1"], o["dx2"], o["x0"], o["res"], o["Ri"], o["C_left"], o["C_right"],
            o["C0"], o["Mass"], o["Mass0"])


def B1(lib=None, **over):
    """examples/B1.jl -> (x_left, x_right, x0, C_left, C_right, C0, t, Di, maxC, Ri)"""
    p = pars_B1(); p.update(over)
    o = run_couple(p, lib)
    return o["x_left"], o["x_right"], o["x0"], o["C_left"], o["C_right"], o["C0"], o["t"], np.array(p["Di"]), o["maxC"], o["Ri"]


def B2(lib=None, **over):
    """examples/B2.jl:139 -> (x_left, x_right, x0, C_left, C_right, C0, t, Di, D0, T_ar[1], T, Ea1, R, Ri, maxC)"""
    p = pars_B2(); p.update(over)
    o = run_couple(p, lib)
    return (o["x_left"], o["x_right"], o["x0"], o["C_left"], o["C_right"], o["C0"], o["t"], np.array(p["Di"]),
            np.array(p["D0"]), p["T_ar"][0], float(o["last"][3]), p["Ea1"], p["R"], o["Ri"], o["maxC"])


def B4(lib=None, **over):
    """examples/B4.jl:143 -> (x_left, x_right, x0, Ri, Ri0, C_left, C_right, C0, C0_r, KD0, n, maxC)"""
    p = pars_B4(); p.update(over)
    o = run_couple(p, lib)
    return (o["x_left"], o["x_right"], o["x0"], o["Ri"], o["Ri0"], o["C_left"], o["C_right"], o["C0"], o["C0_r"],
            o["KD0"], o["n"], o["maxC"])


def B5(lib=None, **over):
    """examples/B5.jl:143 -> (..., maxC, Di, V_ip, t_tot) (Di, V_ip, t_tot rescaled to input units)"""
    p = pars_B5(); p.update(over)
    o = run_couple(p, lib)
    return (o["x_left"], o["x_right"], o["x0"], o["Ri"], o["Ri0"], o["C_left"], o["C_right"], o["C0"], o["C0_r"],
            o["KD0"], o["n"], o["maxC"], np.array(p["Di"]), p["V_ip"], p["t_tot"])


def C1(lib=None, **over):
    """examples/C1.jl -> (x_left, x_right, x0, Ri, Ri0, C_left, C_right, C0, C0_r, n, maxC, KD0)"""
    p = pars_C1(); p.update(over)
    o = run_couple(p, lib)
    return (o["x_left"], o["x_right"], o["x0"], o["Ri"], o["Ri0"], o["C_left"], o["C_right"], o["C0"], o["C0_r"],
            o["n"], o["maxC"], o["KD0"])


def B6(lib=None, **over):
    """examples/B6.jl:142 -> (x_left, x_right, x0, Ri, Ri0, C_left, C_right, C0, C0_r, KD0, n, maxC)"""
    p = pars_B6(); p.update(over)
    o = run_couple(p, lib)
    return (o["x_left"], o["x_right"], o["x0"], o["Ri"], o["Ri0"], o["C_left"], o["C_right"], o["C0"], o["C0_r"],
            o["KD0"], o["n"], o["maxC"])


def B7(lib=None, **over):
    """examples/B7.jl:142 -> B6's tuple (as shipped ~3e5 steps: pass max_steps through run_couple for a prefix)"""
    p = pars_B7(); p.update(over)
    o = run_couple(p, lib)
    return (o["x_left"], o["x_right"], o["x0"], o["Ri"], o["Ri0"], o["C_left"], o["C_right"], o["C0"], o["C0_r"],
            o["KD0"], o["n"], o["maxC"])


def C2(lib=None, **over):
    """examples/C2.jl:144 -> (x_left, x_right, x0, Ri, Ri0, C_left, C_right, C0, C0_r, n, maxC, KD0)"""
    p = pars_C2(); p.update(over)
    o = run_couple(p, lib)
    return (o["x_left"], o["x_right"], o["x0"], o["Ri"], o["Ri0"], o["C_left"], o["C_right"], o["C0"], o["C0_r"],
            o["n"], o["maxC"], o["KD0"])


def _couple_12(p, lib):
    o = run_couple(p, lib, record=True)
    return (o["x_left"], o["x_right"], o["dx1"], o["dx2"], o["x0"], o["res"], o["Ri"], o["C_left"], o["C_right"],
            o["C0"], o["Mass"], o["Mass0"])


def DC_Flux_growth(lib=None, **over):
    """examples/Diff_couple_Flux_growth.jl:138 -> DCGF's 12-tuple"""
    p = pars_Flux_growth(); p.update(over)
    return _couple_12(p, lib)


def DC_MB_growth(lib=None, **over):
    """examples/Diff_couple_MB_growth.jl -> (x_left, x_right, dx1, dx2, x0, res, Ri, C_left, C_right, C0, Mass, Mass0)"""
    p = pars_MB_growth(); p.update(over)
    return _couple_12(p, lib)


def DC_Flux(lib=None, **over):
    """examples/Diff_couple_Flux.jl:130 -> (x_left, x_right, dx1, dx2, x0, res, Ri, C_left, C_right, C0)"""
    p = pars_DC_Flux(); p.update(over)
    return _couple_12(p, lib)[:10]


def DC_MB(lib=None, **over):
    """examples/Diff_couple_MB.jl:132 -> DCGF's 12-tuple"""
    p = pars_DC_MB(); p.update(over)
    return _couple_12(p, lib)


def DCNI(lib=None, **over):
    """examples/Diff_couple_no_interaction.jl -> (x_left, x_right, x0, C_left, C_right, C0, Cini_l, Cini_r,
    nmodes_l, nmodes_r, Amp_l, Amp_r, Di, Ri, t)"""
    p = pars_DCNI(); p.update(over)
    o = run_couple(p, lib)
    return (o["x_left"], o["x_right"], o["x0"], o["C_left"], o["C_right"], o["C0"], 0.0, 0.0, DCNI_MODES_L[0],
            DCNI_MODES_R[0], DCNI_MODES_L[1], DCNI_MODES_R[1], np.array(p["Di"]), o["Ri"], o["t"])


# ------------------------------------------------------------------------------------------------
# single phase: main_codes/Simple_Diff_1D_general.jl, examples/A1.jl, A2.jl
# ------------------------------------------------------------------------------------------------
def pars_SDG():
    """main_codes/Simple_Diff_1D_general.jl:6-31"""
    return dict(Di=2.65e-18, D0=2.75e-6, L=0.001, rho=2700.0, R=8.314472, Ea1=292880.0, t_tot=1e-4 * Myr2Sec, n=1,
                res=100, CFL=0.99, BCout=[1, 1], T_ar=(1273.15, 923.15), init="sdg", Cinf=0.0, Cstart=4.0)


def pars_Simple_Diff():
    """examples/Simple_Diff.jl:7-36: SDG's column with the Arrhenius diffusivity (Di = -1) on a cooling path"""
    p = pars_SDG()
    p.update(Di=-1.0)
    return p


def pars_A1():
    """examples/A1.jl:6-36"""
    return dict(Di=2.65e-18, D0=np.nan, L=0.005, rho=2700.0, R=np.nan, Ea1=np.nan, t_tot=1e-3 * Myr2Sec, n=1,
                res=500, CFL=0.99, BCout=[1, 1], T_ar=(1273.15, 1273.15), init="a1",
                nmodes=np.array([1, 2, 5, 7, 12.0]), Amp=np.array([12, 0.5, 3, -2, 1.0]))


def pars_A2():
    """examples/A2.jl:6-34"""
    return dict(Di=2.65e-18, D0=np.nan, L=0.001, rho=2700.0, R=np.nan, Ea1=np.nan, t_tot=1e-3 * Myr2Sec, n=3,
                res=500, CFL=0.5, BCout=[0, 1], T_ar=(1273.15, 1273.15), init="a2", Cinf=1.0, Cstart=0.0)


def build_simple_ensemble(pars_list, hist_len=1000, cap=None):
    M = len(pars_list)
    p0 = pars_list[0]
    cap = cap or max(p["res"] for p in pars_list)
    ens = SimpleEnsemble(M, int(p0["n"]), cap, hist_len, R=float(p0["R"]),
                         arrays=dict(t_ar=np.zeros((M, hist_len)), T_ar=np.zeros((M, hist_len))))
    a = ens.arrays
    for m, p in enumerate(pars_list):
        res, L = p["res"], p["L"]
        dx = L * (1.0 / (res - 1))                         # SDG :22
        x = steprange(dx, L)                               # :23
        assert len(x) == res, (len(x), res)
        a["x"][m, :res] = x
        if p["init"] == "sdg":
            C = p["Cinf"] * np.ones(res); C[0] = p["Cstart"]
        elif p["init"] == "a1":
            C = sinusoid_profile(0.0, p["nmodes"], L, p["Di"], 0.0, p["Amp"], x)
        elif p["init"] == "a2":
            C = p["Cstart"] * np.ones(res); C[-1] = p["Cinf"]
        a["C"][m, :res] = C
        a["res"][m] = res
        a["Di"][m], a["D0"][m], a["Ea1"][m], a["rho"][m] = p["Di"], p["D0"], p["Ea1"], p["rho"]
        a["dx"][m] = L * (1.0 / (res - 1.0))               # :45
        a["t_tot"][m], a["CFL"][m] = p["t_tot"], p["CFL"]
        a["BCout"][m] = p["BCout"]
        a["t_ar"][m] = linrange(0.0, p["t_tot"], hist_len)
        a["T_ar"][m] = linrange(p["T_ar"][0], p["T_ar"][1], hist_len)
        if p["BCout"][0] != 0 and p["n"] in (2, 3):
            raise ValueError("The code is only valid for cylindrical/spherical geometry, where the left outer BC has Neumann conditions (0).")
    return ens


def run_simple(pars, lib=None, flags=0):
    lib = _lib(lib)
    ens = build_simple_ensemble([pars])
    x0, C0 = ens.x[0, : pars["res"]].copy(), ens.C[0, : pars["res"]].copy()
    rc = ens.run(lib, flags=flags | abi.FLAG_MASS_FINAL)
    if rc != 0:
        raise RuntimeError(f"library call failed rc={rc}")
    if ens.status[0] != 0:
        raise RuntimeError(lib.mbm_status_string(int(ens.status[0])).decode())
    n = pars["res"]
    return dict(x=ens.x[0, :n].copy(), C=ens.C[0, :n].copy(), x0=x0, C0=C0, D=float(ens.D[0]), t=float(ens.t[0]),
                t_tot=pars["t_tot"], Mass=float(ens.Mass[0]), it=int(ens.it[0]), ens=ens)


def SDG(lib=None, **over):
    """main_codes/Simple_Diff_1D_general.jl:4-103 -> (x, C, x0, C0, D, t, t_tot, Mass, Mass0)"""
    p = pars_SDG(); p.update(over)
    ens0 = build_simple_ensemble([p])
    o = run_simple(p, lib)
    # Mass0 = calc_mass_vol_simple_diff at t = 0 (SDG :54): same trapezoid on the initial profile
    x0, C0 = o["x0"], o["C0"]
    V = p["rho"] * x0 ** p["n"]
    Mass0 = float(np.cumsum(0.5 * np.diff(V) * (C0[:-1] + C0[1:]))[-1])
    del ens0
    return o["x"], o["C"], o["x0"], o["C0"], o["D"], o["t"], o["t_tot"], o["Mass"], Mass0


def Simple_Diff(lib=None, **over):
    """examples/Simple_Diff.jl:107 -> (x, C, x0, C0, D, t, t_tot)"""
    p = pars_Simple_Diff(); p.update(over)
    o = run_simple(p, lib)
    return o["x"], o["C"], o["x0"], o["C0"], o["D"], o["t"], o["t_tot"]


def A1(lib=None, **over):
    """examples/A1.jl -> (x, C, x0, C0, D, t, t_tot, Cini, nmodes, Amp, L)"""
    p = pars_A1(); p.update(over)
    o = run_simple(p, lib)
    return o["x"], o["C"], o["x0"], o["C0"], o["D"], o["t"], o["t_tot"], 0.0, p["nmodes"], p["Amp"], p["L"]


def A2(lib=None, **over):
    """examples/A2.jl -> (x, C, x0, C0, D, t, t_tot)"""
    p = pars_A2(); p.update(over)
    o = run_simple(p, lib)
    return o["x"], o["C"], o["x0"], o["C0"], o["D"], o["t"], o["t_tot"]


# ------------------------------------------------------------------------------------------------
# Stefan: main_codes/Chemical_Stefan_problem_XT.jl
# ------------------------------------------------------------------------------------------------
COEFF_REACTION_LINES = np.array([  # examples/Examples_phase_diagram/Coefficients_Reaction_lines.csv (3 x 2 input data)
    [0.8065319078911909, 9.79427908555563],
    [0.0010185823804784945, -0.010021212961038951],
    [-6.464640891194833e-7, 2.6061746209320835e-6],
])


def pars_CSP():
    """main_codes/Chemical_Stefan_problem_XT.jl:8-26 (density table: the shipped .tab is a missing blob;
    rho = [1 1] as examples/D1.jl:54-55 does)"""
    return dict(Di=[0.002, 0.0004], Ri=2.0, Tstart=1400.0 + 273.0, Tstop=1300.0 + 273.0, t_tot=1e4, n=1, CompInt=0.3,
                coeff=COEFF_REACTION_LINES, CFL=0.8, res=[15, 25], MRefin=5.0, BCout=[0, 0], hist_len=10000)


def pars_D1():
    """examples/D1.jl:8-35: olivine / melt Stefan problem with in-loop diffusivities (olivine after Dohmen &
    Chakraborty 2007 as a function of T and the mean olivine composition, melt after Zhang & Cherniak 2010);
    the density table is replaced by ones exactly as D1.jl:54-55 does.  Angles are passed to cos() as written
    in the reference (90.0 is taken in radians there)."""
    return dict(D0=5.38 * 1e-9, Ea=226000, alpha=0.0, beta=90.0, gamma=90.0, deltaV=7 * 0.1**6, P=10**6, R=8.314472,
                Ri=0.0001, Tstart=1400.0 + 273.0, Tstop=1350.0 + 273.0, t_tot=60 * 60 * 24 * 30, n=1, CompInt=0.50,
                coeff=COEFF_REACTION_LINES, CFL=50.0, res=[50, 50], MRefin=5.0, BCout=[0, 0], hist_len=10000,
                law=abi.LAW_OLIVINE_MELT_D1)


def d1_diffusivities(p, T, meanC):
    """examples/D1.jl:68-74 / :126-132 (host side: initial values, used for the first dt only)"""
    import math
    l001 = math.log10(p["D0"]) - (p["Ea"] + (p["P"] - 10**5) * p["deltaV"]) / (2.303 * p["R"] * T) + 3 * (meanC - 0.14)
    loth = l001 - math.log10(6.0)
    D001, Doth = 10**l001, 10**loth
    D_l = D001 * math.cos(p["alpha"]) ** 2 + Doth * math.cos(p["beta"]) ** 2 + Doth * math.cos(p["gamma"]) ** 2
    D_r = math.exp(-7.92 - 26222 / T)
    return D_l, D_r


def build_stefan_ensemble(pars_list, lib=None, cap_left=None, cap_right=None):
    """Pre-loop part of CSP() (Chemical_Stefan_problem_XT.jl:27-98) and of D1() (examples/D1.jl:36-97) for a list
    of members; members with a `law` entry carry the in-loop material law of D1."""
    lib = _lib(lib)
    M = len(pars_list)
    p0 = pars_list[0]
    hist_len = p0["hist_len"]
    res_max = np.max([p["res"] for p in pars_list], axis=0)
    cap_left = cap_left or int(4 * res_max[0] + 8)
    cap_right = cap_right or int(4 * res_max[1] + 8)
    law = int(p0.get("law", abi.LAW_NONE))
    assert all(int(p.get("law", abi.LAW_NONE)) == law for p in pars_list), "all members must share the material law"
    ens = CoupleEnsemble(M, int(p0["n"]), abi.BC_STEFAN, abi.DT_STEFAN, cap_left, cap_right, hist_len, R=8.314472,
                         arrays=dict(t_ar=np.zeros((M, hist_len)), T_ar=np.zeros((M, hist_len)),
                                     KD_ar=np.zeros((1, hist_len)), coeff_up=np.zeros((M, 3)), coeff_do=np.zeros((M, 3)),
                                     C_leftB=np.zeros(M), C_rightB=np.zeros(M),
                                     law_par=np.zeros((M, 8)) if law else None))
    a = ens.arrays
    if law:
        a["law"] = law
    for m, p in enumerate(pars_list):
        coeff = np.asarray(p["coeff"], float)
        coeff_up, coeff_do = coeff[:, 0].copy(), coeff[:, 1].copy()          # :17-19, coeff_trans_line
        T = p["Tstart"]
        CLB = coeff_do[0] + coeff_do[1] * T + coeff_do[2] * T**2              # composition :67-71
        CRB = coeff_up[0] + coeff_up[1] * T + coeff_up[2] * T**2
        n, Ri0, CompInt = p["n"], p["Ri"], p["CompInt"]
        Xc = (1.0 - CompInt) * CRB + CompInt * CLB                             # :58
        L = (Ri0**n * (Xc - CLB) * (1.0 / (CRB - Xc)) ** (1 * (1.0 / n))) + Ri0  # :59
        if Ri0 >= L:
            raise ValueError("Please change the size of the system. Increase Ri[2].")
        a["coeff_up"][m], a["coeff_do"][m], a["C_leftB"][m], a["C_rightB"][m] = coeff_up, coeff_do, CLB, CRB
        a["Ri"][m] = [Ri0, L]
        a["res"][m] = p["res"]; a["resmin"][m] = p["res"]
        if law:                                                                # D1.jl:68-74: D(T, X) at Tstart
            import math
            p["Di"] = list(d1_diffusivities(p, T, CLB))
            a["law_par"][m, :5] = [math.log10(p["D0"]), p["Ea"] + (p["P"] - 10**5) * p["deltaV"],
                                   math.cos(p["alpha"]) ** 2, math.cos(p["beta"]) ** 2, math.cos(p["gamma"]) ** 2]
        a["Di"][m] = p["Di"]; a["rho"][m] = [1.0, 1.0]
        a["t_tot"][m], a["CFL"][m], a["MRefin"][m], a["BCout"][m] = p["t_tot"], p["CFL"], p["MRefin"], p["BCout"]
        a["T_ar"][m] = linrange(p["Tstart"], p["Tstop"], hist_len)             # Tpath :51
        a["t_ar"][m] = linrange(0.0, p["t_tot"] + 1e-10, hist_len)             # tpath :52
    st = ens.create_grid(lib)
    if np.any(st != 0):
        raise RuntimeError("create_grid failed")
    for m, p in enumerate(pars_list):
        n1, n2 = p["res"]
        a["C_left"][m, :n1] = a["C_leftB"][m]; a["C_right"][m, :n2] = a["C_rightB"][m]      # :76-77
        mn = min(a["dx1"][m], a["dx2"][m])
        a["dt"][m] = mn**2 * (1.0 / max(p["Di"]))                                            # :79
    a["Mass0"][:] = ens.calc_mass(lib)
    return ens


def D1(lib=None, **over):
    """examples/D1.jl:4-233 (the entries of its 25-tuple that do not need plots): the Stefan driver with the
    olivine / melt diffusivities recomputed every step -> dict(x_left, x_right, x0, C_left, C_right, C0, maxC, T,
    Mass0, Mass, KD_sim, T_sim, Ri, it)"""
    p = pars_D1(); p.update(over)
    out = CSP(lib, _pars=p)
    out["maxC"] = float(max(out["C_left"].max(), out["C_right"].max()))
    return out


def CSP(lib=None, **over):
    """main_codes/Chemical_Stefan_problem_XT.jl:4-208 (subset of the 20-tuple that does not need plots):
    -> dict(x_left, x_right, x0, C_left, C_right, C0, T, Mass0, Mass, KD_sim, T_sim, Ri, it)"""
    lib = _lib(lib)
    p = over.pop("_pars", None) or pars_CSP()
    p.update(over)
    ens = build_stefan_ensemble([p], lib)
    x0 = np.concatenate([ens.member(0)[0], ens.member(0)[2]]).copy()
    C0 = np.concatenate([ens.member(0)[1], ens.member(0)[3]]).copy()
    rec_cap = int(over.get("rec_cap", 200000))
    rc = ens.run(lib, flags=abi.FLAG_MASS_FINAL, rec_cap=rec_cap)
    if rc != 0:
        raise RuntimeError(f"library call failed rc={rc}")
    if ens.status[0] != 0:
        raise RuntimeError(lib.mbm_status_string(int(ens.status[0])).decode())
    xl, cl, xr, cr = ens.member(0)
    it = int(ens.it[0])
    rec = ens.rec[0, : min(it, rec_cap)]
    return dict(x_left=xl.copy(), x_right=xr.copy(), x0=x0, C_left=cl.copy(), C_right=cr.copy(), C0=C0,
                T=float(ens.last[0, 3]), Mass0=float(ens.Mass0[0]), Mass=rec[:, 0].copy(), T_sim=rec[:, 1].copy(),
                KD_sim=rec[:, 2].copy(), Ri_sim=rec[:, 3].copy(), Ri=ens.Ri[0].copy(), it=it, res=ens.res[0].copy(), ens=ens)
