"""Synthetic ensembles of the BASELINE.json configurations: lists of per-member
parameter blocks in the reference's own conventions, sampled with numpy PCG64 (seed = 20251019 + config#).
Host-side only; the members are advanced by the C-ABI library."""
from __future__ import annotations

import numpy as np

from . import abi
from .drivers import pars_B1, pars_C1, pars_CSP, pars_MB_growth

SEED0 = 20251019


def _rng(cfg: int, extra: int = 0):
    return np.random.Generator(np.random.PCG64(SEED0 + cfg + 1000003 * extra))


def _logu(rng, lo, hi, n):
    return np.exp(rng.uniform(np.log(lo), np.log(hi), n))


def c4_mb_growth_resorption(M: int, res: int = 1024, rank: int = 0, max_steps: int = 10_000, mrefin: float = 20.0):
    """config 4: MB growth/resorption couples (examples/Diff_couple_MB_growth.jl:8-34 as base), half growth,
    half resorption, n = 2, MRefin = 20, CFL = 0.3, KD constant or linearly drifting, linear T path.
    t_tot is set to 1.5x the time `max_steps` steps need, so no member finishes early."""
    rng = _rng(4, rank)
    Dl, Dr = _logu(rng, 2e-7, 8e-7, M), _logu(rng, 2.5e-8, 1e-7, M)
    KD0 = rng.uniform(0.05, 0.2, M)
    drift = np.where(rng.uniform(size=M) < 0.5, 1.0, rng.uniform(0.5, 1.0, M))
    Vmag = rng.uniform(0.5e-7, 2e-7, M)
    growth = (np.arange(M) % 2) == 0
    r = np.where(growth, rng.uniform(0.4, 0.6, M), rng.uniform(0.30, 0.45, M))
    Tend = rng.uniform(900.0, 1000.0, M)
    pars = []
    for m in range(M):
        p = pars_MB_growth()
        V = Vmag[m] if growth[m] else -Vmag[m]
        # estimate of the first dt in seconds (find_dt, functions.jl:465-490) from the graded left mesh
        Lsc, Dsc = 1e-3, 0.5 * (Dl[m] + Dr[m])
        dx = (r[m] / Lsc) * (1.0 / mrefin) / ((res - 2) * (1.0 + 1.0 / mrefin) / 2.0)
        dt0 = 0.3 * min(0.99 * dx / (abs(V) / (Dsc / Lsc)), 50.0 * dx * dx / (max(Dl[m], Dr[m]) / Dsc))
        t_tot = 1.5 * max_steps * dt0 * (Lsc * Lsc / Dsc)
        p.update(Di=[Dl[m], Dr[m]], Ri=[r[m], 1.0], V_ip=V, t_tot=t_tot, KD_ar=(KD0[m], KD0[m] * drift[m]),
                 T_ar=(1273.15, Tend[m]), res=[res, res], MRefin=mrefin, Cl_i=0.5, Cr_i=0.5 / KD0[m], init="uniform")
        pars.append(p)
    return pars


def c3_flux_fixed(M: int, res: int = 256, rank: int = 0):
    """config 3: B1-style fixed-interface flux couples (examples/B1.jl:7-33), sphere, V = 0, BCout [0 1]"""
    rng = _rng(3, rank)
    D = _logu(rng, 1e-18, 1e-17, M)
    r = rng.uniform(2.5e-4, 7.5e-4, M)
    pars = []
    for m in range(M):
        p = pars_B1()
        p.update(Di=[D[m], D[m]], Ri=[r[m], 2 * r[m]], res=[res, res], t_tot=1e30)
        pars.append(p)
    return pars


def c2a_rayleigh_sweep(M: int, rank: int = 0):
    """config 2a: examples/C1.jl sweep (MB couple vs Rayleigh fractionation)"""
    rng = _rng(2, rank)
    KD0 = _logu(rng, 20.0, 200.0, M)
    Ri1 = rng.uniform(0.005, 0.02, M)
    CFL = rng.uniform(0.2, 0.4, M)
    pars = []
    for m in range(M):
        p = pars_C1()
        Cl = 0.5
        p.update(Ri=[Ri1[m], 1.0], CFL=CFL[m], KD_ar=(KD0[m], KD0[m]), Cl_i=Cl, Cr_i=Cl / KD0[m])
        pars.append(p)
    return pars


def c2b_stefan_sweep(M: int, rank: int = 0):
    """config 2b: main_codes/Chemical_Stefan_problem_XT.jl sweep (rho == 1)"""
    rng = _rng(2, rank + 7)
    Dl, Dr = _logu(rng, 1e-3, 4e-3, M), _logu(rng, 2e-4, 8e-4, M)
    Ri = rng.uniform(1.0, 3.0, M)
    Tstop = rng.uniform(1523.0, 1623.0, M)
    CI = rng.uniform(0.2, 0.4, M)
    pars = []
    for m in range(M):
        p = pars_CSP()
        p.update(Di=[Dl[m], Dr[m]], Ri=Ri[m], Tstop=Tstop[m], CompInt=CI[m])
        pars.append(p)
    return pars


def c5_highres(M: int, res: int = 65536, rank: int = 0, max_steps: int = 100):
    """config 5: the c4 family at n = 65,536 per phase with MRefin = 1 (uniform meshes: the only mesh the
    reference can generate at this size), growth members only"""
    rng = _rng(5, rank)
    Dl, Dr = _logu(rng, 2e-7, 8e-7, M), _logu(rng, 2.5e-8, 1e-7, M)
    KD0 = rng.uniform(0.05, 0.2, M)
    Vmag = rng.uniform(0.5e-7, 2e-7, M)
    r = rng.uniform(0.4, 0.5, M)
    pars = []
    for m in range(M):
        p = pars_MB_growth()
        Lsc, Dsc = 1e-3, 0.5 * (Dl[m] + Dr[m])
        dx = (r[m] / Lsc) / (res - 1)
        dt0 = 0.3 * min(0.99 * dx / (Vmag[m] / (Dsc / Lsc)), 50.0 * dx * dx / (max(Dl[m], Dr[m]) / Dsc))
        t_tot = 1.5 * max_steps * dt0 * (Lsc * Lsc / Dsc)
        p.update(Di=[Dl[m], Dr[m]], Ri=[r[m], 1.0], V_ip=Vmag[m], t_tot=t_tot, KD_ar=(KD0[m], KD0[m]),
                 T_ar=(1273.15, 950.0), res=[res, res], MRefin=1.0, Cl_i=0.5, Cr_i=0.5 / KD0[m], init="uniform")
        pars.append(p)
    return pars


FLOPS_PER_NODE_STEP = {  # algorithmic flops, SURVEY.md §8(d)
    ("mb", 1): 86, ("mb", 2): 100, ("mb", 3): 114, ("flux_moving", 1): 73, ("flux_moving", 2): 86,
    ("flux_moving", 3): 99, ("flux_fixed", 3): 50, ("simple", 1): 24, ("simple", 3): 50, ("stefan", 1): 73,
}
BYTES_PER_NODE_STEP_MOVING = 32   # read + write x and C once (streaming path)
BYTES_PER_NODE_STEP_FIXED = 24
