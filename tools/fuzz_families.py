#!/usr/bin/env python
"""Randomised parity sweep over the families the first sweep (fuzz_parity.py) does not reach: the advective dt rule of
examples/C1.jl / B4.jl / B5.jl, fixed-interface couples (flux / mass balance / Dirichlet hold / Lasaga rows) with constant
and Arrhenius diffusivities, the Stefan driver and single-phase diffusion, at random sizes, geometries and step counts.
CUDA library against the CPU oracle on identical inputs: integers, status codes (and the regrid checksum where the
meshes are bit-identical) equal; profiles within the tolerance of the family (tests/test_gpu_parity.py).
Usage: fuzz_families.py [seed] [batches]"""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import mbm_b200
from mbm_b200 import abi, drivers as d, ensembles as E
from oracle import load as oracle_load

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
batches = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rng = np.random.default_rng(seed)
ora = oracle_load.load()
lib = ora if os.environ.get("FUZZ_ORACLE_ONLY") else mbm_b200.load_cuda_library()


def relerr(a, b):
    den = np.max(np.abs(a))
    return 0.0 if den == 0 else float(np.max(np.abs(a - b)) / den)


def couple_batch():
    fam = str(rng.choice(["c1", "b4", "b5", "b1", "b2", "mbfixed", "dcni", "lasaga"]))
    M = int(rng.integers(1, 5))
    tol, hashed = 1e-10, True
    pars = []
    for _ in range(M):
        if fam in ("c1", "b4", "b5"):
            p = {"c1": d.pars_C1, "b4": d.pars_B4, "b5": d.pars_B5}[fam]()
            kd0 = float(np.exp(rng.uniform(np.log(5.0), np.log(300.0))))
            r1 = int(rng.integers(30, 200)); r2 = int(rng.integers(r1, 260))
            p.update(res=[r1, r2], CFL=float(rng.uniform(0.2, 0.8)), KD_ar=(kd0, kd0), Cr_i=p["Cl_i"] / kd0,
                     Ri=[float(rng.uniform(0.005, 0.03)) * (1.0 if fam != "b5" else 1.0), p["Ri"][1]])
        elif fam in ("b1", "b2"):
            p = d.pars_B1() if fam == "b1" else d.pars_B2()
            r1 = int(rng.integers(20, 300)); r2 = int(rng.integers(r1, 320))
            p.update(res=[r1, r2], MRefin=float(rng.choice([1.0, 5.0, 15.0])), n=int(rng.integers(1, 4)),
                     KD_ar=(1.0, float(rng.uniform(0.6, 1.4))))
            p.update(BCout=[0, int(rng.integers(0, 2))] if p["n"] != 1 else [int(rng.integers(0, 2)), int(rng.integers(0, 2))])
        elif fam == "mbfixed":
            p = d.pars_MB_growth()
            r1 = int(rng.integers(30, 300)); r2 = int(rng.integers(r1, 330))
            p.update(V_ip=0.0, res=[r1, r2], n=int(rng.integers(1, 4)), MRefin=float(rng.choice([1.0, 8.0, 20.0])),
                     KD_ar=(0.1, float(rng.uniform(0.05, 0.2))))
        elif fam == "dcni":
            p = d.pars_DCNI()
            r = int(rng.integers(50, 600))
            p.update(res=[r, r])
        else:
            p = d.pars_DCGF()
            r1 = int(rng.integers(30, 200)); r2 = int(rng.integers(r1, 260))
            p.update(kind="lasaga", V_ip=float(rng.choice([0.0, 1e-8, -1e-8])), Di=[4e-8, 5e-8], Cl_i=float(rng.uniform(0.3, 0.7)),
                     Cr_i=float(rng.uniform(0.3, 0.7)), KD_ar=(float(rng.uniform(0.6, 2.0)), float(rng.uniform(0.6, 2.0))),
                     res=[r1, r2], MRefin=5.0, n=int(rng.integers(1, 4)), t_tot=4e5, rho=[1.0, float(rng.uniform(0.9, 1.2))])
        pars.append(p)
    # members of one call share kind / n / dt_mode
    for p in pars:
        p.update(n=pars[0]["n"], BCout=pars[0]["BCout"] if fam in ("b1", "b2") else p["BCout"])
    if fam == "b2":
        tol = 2e-9                                          # CUDA exp vs glibc exp: 1-ulp D differences feed dt
    if fam == "lasaga":
        tol = 1e-9       # the shipped closed form divides by 2 Dr dx1 rho2 (1 - KD): ill-conditioned where KD(t) passes near 1
                         # (12 000 batches of all families: one Lasaga batch at 1.5e-10, the prescribed values themselves
                         # are computed with identical operations on both sides)
    eo, _ = d.build_couple_ensemble(pars, ora)
    eg, _ = d.build_couple_ensemble(pars, lib)
    return fam, pars, eo, eg, tol, hashed


def simple_batch():
    """single-phase diffusion (Simple_Diff_1D_general / A1 / A2): random size, geometry, outer boundary conditions, CFL,
    constant or Arrhenius diffusivity"""
    M = int(rng.integers(1, 5))
    fam = str(rng.choice(["sdg", "a1", "a2"]))
    n = int(rng.integers(1, 4))
    pars = []
    for _ in range(M):
        p = {"sdg": d.pars_SDG, "a1": d.pars_A1, "a2": d.pars_A2}[fam]()
        bc = [int(rng.integers(0, 2)), int(rng.integers(0, 2))] if n == 1 else [0, int(rng.integers(0, 2))]
        p.update(res=int(rng.integers(12, 3000)), n=n, BCout=bc, CFL=float(rng.uniform(0.3, 0.99)))
        if fam != "sdg" and rng.random() < 0.3:            # Arrhenius diffusivity (Di == -1, functions.jl:1414-1417)
            p.update(Di=-1.0, D0=2.75e-6, R=8.314472, Ea1=292880.0, T_ar=(1273.15, 1073.15))
        pars.append(p)
    arrh = any(p["Di"] == -1.0 for p in pars)
    for p in pars:                                          # one call shares the gas constant
        p.update(R=8.314472 if arrh else p["R"])
    eo = d.build_simple_ensemble(pars)
    return "simple-" + fam, pars, eo, eo.copy(), (2e-9 if arrh else 1e-10), False


def stefan_batch():
    M = int(rng.integers(1, 5))
    pars = E.c2b_stefan_sweep(M, rank=int(rng.integers(0, 10_000)))
    eo, eg = d.build_stefan_ensemble(pars, ora), d.build_stefan_ensemble(pars, lib)
    return "stefan", pars, eo, eg, 1e-9, False              # V comes from the solution; meshes differ in the last bits


bad = 0
t0 = time.time()
for b in range(batches):
    try:
        u_ = rng.random()
        fam, pars, eo, eg, tol, hashed = stefan_batch() if u_ < 0.2 else simple_batch() if u_ < 0.4 else couple_batch()
    except (ValueError, RuntimeError) as e:
        print("skip batch", b, str(e)[:80]); continue
    steps = int(rng.integers(20, 200))
    assert eo.run(ora, max_steps=steps, flags=abi.FLAG_MASS_FINAL) == 0
    if rng.random() < 0.4:
        a = steps // 2
        assert eg.run(lib, max_steps=a, flags=abi.FLAG_MASS_FINAL) == 0 and eg.run(lib, max_steps=steps - a, flags=abi.FLAG_MASS_FINAL) == 0
    else:
        assert eg.run(lib, max_steps=steps, flags=abi.FLAG_MASS_FINAL) == 0
    ok = bool(np.array_equal(eo.status, eg.status))
    worst = 0.0
    for m in range(eo.M):
        if eo.status[m] != 0:
            continue
        ok &= bool(np.array_equal(eo.res[m], eg.res[m]) and eo.it[m] == eg.it[m] and eo.node_steps[m] == eg.node_steps[m])
        if fam.startswith("simple"):
            nn_ = int(eo.res[m])
            worst = max(worst, relerr(eo.C[m, :nn_], eg.C[m, :nn_]))
            continue
        if hashed and fam != "b2":
            ok &= bool(eo.arrays["idx_hash"][m] == eg.arrays["idx_hash"][m])
        xl_o, cl_o, xr_o, cr_o = eo.member(m)
        xl_g, cl_g, xr_g, cr_g = eg.member(m)
        # concentrations are compared on the couple's own scale (a phase the diffusion front has not reached holds rounding
        # noise of 1e-100 and less: relative to ITS maximum every solver differs by O(1), the oracle's two included)
        cscale = max(np.max(np.abs(cl_o)), np.max(np.abs(cr_o)))
        worst = max(worst, relerr(xl_o, xl_g), relerr(xr_o, xr_g), float(np.max(np.abs(cl_o - cl_g)) / cscale),
                    float(np.max(np.abs(cr_o - cr_g)) / cscale))
    ok &= worst < tol
    if not ok:
        bad += 1
    print(f"{'ok ' if ok else 'BAD'} batch {b}: {fam} M={eo.M} steps={steps} res={eo.res.tolist()} status={eo.status.tolist()}/{eg.status.tolist()} "
          f"it={eo.it.tolist()} worst={worst:.2e} (tol {tol:.0e})", flush=True)
print(f"fuzz families seed {seed}: {batches} batches, {bad} bad, {time.time()-t0:.0f} s")
sys.exit(1 if bad else 0)
