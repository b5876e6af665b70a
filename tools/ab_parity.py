#!/usr/bin/env python
"""A/B helper: for the library in MBM_B200_LIB, time a c4 ensemble (kernel seconds, 2 CTAs per SM) and check its
first members against the CPU oracle (integers exact, profiles to 1e-10).  Usage: ab_parity.py [M] [steps] [n_check]"""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
os.environ.setdefault("MBM_PIECES", "1")
sys.path.insert(0, str(ROOT))
import numpy as np
import mbm_b200
from mbm_b200 import drivers as d, ensembles as E
from oracle import load as oracle_load

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
ncheck = int(sys.argv[3]) if len(sys.argv) > 3 else 32
lib, ora = mbm_b200.load_cuda_library(), oracle_load.load()
pars = E.c4_mb_growth_resorption(M, res=1024, max_steps=10_000)
ONLY = os.environ.get("MBM_AB_ONLY")                 # "growth" / "resorption": one kind of member only (same count)
if ONLY:
    sel = E.c4_mb_growth_resorption(2 * M, res=1024, max_steps=10_000)
    pars = [p for i, p in enumerate(sel) if (i % 2 == 0) == (ONLY == "growth")]
eg, _ = d.build_couple_ensemble(pars, lib, cap_left=2048, cap_right=2560)
eo, _ = d.build_couple_ensemble(pars[:ncheck], ora, cap_left=2048, cap_right=2560)
NOHASH = bool(os.environ.get("MBM_AB_NOHASH"))      # the bench runs without the regrid checksum
if NOHASH:
    eg.arrays["idx_hash"] = None; eo.arrays["idx_hash"] = None
eg.run(lib, max_steps=3)
eo.run(ora, max_steps=3)
best = 1e9
n0 = int(eg.node_steps.sum())
for rep in range(2):
    assert eg.run(lib, max_steps=steps) == 0
    ks = lib.mbm_last_kernel_seconds(); ns = int(eg.node_steps.sum()) - n0; n0 += ns
    best = min(best, ks / ns)
assert eo.run(ora, max_steps=2 * steps) == 0
worst = 0.0
for m in range(ncheck):
    for a, b in zip(eo.member(m), eg.member(m)):
        den = np.max(np.abs(a)); worst = max(worst, float(np.max(np.abs(a - b)) / den) if den else 0.0)
ints = all(np.array_equal(getattr(eo, k), getattr(eg, k)[:ncheck]) for k in ("res", "it", "status", "node_steps") + (() if NOHASH else ("idx_hash",)))
print(f"{os.path.basename(os.environ.get('MBM_B200_LIB', 'default'))}: {1 / best:.4g} node-steps/s  failed {int((eg.status != 0).sum())} "
      f"parity relerr {worst:.2e} ints_exact {ints} launches {lib.mbm_last_kernel_launches()}")
