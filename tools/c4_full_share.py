#!/usr/bin/env python
"""One GPU's share of BASELINE.json configs[3] run to its full length: 8 192 members (65 536 / 8), n = 1024 per phase,
10^4 steps in ONE call through the C ABI with host buffers, then the mass-error gate (calc_mass_err on calc_mass_vol,
functions.jl:183-212) for every member on the device.  Usage: c4_full_share.py [members] [steps]"""
import sys, time, os, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import mbm_b200
from mbm_b200 import abi, drivers as d, ensembles as E

M = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
lib = mbm_b200.load_cuda_library()
t0 = time.perf_counter()
ens, _ = d.build_couple_ensemble(E.c4_mb_growth_resorption(M, res=1024, max_steps=steps), lib, cap_left=2048, cap_right=2560)
ens.arrays["idx_hash"] = None
tb = time.perf_counter() - t0
warm = ens.copy(); warm.run(lib, max_steps=2); del warm
t0 = time.perf_counter()
assert ens.run(lib, max_steps=steps, flags=abi.FLAG_MASS_FINAL) == 0
te = time.perf_counter() - t0
ks = lib.mbm_last_kernel_seconds()
out, err = ens.validate(lib, abi.VAL_MASS_ERR)
ns = int(ens.node_steps.sum())
print(json.dumps({"config": f"configs[3], one GPU's share at full length: {M} members x {steps} steps, n=1024/phase",
                  "members_failed": int((ens.status != 0).sum()), "steps_min_max": [int(ens.it.min()), int(ens.it.max())],
                  "node_steps": ns, "e2e_s": te, "node_steps_per_s_e2e": ns / te,
                  "kernel_s_summed_over_the_8_pieces": ks,
                  "mass_err_max_abs": float(np.max(np.abs(err))), "build_s": tb,
                  "res_left_min_max": [int(ens.res[:, 0].min()), int(ens.res[:, 0].max())],
                  "res_right_min_max": [int(ens.res[:, 1].min()), int(ens.res[:, 1].max())]}))
