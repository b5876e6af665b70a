#!/usr/bin/env python
"""How often does a host-buffer call of the bench workload stall?  Repeats mbm_couple_batch (pinned host buffers, 250 time
steps, the bench's 8 192 members) N times from the same initial state and prints the distribution of the call times.
Usage: e2e_jitter.py [calls]"""
import sys, time, os, ctypes as C
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import mbm_b200
from mbm_b200 import abi, drivers as d, ensembles as E

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
lib = mbm_b200.load_cuda_library()
ens, _ = d.build_couple_ensemble(E.c4_mb_growth_resorption(8192, res=1024, max_steps=10_000), lib, cap_left=2048, cap_right=2560)
ens.arrays["idx_hash"] = None
init = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in ens.arrays.items()}
pinned = {k: torch.from_numpy(v).pin_memory() for k, v in ens.arrays.items() if isinstance(v, np.ndarray)}
for k, t in pinned.items():
    ens.arrays[k] = t.numpy()
times = []
for i in range(N + 2):
    if i % 4 == 0:
        for k, v in init.items():
            if isinstance(v, np.ndarray):
                ens.arrays[k][...] = v
    t0 = time.perf_counter()
    assert ens.run(lib, max_steps=250) == 0
    times.append(1e3 * (time.perf_counter() - t0))
t = np.array(times[2:])
print(f"{N} calls: median {np.median(t):.1f} ms, p90 {np.percentile(t, 90):.1f}, max {t.max():.1f}, calls above 1.2 x median: {int((t > 1.2 * np.median(t)).sum())}: "
      f"{[round(v, 1) for v in t[t > 1.2 * np.median(t)]]}")
