import sys, time, os
sys.path.insert(0, '/root/repo')
os.environ.setdefault("MBM_PIECES", "1")
import numpy as np
import mbm_b200
from mbm_b200 import drivers as d, ensembles as E
lib = mbm_b200.load_cuda_library()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
ens, _ = d.build_couple_ensemble(E.c2a_rayleigh_sweep(M), lib)
if os.environ.get('NOHASH'): ens.arrays['idx_hash'] = None
ens.run(lib, max_steps=10)
n0 = int(ens.node_steps.sum())
for rep in range(2):
    t0 = time.perf_counter(); assert ens.run(lib, max_steps=steps) == 0; te = time.perf_counter() - t0
    ks = lib.mbm_last_kernel_seconds(); ns = int(ens.node_steps.sum()) - n0; n0 += ns
    print(f"c2a M={M} steps={steps}: kernel {ks*1e3:.2f} ms -> {ns/ks:.4g} node-steps/s; e2e {te*1e3:.1f} ms -> {ns/te:.4g}; failed {int((ens.status!=0).sum())} res {ens.res[0]}")
