import sys, time, os
sys.path.insert(0, '/root/repo')
os.environ.setdefault("MBM_PIECES", "1")
import numpy as np
import mbm_b200
from mbm_b200 import drivers as d, ensembles as E
lib = mbm_b200.load_cuda_library()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
res = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
pars = E.c5_highres(M, res, max_steps=100)
if os.environ.get('V0'):                      # config 5's fixed-mesh variant: V_ip = 0
    for p in pars: p.update(V_ip=0.0)
ens, _ = d.build_couple_ensemble(pars, lib, cap_left=res + 4, cap_right=res + 4)
if os.environ.get('NOHASH'): ens.arrays['idx_hash'] = None
ens.run(lib, max_steps=2)
n0 = int(ens.node_steps.sum())
for rep in range(2):
    t0 = time.perf_counter(); assert ens.run(lib, max_steps=steps) == 0; te = time.perf_counter() - t0
    ks = lib.mbm_last_kernel_seconds(); ns = int(ens.node_steps.sum()) - n0; n0 += ns
    print(f"c5 M={M} steps={steps} res={res}: kernel {ks*1e3:.2f} ms -> {ns/ks:.4g} node-steps/s = {ns/ks*32e-9:.1f} GB/s algorithmic; "
          f"e2e {te*1e3:.1f} ms -> {ns/te:.4g}; failed {int((ens.status!=0).sum())} launches {lib.mbm_last_kernel_launches()}")
