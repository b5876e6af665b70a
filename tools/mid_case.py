import sys, time, os
sys.path.insert(0, '/root/repo')
os.environ.setdefault("MBM_PIECES", "1")
import numpy as np
import mbm_b200
from mbm_b200 import drivers as d
lib = mbm_b200.load_cuda_library()
pars=[]
rng=np.random.default_rng(1)
for k in range(296):
    p = d.pars_MB_growth(); p.update(V_ip=float(rng.uniform(0.5e-7,2e-7)), res=[1500, 2000]); pars.append(p)
ens,_ = d.build_couple_ensemble(pars, lib)
ens.arrays['idx_hash']=None
ens.run(lib, max_steps=2)
n0=int(ens.node_steps.sum())
for rep in range(2):
    assert ens.run(lib, max_steps=100)==0
    ks=lib.mbm_last_kernel_seconds(); ns=int(ens.node_steps.sum())-n0; n0+=ns
    print(f"mid-size N=3500: kernel {ks*1e3:.2f} ms -> {ns/ks:.4g} node-steps/s; failed {int((ens.status!=0).sum())} launches {lib.mbm_last_kernel_launches()}")
