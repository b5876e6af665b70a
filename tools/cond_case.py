import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np
import mbm_b200
from mbm_b200 import abi, drivers as d
from oracle import load as oracle_load
lib, ora = mbm_b200.load_cuda_library(), oracle_load.load()
def rel(a,b):
    den=np.max(np.abs(a)); return float(np.max(np.abs(a-b))/den)
for res in ([1200,1300],[2400,2600],[5000,5500]):
  for n in (1,2,3):
    p = d.pars_MB_growth(); p.update(V_ip=2.97e-7, res=res, MRefin=1.0, Ri=[0.45, 1.0], t_tot=2e4, KD_ar=(0.5, 0.4), Cl_i=0.5, Cr_i=1.0, n=n)
    eo,_ = d.build_couple_ensemble([p], ora, cap_left=12000, cap_right=12000)
    er = eo.copy(); es = eo.copy()
    eo.run(ora, max_steps=55)
    es.run(lib, max_steps=55, flags=abi.FLAG_FORCE_STREAMING)
    msg = f"res={res} n={n}: streaming vs oracle {max(rel(u,v) for u,v in zip(eo.member(0), es.member(0))):.2e}"
    if sum(res) < 5000:
        er.run(lib, max_steps=55)
        msg += f", resident vs oracle {max(rel(u,v) for u,v in zip(eo.member(0), er.member(0))):.2e}, streaming vs resident {max(rel(u,v) for u,v in zip(er.member(0), es.member(0))):.2e}"
    print(msg)
