import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import mbm_b200
from mbm_b200 import drivers as d, abi
lib = mbm_b200.load_cuda_library()
# tuned kernel (MB + flux, growth/resorption), general kernel (fixed interface, Stefan, simple), hand-over
pars = []
for V, res in ((-1e-7, [60, 90]), (1e-7, [100, 150]), (3e-8, [33, 70])):
    p = d.pars_MB_growth(); p.update(V_ip=V, res=res); pars.append(p)
e, _ = d.build_couple_ensemble(pars, lib); assert e.run(lib, max_steps=12) == 0 and np.all(e.status == 0)
p = d.pars_B5(); e, _ = d.build_couple_ensemble([p], lib); assert e.run(lib, max_steps=12) == 0 and np.all(e.status == 0)
p = d.pars_B1(); e, _ = d.build_couple_ensemble([p], lib); assert e.run(lib, max_steps=12) == 0 and np.all(e.status == 0)
e = d.build_stefan_ensemble([d.pars_CSP()], lib); assert e.run(lib, max_steps=12) == 0 and np.all(e.status == 0)
e = d.build_simple_ensemble([d.pars_A2()]); assert e.run(lib, max_steps=12) == 0 and np.all(e.status == 0)
import os
os.environ["MBM_KCAP"] = "170"
pars = []
for V, res in ((-1e-7, [60, 90]), (1e-7, [60, 90])):
    p = d.pars_MB_growth(); p.update(V_ip=V, res=res); pars.append(p)
e, _ = d.build_couple_ensemble(pars, lib); assert e.run(lib, max_steps=8) == 0 and np.all(e.status == 0)
# streaming kernel (bulk-copy row staging): uniform meshes, growth and resorption, columns of several rows
os.environ.pop("MBM_KCAP", None)
pars = []
for V, res in ((1e-7, [300, 500]), (-1e-7, [300, 500]), (2e-7, [257, 258]), (-3e-7, [1500, 1600])):
    p = d.pars_MB_growth()
    p.update(V_ip=V, res=res, MRefin=1.0, Ri=[0.45, 1.0], t_tot=2e4, KD_ar=(0.1, 0.1), Cl_i=0.5, Cr_i=5.0)
    pars.append(p)
e, _ = d.build_couple_ensemble(pars, lib, cap_left=3000, cap_right=3600)
assert e.run(lib, max_steps=10, flags=abi.FLAG_FORCE_STREAMING) == 0 and np.all(e.status == 0), e.status
print("sanitizer case ok")
