"""examples/D1.jl over its whole run (about 6.6 million steps; in-loop olivine / melt diffusivities): `oracle` writes the
oracle's final state to tools/_d1_oracle_final.npz (about 150 s on one core), `cuda` runs the CUDA library and compares."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
import mbm_b200
from mbm_b200 import drivers as d

F = '/root/repo/tools/_d1_oracle_final.npz'
mode = sys.argv[1] if len(sys.argv) > 1 else "cuda"
if mode == "oracle":
    from oracle import load as oracle_load
    lib = oracle_load.load()
else:
    lib = mbm_b200.load_cuda_library()
ens = d.build_stefan_ensemble([d.pars_D1()], lib)
t0 = time.perf_counter(); rc = ens.run(lib); el = time.perf_counter() - t0
xl, cl, xr, cr = ens.member(0)
print(f"{mode}: rc {rc}, status {int(ens.status[0])}, steps {int(ens.it[0])}, res {ens.res[0].tolist()}, t {float(ens.t[0])!r}, "
      f"Ri {float(ens.Ri[0, 0])!r}, {el:.1f} s", flush=True)
if mode == "oracle":
    np.savez(F, xl=xl, cl=cl, xr=xr, cr=cr, it=ens.it, res=ens.res, Ri=ens.Ri, t=ens.t, status=ens.status, node_steps=ens.node_steps)
else:
    o = np.load(F)
    ints = all(np.array_equal(o[k], getattr(ens, k)) for k in ("it", "res", "status", "node_steps"))
    worst = max(float(np.max(np.abs(a - b)) / np.max(np.abs(a))) for a, b in zip((o["xl"], o["cl"], o["xr"], o["cr"]), (xl, cl, xr, cr))
                ) if ints else float("nan")
    print(f"integers equal {ints}; oracle steps {int(o['it'][0])}, res {o['res'][0].tolist()}; worst rel err {worst:.2e}; "
          f"Ri rel diff {abs(float(o['Ri'][0, 0]) - float(ens.Ri[0, 0])) / float(o['Ri'][0, 0]):.2e}")
