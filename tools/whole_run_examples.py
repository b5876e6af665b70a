"""Whole runs of the long example fixtures (examples/B7.jl, C2.jl: 410 719 steps at CFL 0.005; Diff_couple_Flux_growth.jl:
39 531 steps) through the CUDA library and through the oracle: integers, regrid checksum, worst relative error, times."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
import mbm_b200
from mbm_b200 import abi, drivers as d
from oracle import load as oracle_load

cuda, ora = mbm_b200.load_cuda_library(), oracle_load.load()
for name in [a for a in sys.argv[1:] if a not in ("CSP", "A2")] or (() if sys.argv[1:] else ("Flux_growth", "B7", "C2")):
    pars = getattr(d, "pars_" + name)()
    eo, _ = d.build_couple_ensemble([pars], ora)
    eg, _ = d.build_couple_ensemble([pars], cuda)
    t0 = time.perf_counter(); assert eo.run(ora, flags=abi.FLAG_MASS_FINAL) == 0; to = time.perf_counter() - t0
    t0 = time.perf_counter(); assert eg.run(cuda, flags=abi.FLAG_MASS_FINAL) == 0; tg = time.perf_counter() - t0
    ints = all(np.array_equal(eo.arrays[k], eg.arrays[k]) for k in ("res", "it", "node_steps", "idx_hash", "status"))
    worst = max(float(np.max(np.abs(a - b)) / np.max(np.abs(a))) for a, b in zip(eo.member(0), eg.member(0)))
    print(f"{name}: steps {int(eg.it[0])}, res {eg.res[0].tolist()}, status {int(eg.status[0])}, integers+checksum equal {ints}, "
          f"worst rel err {worst:.2e}, mass err oracle {float((eo.Mass[0]-eo.Mass0[0])/eo.Mass0[0]):.2e} cuda "
          f"{float((eg.Mass[0]-eg.Mass0[0])/eg.Mass0[0]):.2e}, oracle {to:.1f} s, cuda (one member) {tg:.1f} s", flush=True)

if not sys.argv[1:] or "CSP" in sys.argv[1:]:
    # main_codes/Chemical_Stefan_problem_XT.jl (rho == 1): 9 293 steps, the velocity comes from the solution every step
    eo, eg = d.build_stefan_ensemble([d.pars_CSP()], ora), d.build_stefan_ensemble([d.pars_CSP()], cuda)
    assert eo.run(ora) == 0 and eg.run(cuda) == 0
    ints = all(np.array_equal(eo.arrays[k], eg.arrays[k]) for k in ("res", "it", "node_steps", "status"))
    worst = max(float(np.max(np.abs(a - b)) / np.max(np.abs(a))) for a, b in zip(eo.member(0), eg.member(0)))
    print(f"CSP: steps {int(eg.it[0])}, res {eg.res[0].tolist()}, status {int(eg.status[0])}, integers equal {ints}, worst rel err "
          f"{worst:.2e}, Ri {float(eo.Ri[0, 0])!r} vs {float(eg.Ri[0, 0])!r}", flush=True)
if not sys.argv[1:] or "A2" in sys.argv[1:]:
    # examples/A2.jl: 41 647 steps, sphere, Neumann / Dirichlet
    eo = d.build_simple_ensemble([d.pars_A2()]); eg = eo.copy()
    assert eo.run(ora) == 0 and eg.run(cuda) == 0
    n = int(eo.res[0])
    print(f"A2: steps {int(eg.it[0])} / {int(eo.it[0])}, status {int(eg.status[0])}, worst rel err "
          f"{float(np.max(np.abs(eo.C[0, :n] - eg.C[0, :n])) / np.max(np.abs(eo.C[0, :n]))):.2e}", flush=True)
