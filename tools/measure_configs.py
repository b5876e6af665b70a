#!/usr/bin/env python
"""Measure the other BASELINE.json configurations (c2a, c2b, c3, c5) once on a B200 and check a subset of
every ensemble against the CPU oracle.  These are parity-test cases, not bench lines (bench.py measures c4);
the numbers go to profiles/ as a record of where each kernel variant stands."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
os.environ.setdefault("MBM_PIECES", "1")     # one stream: kernel_s is then the plain sum of the kernel launches
sys.path.insert(0, str(ROOT))
import mbm_b200  # noqa: E402
from mbm_b200 import abi, drivers as d, ensembles as E  # noqa: E402
from oracle import load as oracle_load  # noqa: E402


def relerr(a, b):
    den = np.max(np.abs(a))
    return 0.0 if den == 0 else float(np.max(np.abs(a - b)) / den)


def run(name, gen, build, M, max_steps, n_check, flags=0, tol=1e-10, ints=True, gate=None):
    cuda, ora = mbm_b200.load_cuda_library(), oracle_load.load()
    pars = gen(M)
    t0 = time.perf_counter()
    eg = build(cuda, pars)
    tb = time.perf_counter() - t0
    eo = build(ora, pars[:n_check])      # the same first members through the oracle
    eg0 = eg.copy() if gate else None    # the initial state (the known-answer gates need the initial profiles)
    warm = eg.copy()                     # a throw-away copy takes the one-off costs of the first call (allocator pool, module load)
    assert warm.run(cuda, max_steps=2, flags=flags) == 0
    del warm
    t0 = time.perf_counter()
    rc = eg.run(cuda, max_steps=max_steps, flags=flags)
    te = time.perf_counter() - t0
    ks = cuda.mbm_last_kernel_seconds()
    assert rc == 0, rc
    t0 = time.perf_counter()
    assert eo.run(ora, max_steps=max_steps) == 0
    tc = time.perf_counter() - t0
    worst = 0.0
    for m in range(n_check):
        for a, b in zip(eo.member(m), eg.member(m)):
            worst = max(worst, relerr(a, b))
    ok_int = bool(np.array_equal(eo.res, eg.res[:n_check]) and np.array_equal(eo.it, eg.it[:n_check]) and
                  np.array_equal(eo.status, eg.status[:n_check]))
    if ints:
        ok_int = ok_int and bool(np.array_equal(eo.idx_hash, eg.idx_hash[:n_check]))
    ns = int(eg.node_steps.sum())
    out = {"config": name, "members": M, "max_steps": max_steps, "node_steps": ns, "kernel_s": ks, "e2e_s": te,
           "node_steps_per_s_kernel": ns / ks if ks > 0 else None, "node_steps_per_s_e2e": ns / te,
           "members_failed": int((eg.status != 0).sum()), "parity_members": n_check, "parity_max_relerr": worst,
           "parity_integers_exact": ok_int, "oracle_node_steps_per_s": int(eo.node_steps.sum()) / tc,
           "build_s": tb}
    if gate:
        out.update(gate(cuda, eg0, eg))
    print(json.dumps(out), flush=True)
    assert worst < tol, (name, worst)
    assert ok_int, name
    return out


def main():
    which = sys.argv[1:] or ["c2a", "c2b", "c3", "c5"]
    res = []
    if "c2a" in which:   # examples/C1.jl sweep, 4096 members, whole run (~18k steps each) is long: 2000 steps
        res.append(run("c2a: C1 MB/Rayleigh sweep, res [80 120]", E.c2a_rayleigh_sweep, lambda lib, P: d.build_couple_ensemble(P, lib)[0],
                       4096, 2000, 4))
    if "c2a" in which or "c2afull" in which:
        # the whole sweep run to t_tot (~18 000 steps per member), every member checked on the device against its own
        # Rayleigh curve (test_C1.jl's rtol is 1e-4 at the shipped parameters, the sweep widens KD0 / Ri / CFL)
        cuda = mbm_b200.load_cuda_library()
        pars = E.c2a_rayleigh_sweep(4096)
        ens, aux = d.build_couple_ensemble(pars, cuda)
        t0 = time.perf_counter()
        assert ens.run(cuda, max_steps=0) == 0
        te = time.perf_counter() - t0
        ks = cuda.mbm_last_kernel_seconds()
        out, err = ens.validate(cuda, abi.VAL_RAYLEIGH, Ri0=np.array(aux["Ri0"]), C0_r=np.array([c[-1] for c in aux["C0_r"]]), KD0=aux["KD0"])
        ok = ens.status == 0
        r = {"config": "c2a full runs: C1 MB/Rayleigh sweep to t_tot, every member against its own Rayleigh curve (device-side gate)",
             "members": 4096, "members_failed": int((~ok).sum()), "steps_min_max": [int(ens.it[ok].min()), int(ens.it[ok].max())],
             "node_steps": int(ens.node_steps.sum()), "kernel_s": ks, "e2e_s": te, "node_steps_per_s_kernel": int(ens.node_steps.sum()) / ks,
             "rayleigh_gate_max_relerr": float(np.max(err[ok])), "rayleigh_gate_median_relerr": float(np.median(err[ok])),
             "rayleigh_gate_members_within_1e-4": int(np.sum(err[ok] < 1e-4)), "rayleigh_gate_members_within_5e-3": int(np.sum(err[ok] < 5e-3))}
        print(json.dumps(r), flush=True)
        res.append(r)
    if "c2b" in which:
        res.append(run("c2b: Stefan sweep, res [15 25]", E.c2b_stefan_sweep, lambda lib, P: d.build_stefan_ensemble(P, lib),
                       4096, 2000, 4, tol=1e-8, ints=False))
    if "c3" in which:
        def sphere_gate(lib, e0, e1):
            # test/test_B1.jl's gate (sphere series, rtol 1e-2) for every member of the sweep, evaluated on the device
            t0 = time.perf_counter()
            out, err = e1.validate(lib, abi.VAL_SPHERE, zeta=e1.arrays["Di"][:, 0] * e1.arrays["t"], C0_left=e0.arrays["C_left"],
                                   C0_right=e0.arrays["C_right"])
            return {"sphere_gate_max_relerr": float(np.max(err)), "sphere_gate_members_within_1e-2": int(np.sum(err < 1e-2)),
                    "sphere_gate_s": time.perf_counter() - t0}
        res.append(run("c3: flux fixed interface, n=256/phase", lambda M: E.c3_flux_fixed(M, 256), lambda lib, P: d.build_couple_ensemble(P, lib)[0],
                       16384, 1000, 4, gate=sphere_gate))
    if "c5" in which:
        res.append(run("c5: MB growth n=65536/phase MRefin=1 (global-memory variant)",
                       lambda M: E.c5_highres(M, 65536, max_steps=100),
                       lambda lib, P: d.build_couple_ensemble(P, lib, cap_left=65540, cap_right=65540)[0],
                       296, 20, 1, tol=1e-9))
    if "c5" in which or "c5v0" in which:
        def c5v0(M):
            pars = E.c5_highres(M, 65536, max_steps=100)
            for p in pars:
                p.update(V_ip=0.0)
            return pars
        res.append(run("c5 fixed-mesh variant: V_ip = 0, n=65536/phase, LinRange meshes (closed-form streaming kernel, no regrid)",
                       c5v0, lambda lib, P: d.build_couple_ensemble(P, lib, cap_left=65540, cap_right=65540)[0],
                       296, 20, 1, tol=1e-9, ints=False))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "configs.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
