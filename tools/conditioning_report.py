"""Conditioning report of the per-step linear system: the dense system exactly as the reference
hands it to `\\` (functions.jl:1290-1296) is captured from the oracle at chosen steps of a run and solved three ways -
the oracle's restated partial-pivot LU (what `dense \\` does), the structured O(N) solve every kernel of this repo
follows (two one-sided eliminations + the 2x2 interface condition), LAPACK (numpy) - and compared with an mpmath LU
at 60 digits.  CPU only, test infrastructure: uses oracle/.

usage: python tools/conditioning_report.py [C1|MB] [steps ...]      (default: C1 at steps 1 100 5000 18000)
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import mbm_b200  # noqa: E402,F401
from mbm_b200 import abi, drivers as d  # noqa: E402
from oracle import load as oracle_load  # noqa: E402


def capture_step(lib, ens, faithful):
    """advance ens by one step with the capture hook armed; returns (A, rhs, sol) of that step"""
    cap = int(ens.cap_left + ens.cap_right)
    A = np.zeros(cap * cap); rhs = np.zeros(cap); sol = np.zeros(cap); N = C.c_int64(0)
    dp = C.POINTER(C.c_double)
    lib.mbm_oracle_capture_next_system(A.ctypes.data_as(dp), rhs.ctypes.data_as(dp), sol.ctypes.data_as(dp), cap, C.byref(N))
    assert ens.run(lib, max_steps=1, flags=abi.FLAG_SOLVER_FAITHFUL if faithful else 0) == 0 and ens.status[0] == 0
    lib.mbm_oracle_capture_next_system(None, None, None, 0, None)
    n = N.value
    assert n > 0, "nothing captured"
    return A[: n * n].reshape(n, n).copy(), rhs[:n].copy(), sol[:n].copy()


def mp_solve(A, b, dps=60):
    import mpmath as mp
    mp.mp.dps = dps
    x = mp.lu_solve(mp.matrix(A.tolist()), mp.matrix(b.tolist()))
    return x, mp


def rel_err(x, x_mp, mp):
    num = mp.sqrt(sum((mp.mpf(float(a)) - b) ** 2 for a, b in zip(x, x_mp)))
    den = mp.sqrt(sum(b ** 2 for b in x_mp))
    return float(num / den)


def report(kind="C1", steps=(1, 100, 5000, 18000), res=None, out=print):
    lib = oracle_load.load()
    pars = d.pars_C1() if kind == "C1" else d.pars_DCGMB()
    if res is not None:
        pars.update(res=list(res))
    ens, _ = d.build_couple_ensemble([pars], lib)
    rows = []
    out(f"{kind}: res {pars['res']}, ndim {pars.get('ndim', pars.get('n'))}")
    out("| step | N | cond2(L) | cond2 after row equilibration | dense partial-pivot LU (oracle) | LAPACK (numpy) | structured |")
    out("|---|---|---|---|---|---|---|")
    for k in steps:
        if k - 1 > int(ens.it[0]):
            assert ens.run(lib, max_steps=k - 1 - int(ens.it[0])) == 0 and ens.status[0] == 0, ens.status
        if int(ens.it[0]) != k - 1:
            out(f"| {k} | run ended at step {int(ens.it[0])} | | | | | |")
            break
        e1, e2 = ens.copy(), ens.copy()
        A, b, x_dense = capture_step(lib, e1, True)
        A2, b2, x_struct = capture_step(lib, e2, False)
        assert np.array_equal(A, A2) and np.array_equal(b, b2)
        x_mp, mp = mp_solve(A, b)
        s = np.linalg.svd(A, compute_uv=False)
        Aeq = A / np.abs(A).max(axis=1, keepdims=True)
        se = np.linalg.svd(Aeq, compute_uv=False)
        row = dict(step=k, N=A.shape[0], cond=s[0] / s[-1], cond_eq=se[0] / se[-1], dense=rel_err(x_dense, x_mp, mp),
                   lapack=rel_err(np.linalg.solve(A, b), x_mp, mp), structured=rel_err(x_struct, x_mp, mp))
        rows.append(row)
        out("| {step} | {N} | {cond:.1e} | {cond_eq:.1e} | {dense:.1e} | {lapack:.1e} | {structured:.1e} |".format(**row))
    return rows


if __name__ == "__main__":
    kind = sys.argv[1] if len(sys.argv) > 1 else "C1"
    steps = tuple(int(a) for a in sys.argv[2:]) or ((1, 100, 5000, 18000) if kind == "C1" else (1, 6, 12))
    report(kind, steps)
