"""the per-step system as the reference hands it to `\\` (functions.jl:1290-1296), captured
from the oracle, solved by the restated dense LU and by the structured solve the kernels follow, against an mpmath LU.
CPU only."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
import conditioning_report as cr  # noqa: E402

mpmath = pytest.importorskip("mpmath")


def test_structured_solve_is_the_better_conditioned_one_on_the_rayleigh_fixture():
    # examples/C1.jl at a third of its resolution (the mpmath LU is O(N^3) in Python): sphere, MB row with raw dVolC entries
    rows = cr.report("C1", steps=(1, 40), res=(30, 40), out=lambda *_: None)
    assert [r["step"] for r in rows] == [1, 40] and rows[0]["N"] == 71
    for r in rows:
        assert r["cond"] > 1e10 and r["cond_eq"] < 1e-4 * r["cond"]          # the bad scaling is the mass-balance row
        assert r["structured"] < 1e-11                                        # what the kernels reproduce
        assert r["structured"] < r["dense"]                                   # partial-pivot LU on the unscaled rows loses digits
        assert r["dense"] < 1e-3 and r["lapack"] < 1e-3                       # ... but stays usable (the reference's KATs pass)


def test_captured_system_is_the_one_both_solvers_see():
    lib = cr.oracle_load.load()
    from mbm_b200 import drivers as d
    ens, _ = d.build_couple_ensemble([d.pars_DCGMB()], lib)
    e1, e2 = ens.copy(), ens.copy()
    A1, b1, x1 = cr.capture_step(lib, e1, True)
    A2, b2, x2 = cr.capture_step(lib, e2, False)
    assert np.array_equal(A1, A2) and np.array_equal(b1, b2)
    n1 = int(ens.res[0, 0])                    # V_ip < 0: the node is inserted on the right, the left block keeps its size
    assert A1.shape[0] == n1 + int(ens.res[0, 1]) + 1
    # the two interface rows: dense mass-balance row and the partition row (set_inner_bc_mb!, functions.jl:1094-1109)
    assert np.count_nonzero(A1[n1 - 1]) == A1.shape[0] and np.count_nonzero(A1[n1]) == 2
    assert np.allclose(A1 @ x1, b1, rtol=0, atol=1e-9 * np.abs(b1).max()) and np.allclose(x1, x2, rtol=1e-9)
    # capturing does not disturb the step
    e3 = ens.copy(); assert e3.run(lib, max_steps=1) == 0
    assert all(np.array_equal(u, v) for u, v in zip(e3.member(0), e2.member(0)))
