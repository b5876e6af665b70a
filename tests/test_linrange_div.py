"""The closed-form streaming kernel (csrc/mbm_kernel_stream2.cuh) recomputes LinRange nodes with t = i/d evaluated as
i * RN(1/d) corrected by one FMA.  That must be the correctly rounded quotient the reference's Base.lerpi computes
(Julia LinRange getindex): tools/linrange_div_check.c compares it with the IEEE division for every integer pair
0 <= i <= d in a range of d.  The whole range the kernel accepts (d <= 300 000) takes ~40 s on 8 cores and was run once
(0 mismatches); the CPU suite checks the bands around the sizes the configurations use."""
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def test_corrected_reciprocal_division_is_exact(tmp_path):
    exe = tmp_path / "linrange_div_check"
    subprocess.run(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-o", str(exe), str(ROOT / "tools" / "linrange_div_check.c"),
                    "-lm"], check=True)
    for lo, hi in ((2, 6000), (65000, 66100), (131000, 131100), (299900, 300000)):
        out = subprocess.run([str(exe), str(lo), str(hi)], capture_output=True, text=True)
        assert out.returncode == 0 and "mismatches 0" in out.stdout, out.stdout


def test_numpy_restatement_of_the_identity():
    """the same identity in numpy (float64 products are exact to 106 bits through np.longdouble-free arithmetic: the
    remainder i - q0*d is formed with math.fma-free integer arithmetic here, as an independent check of the C program)"""
    from fractions import Fraction
    rng = np.random.default_rng(7)
    for d in (65535, 65536, 131071, 99991, 7):
        r = 1.0 / d
        for i in list(rng.integers(0, d + 1, 200)) + [0, 1, d - 1, d]:
            q0 = float(i) * r
            rem = float(Fraction(int(i)) - Fraction(q0) * d)          # exact remainder, then rounded (it is representable)
            assert Fraction(rem) == Fraction(int(i)) - Fraction(q0) * d
            q = float(Fraction(rem) * Fraction(r) + Fraction(q0))    # one rounding: what the FMA does
            assert q == float(i) / float(d), (i, d)
