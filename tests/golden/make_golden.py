#!/usr/bin/env python
"""Generates tests/golden/couple_fixtures.npz: short runs of the reference's shipped parameter sets through the CPU
oracle (the reference itself is Julia and cannot run in this image, so these vectors pin the RESTATEMENT, which in
turn is pinned to the reference by its known-answer tests, tests/test_oracle_kat.py).  Integers (node counts, step
counter, regrid-index checksum, node-steps) are stored exactly, profiles as float64.  Re-run only when the oracle is
changed on purpose:  python tests/golden/make_golden.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from mbm_b200 import abi, drivers as d  # noqa: E402
from oracle import load as oracle_load  # noqa: E402

# name -> (parameter block of the reference driver, time steps)
FIXTURES = {
    "B1_flux_fixed": (d.pars_B1, 300),            # examples/B1.jl
    "B2_flux_arrhenius": (d.pars_B2, 300),        # examples/B2.jl
    "B4_flux_growth_sphere": (d.pars_B4, 400),    # examples/B4.jl
    "B5_flux_growth_planar": (d.pars_B5, 400),    # examples/B5.jl
    "C1_mb_growth_sphere": (d.pars_C1, 400),      # examples/C1.jl
    "MB_growth_resorption": (d.pars_MB_growth, 400),   # examples/Diff_couple_MB_growth.jl
    "DCGMB_main": (d.pars_DCGMB, 12),             # main_codes/Diff_couple_1D_general_MB.jl (whole run)
    "DCGF_main": (d.pars_DCGF, 6),                # main_codes/Diff_couple_1D_general_Flux.jl (whole run)
    "B6_flux_growth_kdt": (d.pars_B6, 400),       # examples/B6.jl
    "B7_flux_resorption_kdt": (d.pars_B7, 400),   # examples/B7.jl
    "C2_mb_resorption_kdt": (d.pars_C2, 400),     # examples/C2.jl
    "Flux_growth_kdt": (d.pars_Flux_growth, 400), # examples/Diff_couple_Flux_growth.jl
    "DC_Flux_fixed_arrhenius": (d.pars_DC_Flux, 268),   # examples/Diff_couple_Flux.jl (whole run)
    "DC_MB_fixed_arrhenius": (d.pars_DC_MB, 268),       # examples/Diff_couple_MB.jl (whole run)
    "DCNI_hold": (d.pars_DCNI, 209),              # examples/Diff_couple_no_interaction.jl (whole run)
}


def run_fixture(lib, name):
    pars, steps = FIXTURES[name]
    ens, _ = d.build_couple_ensemble([pars()], lib)
    assert ens.run(lib, max_steps=steps, flags=abi.FLAG_MASS_FINAL) == 0
    xl, cl, xr, cr = ens.member(0)
    return dict(status=ens.status.copy(), res=ens.res[0].copy(), it=ens.it.copy(), idx_hash=ens.idx_hash.copy(),
                node_steps=ens.node_steps.copy(), Ri=ens.Ri[0].copy(), t=ens.t.copy(), x_left=xl.copy(), C_left=cl.copy(),
                x_right=xr.copy(), C_right=cr.copy())


def main():
    ora = oracle_load.load()
    out = {}
    for name in FIXTURES:
        for k, v in run_fixture(ora, name).items():
            out[f"{name}/{k}"] = v
    np.savez_compressed(Path(__file__).with_name("couple_fixtures.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
