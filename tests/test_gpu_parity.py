"""Parity tests proper: the CUDA path (through the C ABI of libmbm_b200.so) against the CPU oracle on
identical inputs.  Bar: node counts, step counts, status codes and the regrid-index checksum bit-exact;
profiles, meshes and interface position within 1e-10 relative (FP64), as BASELINE.json states."""
import os

import numpy as np
import pytest

from mbm_b200 import abi, drivers as d

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _relerr(a, b):
    den = np.max(np.abs(a))
    return 0.0 if den == 0 else float(np.max(np.abs(a - b)) / den)


def compare(eo, eg, tol=TOL, ints=True, simple=False):
    assert np.array_equal(eo.status, eg.status), (eo.status, eg.status)
    if ints:
        for k in ("res", "it", "node_steps") + (() if simple else ("idx_hash",)):
            assert np.array_equal(eo.arrays[k], eg.arrays[k]), f"{k}: {eo.arrays[k]} vs {eg.arrays[k]}"
    worst = 0.0
    for m in range(eo.M):
        if simple:
            n = int(eo.res[m])
            pairs = [(eo.C[m, :n], eg.C[m, :n])]
        else:
            pairs = list(zip(eo.member(m), eg.member(m)))
        for a, b in pairs:
            worst = max(worst, _relerr(a, b))
    assert worst < tol, f"relative error {worst}"
    if not simple:
        assert _relerr(eo.Ri, eg.Ri) < 1e-13
        assert _relerr(eo.t, eg.t) < 1e-13
    return worst


def run_both(pars_list, oracle, cuda_lib, max_steps, flags=0):
    eo, _ = d.build_couple_ensemble(pars_list, oracle)
    eg, _ = d.build_couple_ensemble(pars_list, cuda_lib)
    assert np.array_equal(eo.x_left, eg.x_left) and np.array_equal(eo.x_right, eg.x_right), "create_grid differs"
    assert _relerr(eo.Mass0, eg.Mass0) < 1e-15
    assert eo.run(oracle, max_steps=max_steps, flags=flags | abi.FLAG_MASS_FINAL) == 0
    assert eg.run(cuda_lib, max_steps=max_steps, flags=flags | abi.FLAG_MASS_FINAL) == 0
    return eo, eg


def test_library_identity(cuda_lib):
    assert cuda_lib.mbm_backend() == b"cuda-sm_100a"
    assert cuda_lib.mbm_device_count() >= 1


@pytest.mark.parametrize("name,pars,steps", [
    ("B1_flux_fixed", d.pars_B1, 400),
    ("B2_flux_arrhenius", d.pars_B2, 400),
    ("B4_flux_growth_sphere", d.pars_B4, 600),
    ("B5_flux_growth_planar", d.pars_B5, 600),
    ("C1_mb_growth_sphere", d.pars_C1, 600),
    ("B1_flux_fixed_whole_run", d.pars_B1, 0),
    ("B2_flux_arrhenius_whole_run", d.pars_B2, 0),
    ("B4_flux_growth_sphere_whole_run", d.pars_B4, 0),
    ("MBgrowth_mb_resorption_cyl", d.pars_MB_growth, 400),
    ("MBmain_kd_t", d.pars_DCGMB, 0),
    ("Fluxmain", d.pars_DCGF, 0),
    ("DCNI_hold_dirichlet", d.pars_DCNI, 0),
    ("B6_flux_growth_kd_t", d.pars_B6, 0),
    ("B7_flux_resorption_kd_t", d.pars_B7, 800),
    ("C2_mb_resorption_kd_t", d.pars_C2, 800),
    ("Fluxgrowth_cyl_kd_t", d.pars_Flux_growth, 800),
    ("DCFlux_fixed_arrhenius_kd_t", d.pars_DC_Flux, 0),
    ("DCMB_fixed_arrhenius_kd_t", d.pars_DC_MB, 0),
])
def test_fixture_parity(oracle, cuda_lib, name, pars, steps):
    """every shipped couple fixture of SURVEY.md §B, first `steps` steps (0 = whole run)"""
    tol = 2e-9 if "arrhenius" in name else TOL   # CUDA exp vs glibc exp: 1-ulp D differences feed dt
    eo, eg = run_both([pars()], oracle, cuda_lib, steps)
    compare(eo, eg, tol=tol)
    # the final mass is one more value of the state: 1e-11 after the prefixes, the profiles' own bar after whole runs
    # (B4: 18 050 steps of a flux row on an MRefin = 50 sphere, measured 4.9e-11)
    assert _relerr(eo.Mass, eg.Mass) < (tol if "whole_run" in name else 1e-11)


@pytest.mark.parametrize("name,pars,steps", [
    ("Fluxgrowth", d.pars_Flux_growth, 39531),          # examples/Diff_couple_Flux_growth.jl as shipped
    ("C2", d.pars_C2, 410719),                           # examples/C2.jl as shipped (CFL 0.005): 12 s of oracle time
])
def test_whole_run_of_long_example_fixtures(oracle, cuda_lib, name, pars, steps):
    """whole runs, not prefixes: every step's node counts, regrid indices and pchip branches (checksum) equal, profiles to
    1e-10 after up to 4e5 steps (measured 2e-12 / 9e-13; examples/B7.jl, the flux twin of C2: tools/whole_run_examples.py)"""
    eo, eg = run_both([pars()], oracle, cuda_lib, 0)
    assert int(eg.it[0]) == steps and eg.status[0] == 0
    compare(eo, eg)
    assert _relerr(eo.Mass, eg.Mass) < 1e-11


def test_full_run_B5_known_answer(cuda_lib):
    """reference test/test_B5.jl through the CUDA path: Smith (1955), rtol 1e-4"""
    import benchmarks_ref as B
    xl, xr, x0, Ri, Ri0, Cl, Cr, C0, C0r, KD0, n, maxC, Di, V, tt = d.B5(cuda_lib)
    xan, Can = B.smith(xr, Cr, Ri, Di, tt, KD0, V, n)
    ok, err = B.isapprox(Can, Cr, 1e-4)
    assert ok, err


def test_full_run_C1_known_answer(cuda_lib):
    """reference test/test_C1.jl through the CUDA path: Rayleigh fractionation, rtol 1e-4"""
    import benchmarks_ref as B
    xl, xr, x0, Ri, Ri0, Cl, Cr, C0, C0r, n, maxC, KD0 = d.C1(cuda_lib)
    Fs, Fl, Cliq, Cs, Clp, phi = B.rayleigh_fractionation(xl, Cl, Ri0, Ri, C0r, KD0, n)
    iR, iM = B.trapezoidal_integration(Fs, Cs), B.trapezoidal_integration(phi**n, Clp)
    assert abs(iR - iM) <= 1e-4 * max(abs(iR), abs(iM))


def test_simple_A1(oracle, cuda_lib):
    """examples/A1.jl whole run (842 steps) + the analytical gate of test/test_A1.jl"""
    from mbm_b200.host import sinusoid_profile
    import benchmarks_ref as B
    eo = d.build_simple_ensemble([d.pars_A1()])
    eg = eo.copy()
    assert eo.run(oracle) == 0 and eg.run(cuda_lib) == 0
    compare(eo, eg, simple=True)
    p = d.pars_A1()
    Can = sinusoid_profile(0.0, p["nmodes"], p["L"], p["Di"], eg.t[0], p["Amp"], eg.x[0, :500])
    ok, err = B.isapprox(Can, eg.C[0, :500], 1e-4)
    assert ok, err


def test_simple_arrhenius_cooling(oracle, cuda_lib):
    """examples/Simple_Diff.jl: SDG's column with D(T(t)) from the Arrhenius law (Di = -1, functions.jl:1415), whole run"""
    eo, eg = d.build_simple_ensemble([d.pars_Simple_Diff()]), d.build_simple_ensemble([d.pars_Simple_Diff()])
    assert eo.run(oracle, flags=abi.FLAG_MASS_FINAL) == 0 and eg.run(cuda_lib, flags=abi.FLAG_MASS_FINAL) == 0
    compare(eo, eg, tol=2e-9, simple=True)        # CUDA exp vs glibc exp (1 ulp in D feeds dt)
    assert _relerr(eo.t, eg.t) < 1e-13 and _relerr(eo.D, eg.D) < 1e-14 and _relerr(eo.Mass, eg.Mass) < 1e-9


def test_simple_A2_prefix(oracle, cuda_lib):
    """examples/A2.jl (sphere, Neumann/Dirichlet): first 2000 steps, then the whole run (41 647 steps, measured 9e-14)"""
    eo = d.build_simple_ensemble([d.pars_A2()])
    eg = eo.copy()
    assert eo.run(oracle, max_steps=2000) == 0 and eg.run(cuda_lib, max_steps=2000) == 0
    compare(eo, eg, simple=True)
    assert eo.run(oracle) == 0 and eg.run(cuda_lib) == 0 and int(eg.it[0]) == 41647
    compare(eo, eg, simple=True)


def test_stefan_main(oracle, cuda_lib):
    """main_codes/Chemical_Stefan_problem_XT.jl (rho == 1), the whole run (9 293 steps): V comes from the solution, so
    Ri and dt are value-level (1e-10), node counts and step counts must still agree exactly (measured 9e-14)"""
    eo = d.build_stefan_ensemble([d.pars_CSP()], oracle)
    eg = d.build_stefan_ensemble([d.pars_CSP()], cuda_lib)
    assert eo.run(oracle) == 0 and eg.run(cuda_lib) == 0
    assert int(eg.it[0]) == 9293 and eg.res[0].tolist() == [28, 25]
    assert np.array_equal(eo.status, eg.status) and np.array_equal(eo.res, eg.res) and np.array_equal(eo.it, eg.it)
    for a, b in zip(eo.member(0), eg.member(0)):
        assert _relerr(a, b) < 1e-9
    assert _relerr(eo.Ri, eg.Ri) < 1e-10


def test_resume_equals_single_call(cuda_lib):
    """state in/out: 3 calls of 100 steps == 1 call of 300 steps, bit for bit"""
    ea, _ = d.build_couple_ensemble([d.pars_MB_growth()], cuda_lib)
    eb = ea.copy()
    assert ea.run(cuda_lib, max_steps=300) == 0
    for _ in range(3):
        assert eb.run(cuda_lib, max_steps=100) == 0
    for k in ("x_left", "C_left", "x_right", "C_right", "res", "it", "Ri", "t", "idx_hash"):
        assert np.array_equal(ea.arrays[k], eb.arrays[k]), k


def test_fixed_interface_resume_and_batches(oracle, cuda_lib):
    """fixed-interface kernel (V_ip = 0): its scalar batches (16 steps of the time chain ahead) must not show: calls of
    5 + 9 + 23 steps (batches cut short by max_steps, resumed mid-history) == one call of 37 steps, bit for bit, and
    both match the oracle; constant diffusivities (B1: batched) and an Arrhenius law (B2: per-step scalar phase)"""
    for pars, tol in (([d.pars_B1()], 1e-10), ([d.pars_B2()], 2e-9)):       # Arrhenius: CUDA exp vs glibc exp (see compare)
        ea, _ = d.build_couple_ensemble(pars, cuda_lib)
        eb = ea.copy()
        eo, _ = d.build_couple_ensemble(pars, oracle)
        assert ea.run(cuda_lib, max_steps=37) == 0
        for n in (5, 9, 23):
            assert eb.run(cuda_lib, max_steps=n) == 0
        assert eo.run(oracle, max_steps=37) == 0
        for k in ("x_left", "C_left", "x_right", "C_right", "res", "it", "Ri", "t", "dt"):
            assert np.array_equal(ea.arrays[k], eb.arrays[k]), k
        assert np.all(ea.status == 0) and np.all(eo.status == 0)
        assert np.array_equal(ea.it, eo.it) and np.array_equal(ea.res, eo.res)
        for a, b in zip(eo.member(0), ea.member(0)):
            assert np.max(np.abs(a - b)) <= tol * np.max(np.abs(a))


@pytest.mark.parametrize("threads", ["32", "64", "128", "256"])
def test_thread_configs(oracle, cuda_lib, threads, monkeypatch):
    """every CTA width gives the same integers and 1e-10 profiles (different chunking of the solve)"""
    monkeypatch.setenv("MBM_THREADS", threads)
    pars = []
    for V, res in ((-1e-7, [60, 90]), (1e-7, [100, 150]), (-2e-7, [130, 131]), (3e-8, [33, 70])):
        p = d.pars_MB_growth(); p.update(V_ip=V, res=res); pars.append(p)
    eo, eg = run_both(pars, oracle, cuda_lib, 150)
    compare(eo, eg)
    pars = []
    for V, res in ((1.0, [80, 120]), (0.5, [40, 64])):
        p = d.pars_B4(); p.update(V_ip=V, res=res); pars.append(p)
    eo, eg = run_both(pars, oracle, cuda_lib, 150)
    compare(eo, eg)


def test_error_status_codes(oracle, cuda_lib):
    """per-member status mirrors the reference's throw sites; a failed member is frozen, others continue"""
    good = d.pars_MB_growth(); good.update(res=[60, 90])
    fast = d.pars_MB_growth(); fast.update(res=[60, 90], CFL=8.0, V_ip=-1e-4, t_tot=50.0)         # interface moves > 1 cell: functions.jl:50-53
    eo, eg = run_both([good, fast, good], oracle, cuda_lib, 50)
    assert list(eo.status) == list(eg.status)
    assert eo.status[1] in (1, 2) and eo.status[0] == 0 and eo.status[2] == 0
    for m in (0, 2):
        for a, b in zip(eo.member(m), eg.member(m)):
            assert _relerr(a, b) < TOL


def test_capacity_status(oracle, cuda_lib):
    p = d.pars_MB_growth(); p.update(res=[60, 90])
    eo, _ = d.build_couple_ensemble([p], oracle, cap_left=64, cap_right=100)   # nr2 jumps to ~120 at step 1
    eg, _ = d.build_couple_ensemble([p], cuda_lib, cap_left=64, cap_right=100)
    assert eo.run(oracle, max_steps=5) == 0 and eg.run(cuda_lib, max_steps=5) == 0
    assert eo.status[0] == 10 and eg.status[0] == 10


def test_streaming_kernel_matches_resident(cuda_lib):
    """the global-memory variant (MBM_FLAG_FORCE_STREAMING) runs the same arithmetic as the resident one"""
    pars = []
    for V in (-1e-7, 1e-7):
        p = d.pars_MB_growth(); p.update(V_ip=V, res=[100, 150]); pars.append(p)
    ea, _ = d.build_couple_ensemble(pars, cuda_lib)
    eb = ea.copy()
    assert ea.run(cuda_lib, max_steps=60) == 0
    assert eb.run(cuda_lib, max_steps=60, flags=abi.FLAG_FORCE_STREAMING) == 0
    for k in ("res", "it", "idx_hash"):
        assert np.array_equal(ea.arrays[k], eb.arrays[k]), k
    for m in range(2):
        for a, b in zip(ea.member(m), eb.member(m)):
            assert _relerr(a, b) < TOL


def test_handover_to_general_kernel(oracle, cuda_lib, monkeypatch):
    """members that outgrow the tuned kernel's shared-memory capacity are handed over (status -1) to the general
    kernel mid-call and finish there; the result is the same as the oracle's"""
    monkeypatch.setenv("MBM_KCAP", "170")
    pars = []
    for V, res in ((-1e-7, [60, 90]), (1e-7, [60, 90]), (-1e-7, [40, 60]), (2e-7, [30, 40])):
        p = d.pars_MB_growth(); p.update(V_ip=V, res=res); pars.append(p)
    eo, eg = run_both(pars, oracle, cuda_lib, 120)
    compare(eo, eg)


def test_tuned_kernel_thread_configs(oracle, cuda_lib, monkeypatch):
    """the tuned kernel at 32/64/128/256 threads per member"""
    for t in ("32", "64", "128", "256"):
        monkeypatch.setenv("MBM_V2_THREADS", t)
        pars = []
        for V, res in ((-1e-7, [60, 90]), (1e-7, [100, 150]), (3e-8, [33, 70])):
            p = d.pars_MB_growth(); p.update(V_ip=V, res=res); pars.append(p)
        eo, eg = run_both(pars, oracle, cuda_lib, 100)
        compare(eo, eg)


def test_c4_shape_small_sample(oracle, cuda_lib):
    """BASELINE.json configs[3] at full member size (n = 1024 per phase), 6 members x 30 steps"""
    from mbm_b200 import ensembles as E
    pars = E.c4_mb_growth_resorption(6, res=1024, max_steps=10_000)
    eo, _ = d.build_couple_ensemble(pars, oracle, cap_left=2048, cap_right=2560)
    eg, _ = d.build_couple_ensemble(pars, cuda_lib, cap_left=2048, cap_right=2560)
    assert np.array_equal(eo.x_left, eg.x_left) and np.array_equal(eo.x_right, eg.x_right)
    assert eo.run(oracle, max_steps=30) == 0 and eg.run(cuda_lib, max_steps=30) == 0
    compare(eo, eg)


def test_multi_device_sharding(oracle, cuda_lib):
    """n_devices = 2 inside one call: contiguous member blocks, one host thread + stream per device"""
    if cuda_lib.mbm_device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    pars = []
    for k in range(7):
        p = d.pars_MB_growth(); p.update(V_ip=(-1) ** k * (1 + 0.1 * k) * 1e-7, res=[60 + 5 * k, 90]); pars.append(p)
    eo, _ = d.build_couple_ensemble(pars, oracle)
    eg, _ = d.build_couple_ensemble(pars, cuda_lib)
    assert eo.run(oracle, max_steps=80) == 0
    assert eg.run(cuda_lib, max_steps=80, n_devices=2) == 0
    compare(eo, eg)


def test_bitwise_determinism(cuda_lib):
    """three runs of the same ensemble give identical bits (no order-dependent reductions, no races)"""
    pars = []
    for V, res in ((-1e-7, [100, 150]), (1e-7, [100, 150]), (-2e-7, [60, 90]), (5e-8, [130, 131])):
        p = d.pars_MB_growth(); p.update(V_ip=V, res=res); pars.append(p)
    e0, _ = d.build_couple_ensemble(pars, cuda_lib)
    runs = []
    for _ in range(3):
        e = e0.copy()
        assert e.run(cuda_lib, max_steps=200, flags=abi.FLAG_MASS_FINAL) == 0
        runs.append(e)
    for k in ("x_left", "C_left", "x_right", "C_right", "res", "it", "Ri", "t", "idx_hash", "Mass", "node_steps"):
        assert np.array_equal(runs[0].arrays[k], runs[1].arrays[k]) and np.array_equal(runs[0].arrays[k], runs[2].arrays[k]), k


def test_edge_cases(oracle, cuda_lib):
    """empty ensemble; members already at t_tot or frozen on entry are left untouched"""
    p = d.pars_MB_growth(); p.update(res=[60, 90])
    e, _ = d.build_couple_ensemble([p, p, p], cuda_lib)
    e.arrays["t"][1] = e.arrays["t_tot"][1]            # finished member
    e.arrays["status"][2] = 3                          # frozen member
    before = {k: e.arrays[k].copy() for k in ("x_left", "C_left", "x_right", "C_right", "res", "it")}
    assert e.run(cuda_lib, max_steps=20) == 0
    assert list(e.status) == [0, 0, 3] and list(e.it) == [20, 0, 0]
    for k, v in before.items():
        assert np.array_equal(v[1:], e.arrays[k][1:]), k
    # empty ensemble: the call is a no-op
    import ctypes as C
    c, pp, io = e.structs()
    c.n_members = 0
    assert cuda_lib.mbm_couple_batch(C.byref(c), C.byref(pp), C.byref(io)) == 0
    # bad arguments are library-level errors, not crashes
    c.n_members = 3; c.abi_version = 999
    assert cuda_lib.mbm_couple_batch(C.byref(c), C.byref(pp), C.byref(io)) == -1
    c.abi_version = abi.ABI_VERSION; c.flags = abi.FLAG_SOLVER_FAITHFUL
    assert cuda_lib.mbm_couple_batch(C.byref(c), C.byref(pp), C.byref(io)) == -4     # oracle-only mode


def test_whole_run_MB_growth(oracle, cuda_lib):
    """examples/Diff_couple_MB_growth.jl as shipped, whole run (9 352 steps): end-of-run parity and mass error"""
    eo, eg = run_both([d.pars_MB_growth()], oracle, cuda_lib, 0)
    compare(eo, eg, tol=5e-10)
    assert int(eg.it[0]) == 9352 and list(eg.res[0]) == [200, 398]
    assert abs((eg.Mass[0] - eg.Mass0[0]) / eg.Mass0[0]) < 1e-10


@pytest.mark.parametrize("kind", ["mb", "flux"])
def test_streaming_interleaved_kernel(oracle, cuda_lib, kind):
    """the chunk-interleaved global-memory kernel (long profiles, MRefin = 1: config 5's shape) against the oracle:
    growth and resorption members, node counts that change during the run"""
    pars = []
    for V, res in ((1e-7, [300, 500]), (-1e-7, [300, 500]), (2e-7, [257, 258]), (-3e-7, [400, 401])):
        p = d.pars_MB_growth() if kind == "mb" else d.pars_DCGF()
        p.update(V_ip=V, res=res, MRefin=1.0, Ri=[0.45, 1.0], t_tot=2e4, KD_ar=(0.1, 0.1), Cl_i=0.5, Cr_i=5.0)
        if kind == "flux":
            p.update(n=2, Di=[4e-7, 5e-8], T_ar=(1273.15, 923.15))
        pars.append(p)
    eo, _ = d.build_couple_ensemble(pars, oracle, cap_left=1200, cap_right=1600)
    eg, _ = d.build_couple_ensemble(pars, cuda_lib, cap_left=1200, cap_right=1600)
    assert np.array_equal(eo.x_left, eg.x_left) and np.array_equal(eo.x_right, eg.x_right)
    assert eo.run(oracle, max_steps=120) == 0
    assert eg.run(cuda_lib, max_steps=120, flags=abi.FLAG_FORCE_STREAMING) == 0
    compare(eo, eg)


def test_streaming_kernel_row_batches(oracle, cuda_lib):
    """streaming kernel with columns of ~40 rows: the forward sweep and the back-substitution stage their rows through
    shared memory in batches of 6 (cp.async.bulk + mbarrier), so this run takes several batches per pass and re-uses
    every staging buffer and barrier phase; node counts change on the way (growth and resorption)"""
    pars = []
    for V, res in ((1e-7, [4000, 6000]), (-1e-7, [4100, 5000]), (2e-7, [5000, 5001])):
        p = d.pars_MB_growth()
        p.update(V_ip=V, res=res, MRefin=1.0, Ri=[0.45, 1.0], t_tot=2e4, KD_ar=(0.1, 0.1), Cl_i=0.5, Cr_i=5.0)
        pars.append(p)
    eo, _ = d.build_couple_ensemble(pars, oracle, cap_left=9500, cap_right=9500)
    eg, _ = d.build_couple_ensemble(pars, cuda_lib, cap_left=9500, cap_right=9500)
    assert eo.run(oracle, max_steps=40) == 0
    assert eg.run(cuda_lib, max_steps=40, flags=abi.FLAG_FORCE_STREAMING) == 0
    compare(eo, eg)


def test_D1_material_law(oracle, cuda_lib):
    """examples/D1.jl: Stefan driver with the olivine / melt diffusivities recomputed every step from T and the mean
    olivine composition (MBM_LAW_OLIVINE_MELT_D1).  CUDA `pow`/`exp` and glibc's differ by an ulp in D, which feeds
    dt, and V comes from the solution: values are held to 2e-9 as for the Arrhenius / Stefan fixtures; node counts and
    step counts must agree exactly (the regrid checksum is not compared: with meshes that differ in the last bits the
    equal / clamp branch of the last node is not defined, as in test_stefan_main)."""
    pars = []
    for Tstop, comp in ((1350.0 + 273.0, 0.50), (1330.0 + 273.0, 0.45), (1375.0 + 273.0, 0.55)):
        p = d.pars_D1(); p.update(Tstop=Tstop, CompInt=comp); pars.append(p)
    eo = d.build_stefan_ensemble(pars, oracle)
    eg = d.build_stefan_ensemble(pars, cuda_lib)
    assert eo.run(oracle, max_steps=3000, flags=abi.FLAG_MASS_FINAL) == 0
    assert eg.run(cuda_lib, max_steps=3000, flags=abi.FLAG_MASS_FINAL) == 0
    assert np.array_equal(eo.status, eg.status) and np.all(eo.status == 0)
    for k in ("res", "it", "node_steps"):
        assert np.array_equal(eo.arrays[k], eg.arrays[k]), k
    for m in range(eo.M):
        for a, b in zip(eo.member(m), eg.member(m)):
            assert _relerr(a, b) < 2e-9
    assert _relerr(eo.Ri, eg.Ri) < 1e-10
    assert _relerr(eo.last[:, :2], eg.last[:, :2]) < 1e-12       # D_l, D_r of the last step


def test_pieces_staggered_equal_single_piece(cuda_lib, monkeypatch):
    """host-buffer calls cut into pieces on their own streams (input copies staggered piece by piece, as large shards are
    by default) return the same bits as one piece, with and without the staggering"""
    pars = []
    for k in range(37):                           # not a multiple of the piece count: the last piece is short
        p = d.pars_MB_growth(); p.update(V_ip=(-1e-7 if k % 2 else 1e-7) * (1.0 + 0.01 * k), res=[60 + k, 90 + k])
        pars.append(p)
    e0, _ = d.build_couple_ensemble(pars, cuda_lib)
    runs = []
    for pieces, nostagger in (("1", None), ("4", None), ("8", None), ("4", "1")):
        monkeypatch.setenv("MBM_PIECES", pieces)
        if nostagger: monkeypatch.setenv("MBM_NO_STAGGER", nostagger)
        else: monkeypatch.delenv("MBM_NO_STAGGER", raising=False)
        e = e0.copy()
        assert e.run(cuda_lib, max_steps=60, flags=abi.FLAG_MASS_FINAL) == 0
        assert np.all(e.status == 0)
        runs.append(e)
    for k in ("x_left", "C_left", "x_right", "C_right", "res", "it", "Ri", "t", "idx_hash", "Mass", "node_steps"):
        for r in runs[1:]:
            assert np.array_equal(runs[0].arrays[k], r.arrays[k]), k


def test_checksum_is_optional_and_does_not_change_values(cuda_lib):
    """io.idx_hash = NULL (what bench.py passes) selects the pchip transfer compiled without the checksum bookkeeping:
    every other output must be bit-identical to a run with the checksum"""
    pars = []
    for V, res in ((-1e-7, [100, 150]), (1e-7, [100, 150]), (-2e-7, [60, 90]), (5e-8, [130, 131])):
        p = d.pars_MB_growth(); p.update(V_ip=V, res=res); pars.append(p)
    e0, _ = d.build_couple_ensemble(pars, cuda_lib)
    ea, eb = e0.copy(), e0.copy()
    eb.arrays["idx_hash"] = None
    assert ea.run(cuda_lib, max_steps=300, flags=abi.FLAG_MASS_FINAL) == 0
    assert eb.run(cuda_lib, max_steps=300, flags=abi.FLAG_MASS_FINAL) == 0
    assert np.all(ea.status == 0) and np.all(eb.status == 0)
    for k in ("x_left", "C_left", "x_right", "C_right", "res", "it", "Ri", "t", "Mass", "node_steps"):
        assert np.array_equal(ea.arrays[k], eb.arrays[k]), k


def test_newton_warm_flag_matches_oracle(oracle, cuda_lib):
    """MBM_FLAG_NEWTON_WARM: warm-start calls take the kernel path that starts newton_solver from the member's
    previous root, as the oracle does (mbm_oracle.c, regrid) - same iterates, so the 1e-10 bar holds with the flag
    set, and the result differs from the cold-start run (the flag is not silently ignored)."""
    pars = []
    for V in (-1e-7, 1e-7):
        p = d.pars_MB_growth(); p.update(V_ip=V, res=[100, 150]); pars.append(p)
    p = d.pars_B5(); pars_flux = [p]
    for plist in (pars, pars_flux):
        eo, eg = run_both(plist, oracle, cuda_lib, 300, flags=abi.FLAG_NEWTON_WARM)
        compare(eo, eg)
        assert _relerr(eo.arrays["newton_R"], eg.arrays["newton_R"]) < 1e-15
    # the flag is honoured, not ignored: warm and cold start stop at different iterates of the same root
    ew, _ = d.build_couple_ensemble(pars, cuda_lib)
    assert ew.run(cuda_lib, max_steps=300, flags=abi.FLAG_MASS_FINAL | abi.FLAG_NEWTON_WARM) == 0
    ec, _ = d.build_couple_ensemble(pars, cuda_lib)
    assert ec.run(cuda_lib, max_steps=300, flags=abi.FLAG_MASS_FINAL) == 0
    assert not np.array_equal(ec.arrays["newton_R"], ew.arrays["newton_R"])
    assert _relerr(ec.arrays["newton_R"], ew.arrays["newton_R"]) < 1e-8


def test_bad_arguments_are_rejected_before_any_launch(cuda_lib):
    """mbm_couple_batch checks every pointer its interface kind dereferences (a NULL reaching a kernel would be an
    illegal-address fault and a sticky context error): Stefan coefficients, density tables, law parameters, dt_mode."""
    import ctypes as C
    ens, _ = d.build_couple_ensemble([d.pars_DCGMB()], cuda_lib)
    common, params, io = ens.structs(max_steps=2)
    assert cuda_lib.mbm_couple_batch(C.byref(common), C.byref(params), C.byref(io)) == 0
    common.inner_bc = abi.BC_STEFAN; common.dt_mode = abi.DT_STEFAN         # Stefan without coefficient arrays
    assert cuda_lib.mbm_couple_batch(C.byref(common), C.byref(params), C.byref(io)) == -1
    common.inner_bc = abi.BC_MB; common.dt_mode = 7
    assert cuda_lib.mbm_couple_batch(C.byref(common), C.byref(params), C.byref(io)) == -1
    common.dt_mode = abi.DT_FIND_DT; params.nd = 4                          # density table announced, pointers NULL
    assert cuda_lib.mbm_couple_batch(C.byref(common), C.byref(params), C.byref(io)) == -1
    params.nd = 0; params.law = abi.LAW_OLIVINE_MELT_D1                     # law without parameters / outside Stefan
    assert cuda_lib.mbm_couple_batch(C.byref(common), C.byref(params), C.byref(io)) == -1
    params.law = 0
    assert cuda_lib.mbm_couple_batch(C.byref(common), C.byref(params), C.byref(io)) == 0   # the context is still healthy


def test_run_ahead_chain_variants(oracle, cuda_lib):
    """the tuned kernel's step records: (a) Arrhenius diffusivities with a moving interface keep the one-pass chain (D(T(t))
    feeds dt), (b) constant diffusivities take the two-pass chain (recurrence by thread 0, history lookups one record per
    thread, bisection instead of the cursor walk) - both against the oracle's step-by-step scalars, node counts changing on
    the way ([60 90] nodes, large steps: rescans of the left mesh, runs that leave the aligned pattern)"""
    # unscaled members: the reference's Arrhenius switch `Di == [-1 -1]` (functions.jl:1367) only fires when `scaling` has
    # not divided Di by Dsc (B2 is unscaled for the same reason); D(1273 K) ~ [4e-7, 5e-8] as in the constant members
    pa = d.pars_MB_growth(); pa.update(res=[60, 90], Di=[-1.0, -1.0], D0=[5.1e-3, 6.3e-4], R=8.314472, Ea1=1e5, Ea2=1e5,
                                       T_ar=(1273.15, 1173.15), scale=False)
    pb = d.pars_MB_growth(); pb.update(res=[60, 90], V_ip=2e-7, scale=False)
    pc = d.pars_MB_growth(); pc.update(res=[60, 90], V_ip=-3e-7, KD_ar=(0.9, 0.3), scale=False)
    eo, eg = run_both([pa, pb, pc], oracle, cuda_lib, 700)
    assert np.all(eo.status == 0), eo.status
    assert len({tuple(r) for r in eo.res}) > 1
    compare(eo, eg, tol=2e-9)          # member 0: CUDA exp vs glibc exp (1-ulp D differences feed dt)
    for m in (1, 2):
        for a, b in zip(eo.member(m), eg.member(m)):
            assert _relerr(a, b) < TOL


def test_history_failures_inside_the_run_ahead_chain(oracle, cuda_lib):
    """update_t_dependent_param!'s throw sites reached in the middle of a batch of step records: the time leaves the
    history (functions.jl:571, status 6) - the member stops at the same step with the same t, dt, it as in the oracle,
    the others go on"""
    good = d.pars_MB_growth(); good.update(res=[60, 90])
    bad = d.pars_MB_growth(); bad.update(res=[60, 90])
    eo, _ = d.build_couple_ensemble([good, bad, good], oracle)
    eg, _ = d.build_couple_ensemble([good, bad, good], cuda_lib)
    for e in (eo, eg):                                  # member 1: its own, shorter history (ends after ~40 steps)
        H = e.hist_len
        for k in ("t_ar", "T_ar", "KD_ar"):
            e.arrays[k] = np.repeat(e.arrays[k], 3, axis=0) if e.arrays[k].shape[0] == 1 else e.arrays[k]
        e.arrays["t_ar"][1] = np.linspace(0.0, 40.0 * float(eo.arrays["t_ar"][0, -1]) / 9352.0, H)
    assert eo.run(oracle, max_steps=200, flags=abi.FLAG_MASS_FINAL) == 0
    assert eg.run(cuda_lib, max_steps=200, flags=abi.FLAG_MASS_FINAL) == 0
    assert list(eo.status) == list(eg.status) and eo.status[1] == 6 and eo.status[0] == 0, (eo.status, eg.status)
    for k in ("it", "res", "node_steps"):
        assert np.array_equal(eo.arrays[k], eg.arrays[k]), k
    assert _relerr(eo.t, eg.t) < 1e-13 and _relerr(eo.dt, eg.dt) < 1e-13
    for m in range(3):
        for a, b in zip(eo.member(m), eg.member(m)):
            assert _relerr(a, b) < TOL


def test_step_history_of_a_moving_couple(oracle, cuda_lib):
    """MBM_FLAG_RECORD with the tuned kernel: per-step (Mass, T, KD, Ri[1]) - the `Mass` vector DCGF returns
    (Diff_couple_1D_general_Flux.jl:137) - against the oracle"""
    p = d.pars_MB_growth(); p.update(res=[60, 90], KD_ar=(0.9, 0.6))
    eo, _ = d.build_couple_ensemble([p], oracle)
    eg, _ = d.build_couple_ensemble([p], cuda_lib)
    assert eo.run(oracle, max_steps=120, rec_cap=128) == 0 and eg.run(cuda_lib, max_steps=120, rec_cap=128) == 0
    ro, rg = eo.arrays["rec"][0, :120], eg.arrays["rec"][0, :120]
    assert np.all(ro[:, 0] > 0)
    for c in range(4):
        assert _relerr(ro[:, c], rg[:, c]) < 1e-11, c


def _stream_pass_counts(lib):
    import ctypes as C
    f, t = C.c_uint64(0), C.c_uint64(0)
    assert lib.mbm_stream_pass_counts(0, C.byref(f), C.byref(t), None) == 0
    return int(f.value), int(t.value)


def _stream_cut_rows(lib):
    import ctypes as C
    r = C.c_uint64(0)
    assert lib.mbm_stream_pass_counts(0, None, None, C.byref(r)) == 0
    return int(r.value)


def test_streaming_closed_form_fused_pass(oracle, cuda_lib):
    """streaming kernel for uniform meshes (MRefin = 1): meshes recomputed from the LinRange closed form, back-substitution
    and regrid in one pass where the node counts stay (growth members at config 5's shape: nr = resmin) and the two-pass
    path where regrid! changes them (resorption: nr[2] jumps from 5000 to 7455 at the first step) - both against the
    oracle, integers and the regrid checksum exact"""
    pars = []
    for V, res, r in ((1e-7, [5000, 5000], 0.45), (2e-7, [4200, 4200], 0.41), (-1e-7, [4100, 5000], 0.45),
                      (1.5e-7, [5001, 5003], 0.48)):
        p = d.pars_MB_growth()
        p.update(V_ip=V, res=res, MRefin=1.0, Ri=[r, 1.0], t_tot=2e4, KD_ar=(0.1, 0.1), Cl_i=0.5, Cr_i=5.0)
        pars.append(p)
    eo, _ = d.build_couple_ensemble(pars, oracle, cap_left=9500, cap_right=9500)
    eg, _ = d.build_couple_ensemble(pars, cuda_lib, cap_left=9500, cap_right=9500)
    f0, t0 = _stream_pass_counts(cuda_lib)
    assert eo.run(oracle, max_steps=60) == 0
    assert eg.run(cuda_lib, max_steps=60, flags=abi.FLAG_FORCE_STREAMING) == 0
    f1, t1 = _stream_pass_counts(cuda_lib)
    assert f1 - f0 >= 4 * 55, (f1 - f0, t1 - t0)          # every step whose node counts stay
    assert t1 - t0 >= 1, (f1 - f0, t1 - t0)               # the resorption member's first step
    compare(eo, eg)
    # split calls resume on the closed form (the x arrays a call returns are the LinRanges the next one verifies)
    eg2, _ = d.build_couple_ensemble(pars, cuda_lib, cap_left=9500, cap_right=9500)
    for _ in range(3):
        assert eg2.run(cuda_lib, max_steps=20, flags=abi.FLAG_FORCE_STREAMING) == 0
    f2, _t2 = _stream_pass_counts(cuda_lib)
    assert f2 - f1 >= 3 * 50
    for k in ("x_left", "x_right", "C_left", "C_right", "res", "it", "idx_hash", "node_steps"):
        assert np.array_equal(eg.arrays[k], eg2.arrays[k]), k


def test_streaming_other_meshes_take_the_first_kernel(oracle, cuda_lib):
    """a member whose x arrays are not the LinRange pair (one node perturbed by an ulp) is handed to the streaming kernel
    that keeps x in memory; results still match the oracle on the same input"""
    p = d.pars_MB_growth()
    p.update(V_ip=1e-7, res=[3000, 3000], MRefin=1.0, Ri=[0.45, 1.0], t_tot=2e4, KD_ar=(0.1, 0.1), Cl_i=0.5, Cr_i=5.0)
    eo, _ = d.build_couple_ensemble([p], oracle, cap_left=3100, cap_right=3100)
    eg, _ = d.build_couple_ensemble([p], cuda_lib, cap_left=3100, cap_right=3100)
    for e in (eo, eg):
        e.arrays["x_left"][0, 1500] = np.nextafter(e.arrays["x_left"][0, 1500], 1.0)
    f0, t0 = _stream_pass_counts(cuda_lib)
    assert eo.run(oracle, max_steps=10) == 0
    assert eg.run(cuda_lib, max_steps=10, flags=abi.FLAG_FORCE_STREAMING) == 0
    assert _stream_pass_counts(cuda_lib) == (f0, t0)
    compare(eo, eg)


@pytest.mark.parametrize("kind", ["mb", "flux"])
def test_streaming_fixed_interface_uniform_mesh(oracle, cuda_lib, kind):
    """config 5's fixed-mesh variant (V_ip = 0 on LinRange meshes): the closed-form streaming kernel without the regrid -
    forward sweep, back-substitution in place; against the oracle (no regrid: no checksum terms)"""
    pars = []
    for res, r in (([3000, 3000], 0.45), ([2500, 4000], 0.3), ([2600, 4001], 0.6)):
        p = d.pars_MB_growth() if kind == "mb" else d.pars_DCGF()
        p.update(V_ip=0.0, res=res, MRefin=1.0, Ri=[r, 1.0], t_tot=2e4, KD_ar=(0.1, 0.3), Cl_i=0.5, Cr_i=5.0)
        if kind == "flux":
            p.update(Di=[4e-7, 5e-8], T_ar=(1273.15, 923.15))
        pars.append(p)
    for n in (1, 2, 3):
        for p in pars:
            p.update(n=n)
        eo, _ = d.build_couple_ensemble(pars, oracle, cap_left=4100, cap_right=4100)
        eg, _ = d.build_couple_ensemble(pars, cuda_lib, cap_left=4100, cap_right=4100)
        f0, t0 = _stream_pass_counts(cuda_lib)
        assert eo.run(oracle, max_steps=50) == 0
        assert eg.run(cuda_lib, max_steps=50, flags=abi.FLAG_FORCE_STREAMING) == 0
        f1, t1 = _stream_pass_counts(cuda_lib)
        assert (f1 - f0, t1 - t0) == (0, 150)
        compare(eo, eg)


def test_streaming_closed_form_failure_states(oracle, cuda_lib):
    """members that stop inside a step on the closed-form streaming path: interface too fast in either direction
    (functions.jl:39 / :52), capacity exceeded by regrid!'s node counts, history too short (status 6) - same status, step
    count, t, dt and Ri as the oracle (the reference throws there, so the arrays of a failed member are not defined; the
    member that fails before the advection keeps its last completed state, which is compared), beside members that run on"""
    def mk(**kw):
        p = d.pars_MB_growth()
        p.update(res=[2000, 2000], MRefin=1.0, Ri=[0.45, 1.0], t_tot=2e4, KD_ar=(0.1, 0.1), Cl_i=0.5, Cr_i=5.0, V_ip=1e-7)
        p.update(**kw)
        return p
    pars = [mk(), mk(CFL=8.0, V_ip=-1e-4, t_tot=50.0), mk(CFL=8.0, V_ip=1e-4, t_tot=50.0), mk(V_ip=-1e-7), mk(V_ip=2e-7)]
    # capacities: the resorption member's nr[2] wants 2000 / 0.55 = 3636 nodes at its first regrid
    eo, _ = d.build_couple_ensemble(pars, oracle, cap_left=2100, cap_right=2600)
    eg, _ = d.build_couple_ensemble(pars, cuda_lib, cap_left=2100, cap_right=2600)
    for e in (eo, eg):                                  # member 4: its history ends after a few steps
        for k in ("t_ar", "T_ar", "KD_ar"):
            e.arrays[k] = np.repeat(e.arrays[k], 5, axis=0) if e.arrays[k].shape[0] == 1 else e.arrays[k]
        e.arrays["t_ar"][4] = np.linspace(0.0, 25.0 * float(eo.arrays["t_ar"][0, -1]) / 9352.0 / 10.0, e.hist_len)
    assert eo.run(oracle, max_steps=40) == 0
    assert eg.run(cuda_lib, max_steps=40, flags=abi.FLAG_FORCE_STREAMING) == 0
    assert list(eo.status) == list(eg.status), (eo.status, eg.status)
    assert list(eo.status) == [0, 2, 1, 10, 6], eo.status
    assert np.array_equal(eo.arrays["it"], eg.arrays["it"])
    assert np.array_equal(eo.Ri, eg.Ri) and np.array_equal(eo.t, eg.t) and np.array_equal(eo.dt, eg.dt)
    for m in (0, 4):
        assert np.array_equal(eo.res[m], eg.res[m]) and eo.arrays["idx_hash"][m] == eg.arrays["idx_hash"][m]
        n1, n2 = (int(v) for v in eo.res[m])
        assert np.array_equal(eo.x_left[m, :n1], eg.x_left[m, :n1]), m
        assert np.array_equal(eo.x_right[m, :n2], eg.x_right[m, :n2]), m
        assert _relerr(eo.C_left[m, :n1], eg.C_left[m, :n1]) < TOL and _relerr(eo.C_right[m, :n2], eg.C_right[m, :n2]) < TOL, m


def test_full_size_properties_c4(oracle, cuda_lib):
    """BASELINE.json configs[3] at its full member size (n = 1024 per phase, growth and resorption members), 300 members x
    400 steps, through size-independent properties: every member valid; the closed system keeps its mass (the mass-balance
    row, functions.jl:1094-1109: |Mass - Mass0| / Mass0, the reference's calc_mass_err, below 1e-9); a call cut into three
    resumed calls returns the same bits; the interface moved by exactly sum(V dt); and two members against the oracle"""
    from mbm_b200 import ensembles as E
    pars = E.c4_mb_growth_resorption(300, res=1024, max_steps=10_000)
    e1, _ = d.build_couple_ensemble(pars, cuda_lib, cap_left=2048, cap_right=2560)
    e2 = e1.copy()
    Ri0 = e1.Ri.copy()
    assert e1.run(cuda_lib, max_steps=400, flags=abi.FLAG_MASS_FINAL) == 0
    assert np.all(e1.status == 0) and np.all(e1.it == 400)
    assert np.max(np.abs(e1.Mass - e1.Mass0) / e1.Mass0) < 1e-9
    for steps in (150, 150, 100):
        assert e2.run(cuda_lib, max_steps=steps, flags=abi.FLAG_MASS_FINAL) == 0
    for k in ("res", "it", "Ri", "t", "dt", "idx_hash", "Mass", "node_steps"):
        assert np.array_equal(e1.arrays[k], e2.arrays[k]), k
    for m in range(e1.M):                                # (beyond res[m] the arrays keep whatever an earlier call stored)
        for a, b in zip(e1.member(m), e2.member(m)):
            assert np.array_equal(a, b), m
    V = np.array([p["V_ip"] for p in pars])
    grew = (e1.Ri[:, 0] - Ri0[:, 0]) * np.sign(V)
    assert np.all(grew > 0) and np.all(e1.Ri[:, 1] == Ri0[:, 1])
    # the solve sees one inserted node per step: between N + 1 and the final N + 1 nodes per step
    assert np.all(e1.node_steps >= 400 * 2049)
    eo, _ = d.build_couple_ensemble(pars[:2], oracle, cap_left=2048, cap_right=2560)
    assert eo.run(oracle, max_steps=400, flags=abi.FLAG_MASS_FINAL) == 0
    for m in range(2):
        assert np.array_equal(eo.res[m], e1.res[m]) and eo.arrays["idx_hash"][m] == e1.arrays["idx_hash"][m]
        for a, b in zip(eo.member(m), e1.member(m)):
            assert _relerr(a, b) < TOL


def test_full_size_properties_c5(oracle, cuda_lib):
    """BASELINE.json configs[4] at its full member size (n = 65 536 per phase, MRefin = 1), 6 members x 12 steps on the
    closed-form streaming kernel: fused pass on every step, mass kept, split calls return the same bits, meshes are
    exactly the LinRanges of the final interface position, one member against the oracle (tolerance 1e-9: the two
    elimination orders differ by ~5e-10 at 131 072 unknowns, profiles/README.md)"""
    from mbm_b200 import ensembles as E
    res = 65536
    pars = E.c5_highres(6, res, max_steps=100)
    e1, _ = d.build_couple_ensemble(pars, cuda_lib, cap_left=res + 4, cap_right=res + 4)
    e2 = e1.copy()
    f0, t0 = _stream_pass_counts(cuda_lib)
    assert e1.run(cuda_lib, max_steps=12, flags=abi.FLAG_MASS_FINAL) == 0
    f1, t1 = _stream_pass_counts(cuda_lib)
    assert (f1 - f0, t1 - t0) == (72, 0)
    assert np.all(e1.status == 0) and np.all(e1.it == 12) and np.all(e1.res == res)
    assert np.max(np.abs(e1.Mass - e1.Mass0) / e1.Mass0) < 1e-9
    for steps in (5, 7):
        assert e2.run(cuda_lib, max_steps=steps, flags=abi.FLAG_MASS_FINAL) == 0
    for k in ("x_left", "C_left", "x_right", "C_right", "res", "it", "Ri", "t", "dt", "idx_hash", "Mass", "node_steps"):
        assert np.array_equal(e1.arrays[k], e2.arrays[k]), k
    for m in range(6):
        assert np.array_equal(e1.x_left[m, :res], _linrange(0.0, e1.Ri[m, 0], res))
        assert np.array_equal(e1.x_right[m, :res], _linrange(e1.Ri[m, 0], e1.Ri[m, 1], res))
    eo, _ = d.build_couple_ensemble(pars[:1], oracle, cap_left=res + 4, cap_right=res + 4)
    assert eo.run(oracle, max_steps=12, flags=abi.FLAG_MASS_FINAL) == 0
    assert np.array_equal(eo.res[0], e1.res[0]) and eo.arrays["idx_hash"][0] == e1.arrays["idx_hash"][0]
    assert np.array_equal(eo.x_left[0], e1.x_left[0]) and np.array_equal(eo.x_right[0], e1.x_right[0])
    for a, b in zip(eo.member(0), e1.member(0)):
        assert _relerr(a, b) < 1e-9


def _linrange(a, b, n):
    """Julia's LinRange(a, b, n) (Base.lerpi): t = i / (n - 1), (1 - t) * a + t * b, every operation rounded once"""
    t = np.arange(n, dtype=np.float64) / np.float64(n - 1)
    return (1.0 - t) * a + t * b


def test_step_history_on_the_global_memory_path(oracle, cuda_lib):
    """MBM_FLAG_RECORD with members that do not fit shared memory (forced here): the streaming kernels keep no per-step
    history, so the call stays with the general kernel - the `Mass` vector DCGF returns (Diff_couple_1D_general_Flux.jl:137)
    must not be silently dropped"""
    p = d.pars_MB_growth(); p.update(res=[300, 400], MRefin=1.0, KD_ar=(0.9, 0.6))
    eo, _ = d.build_couple_ensemble([p], oracle, cap_left=700, cap_right=900)
    eg, _ = d.build_couple_ensemble([p], cuda_lib, cap_left=700, cap_right=900)
    assert eo.run(oracle, max_steps=40, rec_cap=64) == 0
    assert eg.run(cuda_lib, max_steps=40, rec_cap=64, flags=abi.FLAG_FORCE_STREAMING) == 0
    ro, rg = eo.arrays["rec"][0, :40], eg.arrays["rec"][0, :40]
    assert np.all(ro[:, 0] > 0)
    for c in range(4):
        assert _relerr(ro[:, c], rg[:, c]) < 1e-11, c
    compare(eo, eg)


@pytest.mark.parametrize("n", [1, 3])
def test_streaming_closed_form_other_geometries(oracle, cuda_lib, n):
    """closed-form streaming kernel, planar and spherical members (the element matrices of functions.jl:525-540 for
    n = 1 / 3), flux members with a Dirichlet outer boundary on either side (set_outer_bc!, :1230-1243) beside closed
    mass-balance members: fused pass on the steps whose node counts stay, against the oracle"""
    pars = []
    for kind, V, bcout in (("mb", 1e-7, [0, 0]), ("mb", -1e-7, [0, 0])):
        p = d.pars_MB_growth()
        p.update(V_ip=V, res=[3000, 3000], MRefin=1.0, Ri=[0.45, 1.0], t_tot=2e4, KD_ar=(0.1, 0.1), Cl_i=0.5, Cr_i=5.0, n=n)
        pars.append(p)
    eo, eg = None, None
    eo, _ = d.build_couple_ensemble(pars, oracle, cap_left=5600, cap_right=5600)
    eg, _ = d.build_couple_ensemble(pars, cuda_lib, cap_left=5600, cap_right=5600)
    f0, _t0 = _stream_pass_counts(cuda_lib)
    assert eo.run(oracle, max_steps=50) == 0
    assert eg.run(cuda_lib, max_steps=50, flags=abi.FLAG_FORCE_STREAMING) == 0
    assert _stream_pass_counts(cuda_lib)[0] - f0 >= 90
    compare(eo, eg)
    fl = []
    # (a Dirichlet condition at the left outer boundary is only accepted for planar members, Diff_couple_1D_general_Flux.jl:80)
    for V, bcout in ((1e-7, [1, 0] if n == 1 else [0, 0]), (-1e-7, [0, 1]), (2e-7, [1, 1] if n == 1 else [0, 1])):
        p = d.pars_DCGF()
        p.update(V_ip=V, res=[3000, 3000], MRefin=1.0, Ri=[0.45, 1.0], t_tot=2e4, KD_ar=(0.1, 0.1), Cl_i=0.5, Cr_i=5.0, n=n,
                 Di=[4e-7, 5e-8], T_ar=(1273.15, 923.15), BCout=bcout)
        fl.append(p)
    eo, _ = d.build_couple_ensemble(fl, oracle, cap_left=5600, cap_right=5600)
    eg, _ = d.build_couple_ensemble(fl, cuda_lib, cap_left=5600, cap_right=5600)
    assert eo.run(oracle, max_steps=50) == 0
    assert eg.run(cuda_lib, max_steps=50, flags=abi.FLAG_FORCE_STREAMING) == 0
    compare(eo, eg)


def test_streaming_fill_in_cutoff_changes_nothing(cuda_lib, monkeypatch):
    """rows beyond the decay length of a chunk's fill-in column are stored without it (|A_k| < 1e-19, checked in the
    sweep): at config 5's member size (columns of 518 rows) the profiles with and without the cutoff agree to the last
    bits, integers and checksum exactly"""
    from mbm_b200 import ensembles as E
    res = 65536
    pars = E.c5_highres(3, res, max_steps=100)
    ea, _ = d.build_couple_ensemble(pars, cuda_lib, cap_left=res + 4, cap_right=res + 4)
    eb = ea.copy()
    r0 = _stream_cut_rows(cuda_lib)
    assert ea.run(cuda_lib, max_steps=8) == 0
    r1 = _stream_cut_rows(cuda_lib)
    assert r1 - r0 >= 3 * 7 * 200, r1 - r0            # every step but the first stores > 200 of 518 rows per column as (R, G)
    monkeypatch.setenv("MBM_S2_NO_CUTOFF", "1")
    assert eb.run(cuda_lib, max_steps=8) == 0
    monkeypatch.delenv("MBM_S2_NO_CUTOFF")
    assert _stream_cut_rows(cuda_lib) == r1
    for k in ("res", "it", "idx_hash", "node_steps", "x_left", "x_right", "t", "Ri"):
        assert np.array_equal(ea.arrays[k], eb.arrays[k]), k
    for m in range(3):
        for a, b in zip(ea.member(m), eb.member(m)):
            assert _relerr(a, b) < 1e-15


def test_lasaga_interface_rows(oracle, cuda_lib):
    """set_inner_bc_Lasaga! (src/functions.jl:1143-1206; exported, no shipped driver: the loop is DCGF's with this row function,
    SURVEY 8(f) row 4): both interface values prescribed from the old neighbour values by the reference's closed form, restated
    as shipped (sol2 multiplies by the denominator where sol1 divides, :1187).  Fixed and moving interface, planar /
    cylindrical / spherical, against the oracle; the prescribed values are computed with the same operations in the same
    order on both sides, so the interface values agree to the last bits"""
    def mk(**kw):
        p = d.pars_DCGF()
        p.update(kind="lasaga", V_ip=0.0, Di=[4e-8, 5e-8], Cl_i=0.6, Cr_i=0.4, KD_ar=(1.8, 1.2), res=[60, 80], MRefin=5.0, n=1,
                 t_tot=4e5, rho=[1.0, 1.1])
        p.update(**kw)
        return p
    for n, V in ((1, 0.0), (3, 0.0), (2, 1e-8), (2, -1e-8)):
        pars = [mk(n=n, V_ip=V), mk(n=n, V_ip=V, KD_ar=(0.7, 0.9), Cl_i=0.3, Cr_i=0.5, rho=[1.2, 1.0])]
        eo, eg = run_both(pars, oracle, cuda_lib, 150)
        assert np.all(eo.status == 0) and np.all(eo.it == 150), (eo.status, eo.it)
        compare(eo, eg)
        for m in range(2):
            n1 = int(eo.res[m, 0])
            assert eo.C_left[m, n1 - 1] == eg.C_left[m, n1 - 1] or _relerr(eo.C_left[m, :n1], eg.C_left[m, :n1]) < 1e-13


def test_c4_full_length_members(oracle, cuda_lib):
    """BASELINE.json configs[3] as specified - n = 1024 per phase, 10^4 steps - for a growth and a resorption member of the
    synthetic ensemble, against the oracle at the end of the run: step count, node counts and the checksum of all 10^4
    regrids exact, profiles to the 1e-10 of BASELINE.json (measured: 1.3e-12), mass kept to 1e-9"""
    from mbm_b200 import ensembles as E
    pars = E.c4_mb_growth_resorption(2, res=1024, max_steps=10_000)
    eo, _ = d.build_couple_ensemble(pars, oracle, cap_left=2048, cap_right=2560)
    eg, _ = d.build_couple_ensemble(pars, cuda_lib, cap_left=2048, cap_right=2560)
    assert eo.run(oracle, max_steps=10_000, flags=abi.FLAG_MASS_FINAL) == 0
    assert eg.run(cuda_lib, max_steps=10_000, flags=abi.FLAG_MASS_FINAL) == 0
    assert np.all(eo.status == 0) and np.all(eo.it == 10_000)
    worst = compare(eo, eg, tol=TOL)
    print("c4 full length: max relative error", worst)
    assert np.max(np.abs(eg.Mass - eg.Mass0) / eg.Mass0) < 1e-9


def test_members_beyond_two_per_sm_take_the_tuned_kernel(oracle, cuda_lib, monkeypatch):
    """graded members of 3 500 - 4 000 nodes (more than two CTAs per SM hold, and capacities beyond the general kernel's
    shared-memory layout): with host buffers the library knows the node counts and runs the tuned kernel with one CTA per
    SM instead of the general kernel on a global workspace (growth members, and a resorption member whose right phase
    jumps to 1 857 nodes at its first regrid - the reference's Newton solver does not converge much beyond 2 000).
    Against the oracle, and against the global-memory path"""
    pars = []
    for V, res, r in ((1e-7, [1500, 2000], 0.5), (2e-7, [1500, 2000], 0.5), (-1e-7, [1300, 2000], 0.3)):
        p = d.pars_MB_growth(); p.update(V_ip=V, res=res, Ri=[r, 1.0]); pars.append(p)
    eo, eg = run_both(pars, oracle, cuda_lib, 60)
    assert np.all(eo.status == 0), eo.status
    assert eo.res[2, 1] > 1800
    assert cuda_lib.mbm_last_tuned_capacity() > 3500          # one CTA per SM
    compare(eo, eg)
    monkeypatch.setenv("MBM_V2_NO_ONE_CTA", "1")
    eh, _ = d.build_couple_ensemble(pars, cuda_lib)
    assert eh.run(cuda_lib, max_steps=60, flags=abi.FLAG_MASS_FINAL) == 0
    assert cuda_lib.mbm_last_tuned_capacity() == 0
    compare(eo, eh)


def test_config4_members_stay_two_per_sm(cuda_lib):
    """BASELINE.json configs[3] through host buffers (the path bench.py's e2e takes): the resorption members reach 2 887
    of the ~2 944 nodes that two CTAs per SM hold, so the shard must keep the two-per-SM plan at every call of a resumed
    run - a capacity above that means one CTA per SM and half the throughput"""
    from mbm_b200 import ensembles as E
    pars = E.c4_mb_growth_resorption(64, res=1024, max_steps=10_000)
    e, _ = d.build_couple_ensemble(pars, cuda_lib, cap_left=2048, cap_right=2560)
    for _ in range(3):
        assert e.run(cuda_lib, max_steps=100) == 0
        cap = cuda_lib.mbm_last_tuned_capacity()
        assert 2890 <= cap <= 3100, cap
    assert np.all(e.status == 0) and int(np.max(e.res.sum(axis=1))) > 2800


def test_randomised_parity_sweep():
    """tools/fuzz_parity.py with a fixed seed: 30 batches of random couple members (kind, geometry, velocity sign incl. 0,
    radii, node counts up to 6 000 per phase, KD drift, outer boundary conditions, uniform / graded meshes, resident and
    global-memory paths, split calls) through the CUDA library and the oracle - integers, status codes and the regrid
    checksum equal, profiles within 1e-10"""
    import subprocess, sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    out = subprocess.run([sys.executable, str(root / "tools" / "fuzz_parity.py"), "20261020", "30"], capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "30 batches, 0 bad" in out.stdout


def test_randomised_parity_sweep_other_families():
    """tools/fuzz_families.py with a fixed seed: 40 batches over the families the first sweep does not reach - the advective dt
    rule of C1 / B4 / B5, fixed-interface flux / mass-balance / hold / Lasaga members (constant and Arrhenius diffusivities,
    drifting KD, Dirichlet outer boundaries), Stefan sweeps - random sizes, geometries, step counts, split calls"""
    import subprocess, sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    out = subprocess.run([sys.executable, str(root / "tools" / "fuzz_families.py"), "20261020", "40"], capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "40 batches, 0 bad" in out.stdout
