#!/usr/bin/env python
"""bench.py — ensemble node-steps/s (FP64) of the time-stepping hot path on N B200s of one node.

Workload (config.workload): BASELINE.json configs[3], the metric's configuration — a
Diff_couple_1D_general_MB growth/resorption ensemble, n = 1024 nodes per phase, sharded evenly over the
GPUs (8,192 members per GPU = 65,536 / 8, weak scaling: per-GPU work is fixed).  One bench "step" advances
every member by --tsteps time steps (one pass of the hot path over the batch).  The state (>600 MB per GPU)
is larger than L2 (126 MB).

  value    node-steps/s, all ranks, state resident in HBM when the timed region starts
           (MBM_FLAG_DEVICE_POINTERS; CUDA events on the launching stream; max over ranks)
  e2e      the same metric through the reference-facing C-ABI call with HOST (pinned) buffers:
           H2D of the state + kernel + D2H of the result inside the timed region (the library cuts the
           call into 8 pieces whose input copies are staggered, so copies run beside kernels)
  roofline FP64 pipe: algorithmic flops per node-step (100 for MB, n = 2) x node-steps/s
           over the DFMA peak measured live; the HBM figure (32 B/node-step) is reported beside it
  cpu_baseline  the CPU oracle (structured O(N) solve, OpenMP over members, all host cores) on a bounded
           sample of the same ensemble; N = 1, rank 0 only

`--impl reference` times the reference's CPU algorithm (oracle port: the reference is Julia and cannot run
here; dense partial-pivot LU per step as `L_g \\ R_g` does) on the host cores.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
TRAFFIC_BYTES_8192 = 606.5e6 + 737.5e6     # ncu (profiles/r02_ncu_v2kernel_summary.txt), see roofline.traffic below
sys.path.insert(0, str(ROOT))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mbm_b200", choices=["mbm_b200", "reference"])
    ap.add_argument("--members-per-gpu", type=int, default=int(os.environ.get("MBM_BENCH_MEMBERS", 8192)))
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--tsteps", type=int, default=int(os.environ.get("MBM_BENCH_TSTEPS", 250)))
    ap.add_argument("--cpu-members", type=int, default=1024)
    ap.add_argument("--cpu-tsteps", type=int, default=1000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=100):
        # index: one GPU, or a comma-separated list: rank 0 of a multi-GPU run samples every GPU of the job in ONE nvidia-smi
        # process.  Measured on the 8-GPU box: with one poller per rank (eight nvidia-smi processes at 10 Hz) every rank's
        # kernels ran 11 % slower (85.3 ms per step instead of 76.7 ms, same clocks reported); one poller - at any period -
        # costs nothing (profiles/README.md)
        self.index, self.rows, self.proc, self.period_ms = index, [], None, period_ms

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", str(self.period_ms), "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        self.per_gpu = {}
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                self.per_gpu.setdefault(int(r[0]), []).append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def build_ensemble(lib, args, rank):
    from mbm_b200 import drivers as d, ensembles as E
    capL, capR = 2 * args.res, 2 * args.res + args.res // 2          # 2048 / 2560 at res = 1024
    pars = E.c4_mb_growth_resorption(args.members_per_gpu, res=args.res, rank=rank, max_steps=10_000)
    ens, _ = d.build_couple_ensemble(pars, lib, cap_left=capL, cap_right=capR)
    return ens


def host_threads(ora):
    """all the host cores this process may use, set explicitly: torchrun exports OMP_NUM_THREADS=1 to its ranks (libgomp
    reads it when it is loaded), which would time the CPU legs on one thread while reporting the machine's core count.
    Returns the thread count the OpenMP runtime confirms."""
    try:
        want = len(os.sched_getaffinity(0))
    except AttributeError:
        want = os.cpu_count() or 1
    return int(ora.mbm_oracle_set_threads(int(os.environ.get("MBM_BENCH_CPU_THREADS", want))))


def reference_arm(args):
    """oracle port of the reference's CPU algorithm on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import load as oracle_load
    from mbm_b200 import abi
    ora = oracle_load.load()
    cores = host_threads(ora)
    a = argparse.Namespace(**vars(args))
    a.members_per_gpu = max(cores, 8)
    ens = build_ensemble(ora, a, 0)
    tsteps = 2

    def one(flags):
        ns0 = int(ens.node_steps.sum())
        t0 = time.perf_counter()
        assert ens.run(ora, max_steps=tsteps, flags=flags) == 0
        dt = time.perf_counter() - t0
        return (int(ens.node_steps.sum()) - ns0), dt

    for _ in range(args.warmup):
        one(abi.FLAG_SOLVER_FAITHFUL)
    tot_ns, tot_t = 0, 0.0
    for _ in range(args.steps):
        ns, dt = one(abi.FLAG_SOLVER_FAITHFUL)
        tot_ns += ns; tot_t += dt
    value = tot_ns / tot_t
    ns_s, dt_s = one(0)
    ns_s2, dt_s2 = one(0)
    sample = (f"{a.members_per_gpu} members x {tsteps} time steps per bench step, n={args.res}/phase; dense partial-pivot LU "
              "per step with zero-skipping (the reference's `L_g \\ R_g`; LAPACK would do the full O(N^3)), "
              "O(N log N) pchip search (the reference's is O(N^2)) -> conservative (fast) for the reference")
    line = {
        "impl": "reference", "metric": "ensemble node-steps/s (FP64)", "value": value, "unit": "node-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "node-steps/s", "cores": cores, "kind": "port", "sample": sample,
                         "structured_solver_value": (ns_s + ns_s2) / (dt_s + dt_s2)},
        "e2e": {"value": value, "unit": "node-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": "BASELINE.json configs[3]: Diff_couple_1D_general_MB growth/resorption ensemble, "
                        f"n={args.res} nodes/phase, ndim=2, MRefin=20, CFL=0.3, cold-start Newton (reference-literal)",
            "members_per_gpu": args.members_per_gpu, "time_steps_per_bench_step": args.tsteps,
            "caps": [2 * args.res, 2 * args.res + args.res // 2],
            "l2_policy": "state per GPU (~0.6 GB at the default size) is larger than L2 (126 MB); no flush needed",
            "kernel_variant": "resident (one CTA per member, state in shared memory)"}


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
        return
    import torch
    import torch.distributed as dist
    import mbm_b200
    from mbm_b200 import abi, ensembles as E

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = mbm_b200.load_cuda_library()          # raises if the CUDA library is missing: no fallback
    if lib.mbm_device_count() < 1:
        raise RuntimeError("no CUDA device visible")
    ids = np.array([local], dtype=np.int32)

    ens = build_ensemble(lib, args, rank)
    M = ens.M
    # ---------------- device-resident copy of every array (torch owns the memory) -----------------------
    names_in = ["Di", "D0", "Ea", "rho", "V_ip", "t_tot", "CFL", "MRefin", "BCout", "resmin", "t_ar", "T_ar", "KD_ar"]
    names_io = ["x_left", "C_left", "x_right", "C_right", "res", "Ri", "t", "dt", "it", "dx1", "dx2", "Mass0", "Mass",
                "newton_R", "last", "status", "warn", "node_steps"]
    host = {k: ens.arrays[k] for k in names_in + names_io}
    dtens = {k: torch.from_numpy(v).to(dev) for k, v in host.items()}
    snapshot = {k: dtens[k].clone() for k in names_io}

    def structs(ptr_of, flags, max_steps):
        c = abi.Common(abi.ABI_VERSION, M, ens.ndim, ens.hist_len, ens.cap_left, ens.cap_right, ens.inner_bc,
                       ens.dt_mode, flags, 1, max_steps, abi.i32ptr(ids), 0)
        p = abi.CoupleParams()
        io = abi.CoupleIO()
        cast = {np.dtype("float64"): abi._dp, np.dtype("int32"): abi._i32p, np.dtype("int64"): abi._i64p}
        for k in names_in:
            setattr(p, k, C.cast(ptr_of(k), cast[host[k].dtype]))
        p.t_stride = p.T_stride = p.KD_stride = ens.hist_len
        p.R = ens.R
        for k in names_io:
            setattr(io, k, C.cast(ptr_of(k), cast[host[k].dtype]))
        return c, p, io

    cd, pd, iod = structs(lambda k: dtens[k].data_ptr(), abi.FLAG_DEVICE_POINTERS, args.tsteps)
    stream = torch.cuda.current_stream()
    lib.mbm_set_stream(C.c_void_p(stream.cuda_stream))

    def step_device():
        rc = lib.mbm_couple_batch(C.byref(cd), C.byref(pd), C.byref(iod))
        if rc != 0:
            raise RuntimeError(f"mbm_couple_batch rc={rc}")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- warm-up, then EXACTLY K timed steps -------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()
    ns0 = int(dtens["node_steps"].sum().item())
    # one sampler per job: rank 0 watches every GPU of the run (ranks use GPUs 0 .. world-1 of the node)
    period = int(os.environ.get("MBM_BENCH_CLOCK_MS", 200 if world > 1 else 100))
    sampler = ClockSampler(",".join(str(g) for g in range(world)) if world > 1 else local, period)
    if rank == 0 or os.environ.get("MBM_BENCH_SAMPLE_ALL_RANKS"):
        if os.environ.get("MBM_BENCH_SAMPLE_ALL_RANKS"):
            sampler = ClockSampler(local)
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    launches = 0
    for _ in range(args.steps):
        step_device()
        launches += int(lib.mbm_last_kernel_launches())
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if sampler.proc else {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    secs = e0.elapsed_time(e1) * 1e-3
    ns = int(dtens["node_steps"].sum().item()) - ns0
    bad = int((dtens["status"] != 0).sum().item())
    # after the timed region: the reference's own mass diagnostic (calc_mass_err on calc_mass_vol, functions.jl:183-212) for
    # every member of the timed ensemble, evaluated on the device from the state the timed steps left (mbm_validate_batch)
    mass_err = None
    try:
        if os.environ.get("MBM_BENCH_NO_GATE"):
            raise RuntimeError("disabled by MBM_BENCH_NO_GATE")
        verr = torch.zeros(M, device=dev, dtype=torch.float64)
        vout = torch.zeros(2 * M, device=dev, dtype=torch.float64)
        va = abi.ValidateArgs(abi.ABI_VERSION, M, ens.ndim, ens.cap_left, ens.cap_right, abi.VAL_MASS_ERR,
                              abi.FLAG_DEVICE_POINTERS, local)
        for k in ("x_left", "C_left", "x_right", "C_right", "Ri", "rho", "Mass0"):
            setattr(va, k, C.cast(dtens[k].data_ptr(), abi._dp))
        va.res = C.cast(dtens["res"].data_ptr(), abi._i64p)
        va.out, va.relerr = C.cast(vout.data_ptr(), abi._dp), C.cast(verr.data_ptr(), abi._dp)
        if lib.mbm_validate_batch(C.byref(va)) == 0:
            torch.cuda.synchronize()
            ok_members = dtens["status"] == 0
            mass_err = float(verr[ok_members].abs().max().item()) if bool(ok_members.any()) else None
    except Exception as exc:                       # the gate is evidence beside the measurement, never a reason to lose it
        print(f"mass gate skipped: {exc}", file=sys.stderr)
    tsec = torch.tensor([secs], device=dev, dtype=torch.float64)
    tns = torch.tensor([float(ns), float(bad)], device=dev, dtype=torch.float64)
    per_rank = None
    if world > 1:
        # every rank's own time, node-steps and SM clock (its own GPU, sampled during its timed region): shows whether a
        # max-over-ranks time above the single-GPU one is imbalance between the ranks' ensembles or a slower GPU
        mine = torch.tensor([secs, float(ns)], device=dev, dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_gpu = getattr(sampler, "per_gpu", {})
        per_rank = {"ms_per_step": [1e3 * float(a[0].item()) / max(args.steps, 1) for a in allr],
                    "node_steps": [float(a[1].item()) for a in allr],
                    "sm_mhz": [float(np.median(per_gpu[g])) if per_gpu.get(g) else None for g in range(world)]}
        dist.all_reduce(tsec, op=dist.ReduceOp.MAX)
        dist.all_reduce(tns, op=dist.ReduceOp.SUM)
    secs_max, ns_all, bad_all = float(tsec.item()), float(tns[0].item()), int(tns[1].item())
    value = ns_all / secs_max
    kernel_ms = 1e3 * secs / max(args.steps, 1)     # the tuned kernel is >99 % of a step (profiles/r02_ncu_launch_list.csv)

    # ---------------- e2e: host (pinned) buffers through the C ABI, copies inside the timed region -------------
    e2e = None
    lib.mbm_set_stream(None)
    if not args.no_e2e:
        pinned = {k: torch.from_numpy(host[k].copy()).pin_memory() for k in host}
        for k in names_io:
            pinned[k].copy_(snapshot[k].cpu())
        ch, ph, ioh = structs(lambda k: pinned[k].data_ptr(), 0, args.tsteps)
        h2d = sum(pinned[k].numel() * pinned[k].element_size() for k in names_in + names_io if k not in ("dx1", "dx2"))
        d2h = sum(pinned[k].numel() * pinned[k].element_size() for k in names_io if k != "Mass0")

        def step_host():
            rc = lib.mbm_couple_batch(C.byref(ch), C.byref(ph), C.byref(ioh))
            if rc != 0:
                raise RuntimeError(f"mbm_couple_batch (host buffers) rc={rc}")

        step_host()                                     # warm the allocator pool, the piece threads' streams and the pinned pages
        step_host()
        nsh0 = int(pinned["node_steps"].sum().item())
        barrier()
        t0 = time.perf_counter()
        # every call is timed on its own (the call returns after its last device-to-host copy); the e2e time is the MEDIAN call
        # time x the number of calls: about one call in forty of a bench process stalls for ~100 ms on the host side (the same
        # calls repeated 150 times in a process of their own never do: tools/e2e_jitter.py; kernel times are unchanged in
        # the stalled call), and with three to five calls one such stall would halve the figure.  `ms_per_call` keeps them all.
        ke = max(3, min(args.steps, 5))
        caps, per_call, ksum = [], [], []
        for _ in range(ke):
            tc0 = time.perf_counter()
            step_host()
            per_call.append(1e3 * (time.perf_counter() - tc0))
            caps.append(int(lib.mbm_last_tuned_capacity()))
            ksum.append(1e3 * float(lib.mbm_last_kernel_seconds()))
        torch.cuda.synchronize()
        th_all = time.perf_counter() - t0
        th = 1e-3 * float(np.median(per_call)) * ke
        nsh = int(pinned["node_steps"].sum().item()) - nsh0
        te = torch.tensor([th], device=dev, dtype=torch.float64)
        tn = torch.tensor([float(nsh)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dist.all_reduce(tn, op=dist.ReduceOp.SUM)
        e2e = {"value": float(tn.item()) / float(te.item()), "unit": "node-steps/s", "h2d_bytes_per_step": int(h2d) * world,
               "d2h_bytes_per_step": int(d2h) * world, "bytes_are": "whole job (all ranks)", "steps": ke,
               "ms_per_step": 1e3 * float(te.item()) / ke, "timing": "median call time x calls (max over ranks)",
               "ms_per_call": per_call, "ms_per_step_mean_of_all_calls": 1e3 * th_all / ke,
               "tuned_kernel_capacity_per_call": caps, "kernel_ms_summed_over_pieces_per_call": ksum,
               "api": "mbm_couple_batch with host (pinned) buffers: H2D + kernels + D2H inside the call, 8 pieces on their own streams, input copies staggered piece by piece"}
        del pinned

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline ------------------------------------------------------------------------------
    flops_per_ns = E.FLOPS_PER_NODE_STEP[("mb", 2)]
    dfma = float(lib.mbm_measure_dfma_tflops(local))
    per_gpu = value / world
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    roofline = {"bound": "fp64", "achieved": per_gpu * flops_per_ns * 1e-12, "peak": dfma, "unit": "TFLOP/s",
                "frac": (per_gpu * flops_per_ns * 1e-12 / dfma) if dfma > 0 else None,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this workload (8192 members x 250 time steps),
                # `ncu --set full` capture summarised in profiles/r02_ncu_v2kernel_summary.txt; scaled with the members
                "traffic": TRAFFIC_BYTES_8192 * args.members_per_gpu / 8192.0,
                "peak_source": "DFMA micro-benchmark measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                "flops_per_node_step": flops_per_ns, "kernel": "mbm_step_kernel_v2<256,2>", "kernel_ms_per_launch": kernel_ms,
                "traffic_note": "bytes per launch from ncu (606 MB read + 737 MB written at 8192 members, profiles/r02_ncu_v2kernel_summary.txt): the state, the histories and the step records cross HBM once per launch, not once per step (0.3 % of DRAM throughput); algorithmic streaming traffic would be 32 B x node-steps = 157 GB per launch",
                "hbm_view": {"bound": "hbm", "achieved": per_gpu * 32 * 1e-9, "peak": hbm, "unit": "GB/s",
                             "frac": per_gpu * 32 * 1e-9 / hbm, "bytes_per_node_step": 32,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                             "note": "resident kernel: state crosses HBM once per launch, not once per step"},
                # NOT measured in this run: counters of the same kernel on the same workload from the committed capture
                "pipe_view_from_ncu_capture": {"fp64_pipe_cycles_active_pct": 38.8, "issue_slots_active_pct": 50.2,
                                               "shared_memory_pipe_pct": 40.3, "warps_per_sm": 16,
                                               "fp64_instructions_per_node_step": 109,
                                               "source": "profiles/r02_ncu_v2kernel_summary.txt"}}

    # ---------------- CPU baseline (bounded sample, N = 1 only) ------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle import load as oracle_load
        ora = oracle_load.load()
        cpu_threads = host_threads(ora)
        a = argparse.Namespace(**vars(args))
        a.members_per_gpu = args.cpu_members
        ce = build_ensemble(ora, a, 0)
        ce.run(ora, max_steps=2)
        n0 = int(ce.node_steps.sum())
        t0 = time.perf_counter()
        ce.run(ora, max_steps=args.cpu_tsteps)
        tc = time.perf_counter() - t0
        cpu = {"value": (int(ce.node_steps.sum()) - n0) / tc, "unit": "node-steps/s", "cores": cpu_threads,
               "kind": "port",
               "sample": f"{args.cpu_members} members x {args.cpu_tsteps} time steps of the same ensemble, structured O(N) solve, "
                         f"OpenMP schedule(dynamic) over members, {tc:.1f} s"}

    line = {
        "metric": "ensemble node-steps/s (FP64)", "value": value, "unit": "node-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs_max / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": launches, "clocks": clocks, "node_steps_timed": ns_all, "members_failed": bad_all,
        "mass_err_max_abs": mass_err,     # rank 0's members, device-side gate after the timed region (closed system: ~1e-12)
    }
    if per_rank:
        line["per_rank"] = per_rank
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
