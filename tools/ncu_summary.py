#!/usr/bin/env python
"""Summarise one `ncu --set full --import-source on` capture of a kernel: headline raw metrics, warp stall cycles per
issue, executed warp instructions per unit (divisor = member-steps of the launch) and a table of the segments between
block barriers (instructions, share of instructions, share of stall samples, top stall reasons, top opcodes).
Usage: ncu_summary.py report.ncu-rep <member_steps_of_the_launch>"""
import collections, csv, io, subprocess, sys

rep, units = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit, val = rows[0], rows[1], rows[2]
want = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
print("Kernel:", val[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
for w in want:
    if w in hdr:
        i = hdr.index(w); print(f"{w:90s} {val[i]:>18s} {unit[i]}")
print("\nwarp stall cycles per issued instruction:")
st = [(h, val[i]) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for h, v in sorted(st, key=lambda t: -float(t[1] or 0))[:12]:
    print(f"  {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:22s} {float(v):.2f}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
def f(r, k):
    try: return float(r[ix[k]])
    except ValueError: return 0.0
def op(s):
    t = s.split(); o = t[1] if t[0].startswith("@") else t[0]
    return o.split(".")[0]
stalls = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" not in h]
totS = sum(f(r, "# Samples") for r in data); totI = sum(f(r, "Instructions Executed") for r in data)
ops = collections.Counter()
for r in data: ops[op(r[ix["Source"]])] += f(r, "Instructions Executed")
fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
wf = sum(f(r, "L1 Wavefronts Shared") for r in data); wfx = sum(f(r, "L1 Wavefronts Shared Excessive") for r in data)
print(f"\nexecuted warp instructions per unit: {totI / units:.0f}  (FP64 pipe {fp64 / units:.0f} = {100 * fp64 / totI:.0f} %)")
print("  top opcodes: " + ", ".join(f"{k}={v / units:.0f}" for k, v in ops.most_common(22)))
print(f"  shared-memory wavefronts per unit: {wf / units:.0f}, of which excessive (bank conflicts) {wfx / units:.0f}")
print("\nsegments between block barriers, in program order (instr per unit, share of instr, share of samples | stalls | opcodes):")
seg = []; cur = []
for i, r in enumerate(data):
    cur.append(i)
    if "BAR.SYNC" in r[ix["Source"]] and "0x1," not in r[ix["Source"]]: seg.append(cur); cur = []
seg.append(cur)
for s in seg:
    S = sum(f(data[i], "# Samples") for i in s); I = sum(f(data[i], "Instructions Executed") for i in s)
    if S / totS < 0.004 and I / totI < 0.004: continue
    sc = collections.Counter(); oc = collections.Counter()
    for i in s:
        for h in stalls: sc[h] += f(data[i], h)
        oc[op(data[i][ix["Source"]])] += f(data[i], "Instructions Executed")
    print(f"  sass[{s[0]:5d}..{s[-1]:5d}] {I / units:7.0f} {100 * I / totI:5.1f}% {100 * S / totS:5.1f}% | "
          + ", ".join(f"{k[6:]} {100 * v / max(S, 1):.0f}%" for k, v in sc.most_common(3)) + " | "
          + ", ".join(f"{k}={v / units:.0f}" for k, v in oc.most_common(5)))
