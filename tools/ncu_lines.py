#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump: executed warp instructions and stall
samples per source line, split by opcode class.  Usage: ncu_lines.py dump.csv [top_n] [divisor]"""
import csv, sys, collections, re

def main():
    path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    div = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    rows = list(csv.reader(open(path, newline="")))
    cur_file = None; hdr = None
    per_line = collections.OrderedDict()
    cls_tot = collections.Counter(); op_tot = collections.Counter()
    cur = None
    for r in rows:
        if not r: continue
        if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
        if r[0] == "Function Name": continue
        if r[0] == "Line No": hdr = r; iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); continue
        if hdr is None: continue
        if r[0] != "":
            cur = (cur_file, int(r[0])); per_line.setdefault(cur, dict(src=r[1].strip(), ins=0, smp=0, cls=collections.Counter()))
            continue
        sass = r[3].strip(); 
        try: ins = int(r[iI]); smp = int(r[iS])
        except ValueError: continue
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", sass)
        op = m.group(2) if m else "?"
        if op in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"): c = "fp64"
        elif op.startswith("MUFU"): c = "mufu"
        elif op in ("LDS", "STS", "LDSM"): c = "smem"
        elif op in ("LDG", "STG", "LD", "ST", "LDC", "LDCU", "ATOMS", "ATOMG", "RED", "CCTL"): c = "gmem"
        elif op in ("BAR", "BSSY", "BSYNC", "BRA", "WARPSYNC", "EXIT", "CALL", "RET", "BREAK", "NANOSLEEP", "YIELD", "DEPBAR"): c = "ctrl"
        elif op.startswith("SHFL") or op in ("VOTE", "VOTEU", "REDUX", "MATCH"): c = "shfl"
        else: c = "int/misc"
        d = per_line[cur]; d["ins"] += ins; d["smp"] += smp; d["cls"][c] += ins
        cls_tot[c] += ins; op_tot[op] += ins
    tot = sum(cls_tot.values()); tots = sum(d["smp"] for d in per_line.values())
    print(f"total warp instr {tot/div:.0f}  samples {tots}")
    for c, v in cls_tot.most_common(): print(f"  {c:9s} {v/div:12.0f} {100*v/tot:5.1f}%")
    print("top opcodes:", ", ".join(f"{o}={v/div:.0f}" for o, v in op_tot.most_common(30)))
    print("--- by instructions")
    for k, d in sorted(per_line.items(), key=lambda kv: -kv[1]["ins"])[:top]:
        cl = " ".join(f"{c}={v/div:.0f}" for c, v in d["cls"].most_common(4))
        print(f"{d['ins']/div:9.0f} {100*d['ins']/tot:5.1f}% smp {100*d['smp']/max(tots,1):4.1f}% {k[0]}:{k[1]:4d} [{cl}] {d['src'][:90]}")
    print("--- by samples")
    for k, d in sorted(per_line.items(), key=lambda kv: -kv[1]["smp"])[:top]:
        print(f"{d['smp']:8d} {100*d['smp']/max(tots,1):5.1f}% ins {d['ins']/div:9.0f} {k[0]}:{k[1]:4d} {d['src'][:100]}")
main()
