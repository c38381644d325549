#!/usr/bin/env python
"""Per-source-line executed warp-instructions and stall samples from an ncu report captured with --import-source on.
Usage: ncu_lines.py <report.ncu-rep> [top-n]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None; cur_file = None
agg = collections.OrderedDict()
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; ist = r.index("Warp Stall Sampling (All Samples)"); ie = r.index("Instructions Executed"); continue
    if hdr is None or len(r) < len(hdr) or r[0] == "": continue
    try: ln = int(r[0]); st = int(r[ist] or 0); ex = int(r[ie] or 0)
    except ValueError: continue
    k = (cur_file, ln); a = agg.setdefault(k, [0, 0, r[1].strip()[:110]]); a[0] += st; a[1] += ex
tot_s = sum(a[0] for a in agg.values()) or 1; tot_e = sum(a[1] for a in agg.values()) or 1
print(f"total stall samples {tot_s}, executed warp-instr {tot_e}")
print("---- by executed instructions ----")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"instr {100*a[1]/tot_e:5.1f}%  stall {100*a[0]/tot_s:5.1f}%  {k[0]}:{k[1]}  {a[2]}")
print("---- by stall samples ----")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top // 2]:
    print(f"stall {100*a[0]/tot_s:5.1f}%  instr {100*a[1]/tot_e:5.1f}%  {k[0]}:{k[1]}  {a[2]}")
