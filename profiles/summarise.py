#!/usr/bin/env python
"""Summarises ncu outputs kept in gpurun_out/ into the small text/JSON files committed under profiles/.
  python profiles/summarise.py launches <launches.csv> <out.txt>
  python profiles/summarise.py full <report.ncu-rep> <kernel-substring> <out.json>
"""
import csv
import json
import subprocess
import sys


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = {}
    for r in rows[1:]:
        k = r[ki].split("(")[0].replace("void ", "").replace("unnamed>::", "").strip()
        agg.setdefault(k, []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    lines = ["kernel | launches | total_us | share | avg_us   (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"{k} | {len(v)} | {sum(v) / 1e3:.1f} | {sum(v) / tot:.3f} | {sum(v) / len(v) / 1e3:.1f}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct"]


def full(rep, kernel, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if kernel not in name:
            continue
        d = {"kernel": name[:120]}
        for i, k in enumerate(hdr):
            if k in WANT or ("issue_stalled" in k and "per_issue_active" in k and k.startswith("smsp__average_warps")):
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                if "issue_stalled" in k and v < 0.3:
                    continue
                d[k] = {"value": v, "unit": units[i]}
        res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1)[:1500])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4])
