#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the committed summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
    python profiles/summarize.py full gpurun_out/prof_r1.ncu-rep profiles/r1_ncu_full.md
"""
import csv
import subprocess
import sys
from collections import OrderedDict

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum"]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src, errors="ignore")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows:
        if r is hdr or len(r) <= vi or r[ki] == "Kernel Name":
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)  # -> us
        name = r[ki].split("(")[0].replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none) - {src}\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write(f"{sum(a[0] for a in agg.values())} launches, {total / 1e3:.2f} ms in total\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write(f"| `{name[:90]}` | {n} | {t:.1f} | {100 * t / total:.1f}% |\n")


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none - {src}\n\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
            f.write(f"## {name}\n\n| metric | value |\n|---|---|\n")
            for m in METRICS:
                if m in hdr:
                    f.write(f"| {m} | {r[hdr.index(m)]} {units[hdr.index(m)]} |\n")
            f.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
