#!/usr/bin/env python
"""Opcode mix, stall mix and hottest source lines of one kernel from an ncu report (source page).

    python profiles/sass_mix.py gpurun_out/prof.ncu-rep field_fwd_kernel [launch-id]
"""
import collections
import os
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"]
    if len(sys.argv) > 3:
        cmd += ["--launch-skip", sys.argv[3], "--launch-count", "1"]
    txt = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    sel = os.environ.get("KSEL", "")  # substring of the full kernel name (template arguments), e.g. KSEL="(bool)0"
    k0 = next(i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and sel in r[1])
    rows = rows[k0:]
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    ops, samp, stalls, lines = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
    linesamp = collections.Counter()
    tot = tots = 0
    cur = "?"
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            if len(r) >= 2 and r[0] == "Kernel Name":
                break  # only the first matching launch
            continue
        src = r[ix["Source"]].strip()
        if not r[ix["Address"]].startswith("0x"):
            cur = src
            continue
        t = src.split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        n, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
        ops[op] += n; samp[op] += s; tot += n; tots += s
        lines[cur] += n; linesamp[cur] += s
        for h in hdr:
            if h.startswith("stall_") and "Not Issued" not in h:
                stalls[h] += int(r[ix[h]])
    print(f"warp instructions {tot}, samples {tots}, SASS lines {len(rows) - hi - 1}")
    for op, n in ops.most_common(28):
        print(f"  {op:10s} {n:12d} {n / tot * 100:5.1f}%   samples {samp[op] / max(tots, 1) * 100:5.1f}%")
    print("stalls:", ", ".join(f"{k[6:]} {v / max(tots, 1) * 100:.1f}%" for k, v in stalls.most_common(10)))
    if len(lines) > 1:
        print("hottest source lines (by samples):")
        for l, s in linesamp.most_common(25):
            print(f"  {s / max(tots, 1) * 100:5.1f}% smp {lines[l] / tot * 100:5.1f}% inst | {l[:130]}")


if __name__ == "__main__":
    main()
