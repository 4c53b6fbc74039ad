#!/usr/bin/env python
"""Executed warp instructions and stall samples per SOURCE line: joins an ncu source-page export with nvdisasm -g line info
of the object the kernel was built from (same instruction order).

    python profiles/line_mix.py gpurun_out/prof.ncu-rep field_fwd_kernel cednerf_b200/csrc/build/field.o [launch-skip]
"""
import collections, csv, io, os, re, subprocess, sys, tempfile


def main():
    rep, kern, obj = sys.argv[1:4]
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"]
    if len(sys.argv) > 4:
        cmd += ["--launch-skip", sys.argv[4], "--launch-count", "1"]
    rows = list(csv.reader(io.StringIO(subprocess.run(cmd, capture_output=True, text=True).stdout)))
    sel = os.environ.get("KSEL", "")  # substring of the full kernel name (template arguments), e.g. KSEL="(bool)0"
    k0 = next(i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and sel in r[1])
    rows = rows[k0:]
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
    kname = rows[hi - 1][1]
    inst = []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            if r and r[0] == "Kernel Name":
                break
            continue
        inst.append((r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])))
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    short = re.sub(r"[^A-Za-z0-9_]", "", kname.split("(")[0].split("::")[-1].split("<")[0])
    lines, cur, active = [], "?", False
    for l in dis.splitlines():
        if l.startswith("//-") and ".text." in l:
            active = short in l
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = f"{os.path.basename(m.group(1))}:{m.group(2)}"
            m2 = re.findall(r'inlined at "([^"]+)", line (\d+)', l)
            if m2:
                cur += " <- " + " <- ".join(f"{os.path.basename(a)}:{b}" for a, b in m2)
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            lines.append(cur)
    # several template instances may match: keep the first whose length equals the profile's
    if len(lines) != len(inst):
        print(f"warning: {len(lines)} disassembled vs {len(inst)} profiled instructions", file=sys.stderr)
    tot = sum(n for _, n, _ in inst); tots = sum(s for _, _, s in inst)
    by, bys = collections.Counter(), collections.Counter()
    outer, outers = collections.Counter(), collections.Counter()
    for (src, n, s), ln in zip(inst, lines):
        by[ln] += n; bys[ln] += s
        o = ln.split(" <- ")[-1]
        outer[o] += n; outers[o] += s
    print(f"{kname[:80]}: {tot} warp instructions, {tots} samples")
    print("-- by instructions executed")
    for k, n in by.most_common(25):
        print(f"  {n / tot * 100:5.1f}% inst {bys[k] / max(tots,1) * 100:5.1f}% smp | {k[:110]}")
    print("-- by stall samples")
    for k, n in bys.most_common(25):
        print(f"  {by[k] / tot * 100:5.1f}% inst {n / max(tots,1) * 100:5.1f}% smp | {k[:110]}")


if __name__ == "__main__":
    main()
