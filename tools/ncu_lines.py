#!/usr/bin/env python
"""Per-source-line cost of one kernel from an `ncu --set full --import-source on` report.

    python tools/ncu_lines.py <report.ncu-rep> <kernel-name-substring> [top N] [launch index]

ncu's source page (CSV) lists the kernel's SASS with executed-instruction counts and stall samples per instruction;
nvdisasm -g on the cubin inside libb200deflate.so maps instruction offsets to file:line (-lineinfo build).  The two are
joined on the instruction offset (the report's addresses are absolute; the kernel's first instruction is offset 0) and
summed per source line.  Works without a GPU.  The library must be the one the report was captured with.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
LIB = os.path.join(ROOT, "deflate.hpp_b200", "libb200deflate.so")


def line_map(kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    out, cur, inside, name = {}, None, False, None
    for ln in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            name = m.group(1)
            inside = kernel in name
            cur = None
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return out


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    launch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel], capture_output=True, text=True).stdout
    # the CSV holds one table per launch: "Kernel Name",... then a header row that starts with "Address"
    tables, cur = [], None
    for r in csv.reader(raw.splitlines()):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            tables.append(cur)
        elif cur is not None and r:
            cur["rows"].append(r)
    if not tables:
        print("no such kernel in the report")
        return 1
    t = tables[min(launch, len(tables) - 1)]
    hdr = t["rows"][0]
    col = {h: i for i, h in enumerate(hdr)}
    rows = t["rows"][1:]
    a0 = int(rows[0][col["Address"]], 16)
    lm = line_map(kernel)
    agg = {}
    tot_i = tot_s = 0
    miss = 0
    for r in rows:
        off = int(r[col["Address"]], 16) - a0
        inst = int(float(r[col["Instructions Executed"]] or 0))
        smp = int(float(r[col["Warp Stall Sampling (All Samples)"]] or 0))
        exc = int(float(r[col.get("L1 Wavefronts Shared Excessive", 0)] or 0)) if "L1 Wavefronts Shared Excessive" in col else 0
        key, _ = lm.get(off, (None, ""))
        if key is None:
            miss += 1
            key = ("?", 0)
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += inst; a[1] += smp; a[2] += exc
        tot_i += inst; tot_s += smp
    print(f"kernel {t['name'][:100]}\nlaunch {launch} of {len(tables)}; {len(rows)} SASS instructions ({miss} without line info); "
          f"{tot_i / 1e6:.1f} M warp instructions, {tot_s} stall samples")
    src_cache = {}

    def src(key):
        f, n = key
        if f == "?":
            return ""
        if f not in src_cache:
            p = os.path.join(ROOT, "deflate.hpp_b200", "csrc", f)
            src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
        L = src_cache[f]
        return L[n - 1].strip()[:110] if 0 < n <= len(L) else ""
    print(f"{'file:line':28} {'inst %':>7} {'samples %':>9} {'smem exc (M)':>12}  source")
    for key, (i, s, e) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{key[0] + ':' + str(key[1]):28} {100 * i / max(tot_i, 1):7.2f} {100 * s / max(tot_s, 1):9.2f} {e / 1e6:12.2f}  {src(key)}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
