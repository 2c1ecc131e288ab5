#!/usr/bin/env python
"""SASS evidence for profiles/: which sm_100a instructions the shipped library uses, per kernel (no GPU needed).

    python tools/sass_evidence.py > profiles/rNN_sass_evidence.txt
"""
import collections
import hashlib
import os
import re
import subprocess

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
LIB = os.path.join(ROOT, "deflate.hpp_b200", "libb200deflate.so")
COLS = [("UBLKCP", "TMA 1-D bulk copy"), ("LDGSTS", "cp.async"), ("SYNCS", "mbarrier"), ("MATCH", "match.any"), ("VOTE", "ballot / any"),
        ("SHFL", "shuffle"), ("ATOMS", "shared-memory atomics"), ("ATOMG", "global atomics"), ("REDUX", "warp reduce"), ("LDS", ""), ("STS", ""),
        ("LDG", ""), ("STG", ""), ("LOCAL", "LDL + STL")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    print(f"# cuobjdump -sass deflate.hpp_b200/libb200deflate.so   sha256 {hashlib.sha256(open(LIB, 'rb').read()).hexdigest()[:16]}   HEAD {head}")
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    print(f"# cubin architectures: {', '.join(arch)}")
    print("# columns: instructions, then occurrences of " + ", ".join(f"{a}{' (' + b + ')' if b else ''}" for a, b in COLS))
    kernels = {}
    name = None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m and name:
            op = m.group(1)
            c = kernels[name]
            c["n"] += 1
            if op in ("LDL", "STL"):
                c["LOCAL"] += 1
            elif op in ("VOTE", "VOTEU"):
                c["VOTE"] += 1
            elif op in ("ATOM", "ATOMG", "RED"):
                c["ATOMG"] += 1
            else:
                c[op] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    total = collections.Counter()
    rows = []
    for (k, c), d in zip(kernels.items(), demangled):
        short = re.sub(r"\(.*", "", d.replace("(anonymous namespace)::", "")).replace("b200::", "").replace("void ", "")
        rows.append((short, c))
        total.update(c)
    w = max(len(r[0]) for r in rows)
    print(f"{'kernel':{w}} {'inst':>6} " + " ".join(f"{a:>6}" for a, _ in COLS))
    for short, c in sorted(rows):
        print(f"{short:{w}} {c['n']:6d} " + " ".join(f"{c[a]:6d}" for a, _ in COLS))
    print(f"{'TOTAL':{w}} {total['n']:6d} " + " ".join(f"{total[a]:6d}" for a, _ in COLS))


if __name__ == "__main__":
    main()
