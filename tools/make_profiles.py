#!/usr/bin/env python
"""Turn one evidence session's scratch files (gpurun_out/<tag>_*) into the tracked files under profiles/.

    python tools/make_profiles.py <tag> [round label, default r02]

Needs no GPU: reads the .ncu-rep files with `ncu -i` (tools/ncu_summary.py, tools/ncu_lines.py), copies the bench lines
and the launch list, condenses the compute-sanitizer logs, and writes the SASS evidence of the library as it is built now
(the session must have run with this very build: the script refuses if the bench line's library hash differs).
"""
import glob
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")


def sh(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, **kw)


def main():
    tag = sys.argv[1]
    rnd = sys.argv[2] if len(sys.argv) > 2 else "r02"
    os.makedirs(PROF, exist_ok=True)
    notes = []
    # ---- bench lines ----
    for src, dst in ((f"{tag}_bench.json", f"{rnd}_bench_1gpu.json"),):
        p = os.path.join(OUT, src)
        if os.path.exists(p) and os.path.getsize(p):
            shutil.copy(p, os.path.join(PROF, dst))
            notes.append(f"{dst}: the full `python bench.py` line of this session")
    # ---- launch list ----
    p = os.path.join(OUT, f"{tag}_launches.csv")
    if os.path.exists(p):
        shutil.copy(p, os.path.join(PROF, f"{rnd}_launches_bench.csv"))
        notes.append(f"{rnd}_launches_bench.csv: `ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py --steps 1 --warmup 1` "
                     "(cold-cache, serialised: kernel SHARES are what must agree with the event timings)")
    # ---- ncu --set full captures ----
    md = os.path.join(PROF, f"{rnd}_ncu_summary.md")
    reps = sorted(glob.glob(os.path.join(OUT, f"{tag}_ncu_*.ncu-rep")))
    boxmd = os.path.join(OUT, f"{tag}_ncu_summary.md")
    if os.path.exists(boxmd):
        # the session condensed its captures on the GPU box (tools/gpu_session.sh ncu_each): one table from its rows
        lines = [ln for ln in open(boxmd).read().splitlines() if ln.startswith("|")]
        head, rows_ = lines[:2], [ln for ln in lines if not ln.startswith("| kernel") and not ln.startswith("|---")]
        with open(md, "w") as f:
            f.write(f"# Round {rnd[1:]} -- `ncu --set full --clock-control none --import-source on`, one capture per kernel (first launch), B200\n\n"
                    "Each row comes from its own run of `python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --only <section>` under ncu\n"
                    "(`tools/gpu_session.sh ncu_each`), after the same bench command had exited 0 without ncu; the reports were condensed on the box\n"
                    "(`tools/ncu_summary.py`), the lz77_fast_kernel report also came back.  Compress kernels see one batch = 16 384 chunks = 1 GiB per\n"
                    "launch, the inflate kernels the whole 1 GiB stream.  Times are cold-cache and serialised.  `traffic.json` holds dram read + write\n"
                    "bytes of these launches (what `bench.py` reports as `roofline.traffic`).\n\n" + "\n".join(head + rows_) + "\n")
        if os.path.exists(os.path.join(OUT, f"{tag}_traffic.json")):
            shutil.copy(os.path.join(OUT, f"{tag}_traffic.json"), os.path.join(PROF, "traffic.json"))
        for lf in sorted(glob.glob(os.path.join(OUT, f"{tag}_lines_*.txt"))):
            short = os.path.basename(lf)[len(tag) + 7:].replace("_kernel.txt", ".txt")
            shutil.copy(lf, os.path.join(PROF, f"{rnd}_lines_{short}"))
        notes.append(f"{rnd}_ncu_summary.md / traffic.json / {rnd}_lines_<kernel>.txt: {len(rows_)} `--set full` captures, condensed on the box")
    elif reps:
        with open(md, "w") as f:
            f.write(f"# Round {rnd[1:]} -- `ncu --set full --clock-control none --import-source on`, one capture per kernel (first launch), B200\n\n"
                    "Each row comes from its own run of `python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --only <section>` under ncu\n"
                    "(`tools/gpu_session.sh ncu_each`), after the same bench command had exited 0 without ncu.  Compress kernels see one batch =\n"
                    "16 384 chunks = 1 GiB per launch, the inflate kernels the whole 1 GiB stream.  Times are cold-cache and serialised.\n"
                    "`traffic.json` holds dram read + write bytes of these launches (what `bench.py` reports as `roofline.traffic`).\n\n")
        tj = os.path.join(PROF, "traffic.json")
        for r in reps:
            sh([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), r, md, tj])
        notes.append(f"{rnd}_ncu_summary.md / traffic.json: from {len(reps)} `--set full` captures")
        # per-source-line stall profiles of the kernels that matter
        for kern, short in (("lz77_fast_kernel", "lz77_fast"), ("inflate_segments_kernel", "inflate_segments"), ("inflate_copy_kernel", "inflate_copy"),
                            ("encode_kernel", "encode"), ("huffman_kernel", "huffman"), ("lz77_better_kernel", "lz77_better"),
                            ("inflate_batch_kernel", "inflate_batch"), ("foreign_decode_kernel", "foreign_decode")):
            rp = os.path.join(OUT, f"{tag}_ncu_{kern}.ncu-rep")
            if not os.path.exists(rp):
                continue
            r = sh([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rp, kern, "40"])
            if r.stdout.strip():
                with open(os.path.join(PROF, f"{rnd}_lines_{short}.txt"), "w") as f:
                    f.write(f"# tools/ncu_lines.py {os.path.basename(rp)} {kern} 40   (stall samples and executed warp instructions per source line)\n" + r.stdout)
        notes.append(f"{rnd}_lines_<kernel>.txt: per-source-line profiles (tools/ncu_lines.py)")
    # ---- sanitizer ----
    for tool in ("memcheck", "racecheck"):
        log = os.path.join(OUT, f"{tag}_{tool}.log")
        pt = os.path.join(OUT, f"{tag}_{tool}_pytest.log")
        if not os.path.exists(log):
            continue
        txt = open(log, errors="replace").read()
        keep = [ln for ln in txt.splitlines() if re.search(r"SUMMARY|Error|error|hazard|Hazard|Invalid|invalid|Race|race", ln)]
        tail = open(pt, errors="replace").read().strip().splitlines()[-6:] if os.path.exists(pt) else []
        with open(os.path.join(PROF, f"{rnd}_{tool}.txt"), "w") as f:
            f.write(f"# compute-sanitizer --tool {tool} python -m pytest tests -m gpu -x -q -k <subset>   (tools/gpu_session.sh sanitizer)\n"
                    f"# log lines that mention errors / hazards / summaries ({len(txt.splitlines())} lines in the full log):\n")
            f.write("\n".join(keep[:200]) + "\n# pytest:\n" + "\n".join(tail) + "\n")
        notes.append(f"{rnd}_{tool}.txt: condensed compute-sanitizer log")
    # ---- SASS evidence ----
    r = sh([sys.executable, os.path.join(ROOT, "tools", "sass_evidence.py")])
    open(os.path.join(PROF, f"{rnd}_sass_evidence.txt"), "w").write(r.stdout)
    # ---- kernel share check: event timings (bench line) against the ncu launch list ----
    bl = os.path.join(PROF, f"{rnd}_bench_1gpu.json")
    ll = os.path.join(PROF, f"{rnd}_launches_bench.csv")
    if os.path.exists(bl) and os.path.exists(ll):
        import csv
        d = json.loads(open(bl).read().strip().splitlines()[-1])
        ev = d["roofline"]["kernels_ms_per_step"]
        rows = list(csv.reader(x for x in open(ll) if x.startswith('"')))
        hdr = rows[0]
        kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
        mu = hdr.index("Metric Unit")
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
        agg = {}
        seen_main = False
        for r_ in rows[1:]:
            name = r_[kn].split("(")[0].replace("void ", "").replace("b200::", "").split("<")[0]
            # only the headline section (warm-up + one step on the 1 GiB corpus): it ends where bench.py's next section starts
            # (torch kernels checking the round trip, then the small-file calls of config 1)
            if name == "lz77_fast_kernel":
                seen_main = True
            if seen_main and name.startswith("at::"):
                break
            agg[name] = agg.get(name, 0.0) + float(r_[mv].replace(",", "")) * scale.get(r_[mu], 1e-6)
        comp = {"lz77_kernel": agg.get("lz77_fast_kernel", 0), "huffman_kernel": agg.get("huffman_kernel", 0),
                "scan_sizes_kernel": None, "encode_kernel": agg.get("encode_kernel", 0)}
        # the launch list covers warm-up + 1 step of compress (2 passes): shares are scale-free
        tot_ev = sum(v for k, v in ev.items() if comp.get(k) is not None)
        tot_nc = sum(v for v in comp.values() if v is not None)
        with open(os.path.join(PROF, f"{rnd}_share_check.txt"), "w") as f:
            f.write("# share of the compress step per kernel: CUDA events inside bench.py vs the ncu launch list (same command)\n")
            for k, v in comp.items():
                if v is None:
                    continue
                f.write(f"{k:18s} events {ev[k] / tot_ev:6.3f}   ncu {v / tot_nc:6.3f}\n")
        notes.append(f"{rnd}_share_check.txt: kernel shares, events vs ncu")
    print("\n".join(notes))


if __name__ == "__main__":
    main()
