#!/usr/bin/env python
"""Summarise an `ncu --set full` report into the markdown table and traffic.json kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_r01.ncu-rep profiles/r01_ncu_summary.md profiles/traffic.json

Reads the report with `ncu -i ... --page raw --csv` (works without a GPU).  One row per profiled
launch; traffic.json maps kernel name -> dram bytes (read + write) per launch, which bench.py reports
as roofline.traffic.
"""
import csv
import json
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time ms", lambda v, u: v * {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3,
                                                            "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(u, 1.0)),
    ("smsp__inst_executed.sum", "warp inst (M)", lambda v, u: v / 1e6),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %", None),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", None),
    ("launch__registers_per_thread", "regs", None),
    ("launch__occupancy_limit_shared_mem", "occ limit smem (CTAs)", None),
    ("dram__bytes_read.sum", "dram read MB", None),
    ("dram__bytes_write.sum", "dram write MB", None),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %", None),
    ("lts__t_sector_hit_rate.pct", "L2 hit %", None),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %", None),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts (M)", lambda v, u: v / 1e6),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait", None),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb", None),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb", None),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier", None),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch", None),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe", None),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected", None),
]


def to_bytes(v, unit):
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def main():
    rep, out_md, out_json = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    lines = ["| kernel | " + " | ".join(m[1] for m in METRICS) + " |", "|---|" + "---|" * len(METRICS)]
    traffic = {}
    seen = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("b200::", "").split("<")[0]
        k = seen.get(name, 0)
        seen[name] = k + 1
        cells = []
        for key, _, conv in METRICS:
            if key not in col or r[col[key]] in ("", "n/a"):
                cells.append("-")
                continue
            v = float(r[col[key]].replace(",", ""))
            u = units[col[key]]
            if key.startswith("dram__bytes"):
                v = to_bytes(v, u) / 1e6
            elif conv:
                v = conv(v, u)
            cells.append(f"{v:.2f}" if abs(v) < 1000 else f"{v:.0f}")
        lines.append(f"| {name} #{k} | " + " | ".join(cells) + " |")
        if k == 0 and "dram__bytes_read.sum" in col:
            rd = to_bytes(float(r[col["dram__bytes_read.sum"]].replace(",", "")), units[col["dram__bytes_read.sum"]])
            wr = to_bytes(float(r[col["dram__bytes_write.sum"]].replace(",", "")), units[col["dram__bytes_write.sum"]])
            traffic[name] = int(rd + wr)
    open(out_md, "a").write("\n".join(lines) + "\n")
    old = {}
    try:
        old = json.load(open(out_json))
    except Exception:  # noqa: BLE001
        pass
    old.update(traffic)
    json.dump(old, open(out_json, "w"), indent=1, sort_keys=True)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
