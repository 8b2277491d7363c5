#!/usr/bin/env python
"""Summarises an .ncu-rep (read here on the CPU box with `ncu -i`): headline metrics, SASS opcode
mix and the per-source-line / per-region share of executed instructions and stall samples.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep [--regions search] > profiles/<name>.txt
"""
from __future__ import annotations

import collections
import csv
import io
import subprocess
import sys

HEADLINE = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "sm__cycles_elapsed.max", "local_load_sectors", "smsp__inst_executed_op_local_ld.sum",
    "smsp__inst_executed_op_local_st.sum",
]


def ncu_csv(rep, *args):
    out = subprocess.run(["ncu", "-i", rep, "--csv", *args], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def num(x):
    try:
        return float(x)
    except Exception:
        return 0.0


def main():
    rep = sys.argv[1]
    raw = ncu_csv(rep, "--page", "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    print("== kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for i, h in enumerate(hdr):
        if h in HEADLINE:
            print(f"{h:85s} {vals[i]:>18s} {units[i]}")
    rows = ncu_csv(rep, "--page", "source", "--print-source", "cuda,sass")
    cur, head, lines, sass = None, None, [], []
    for r in rows:
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            head = r
            continue
        if not head or len(r) < 6:
            continue
        d = dict(zip(head[4:], r[4:]))
        if r[0] not in ("", "Function Name") and r[2] == "-":
            lines.append((cur, int(r[0]), r[1].strip()[:90], d))
        elif r[0] == "" and r[2] not in ("-", "..."):
            sass.append((r[3].strip(), d))
    ti = sum(num(d.get("Instructions Executed")) for *_x, d in lines) or 1.0
    ts = sum(num(d.get("Warp Stall Sampling (All Samples)")) for *_x, d in lines) or 1.0
    print(f"\n== per source line (top 40 by stall samples; total warp-inst {ti:.4g}, samples {ts:.4g})")
    for f, l, s, d in sorted(lines, key=lambda t: -num(t[3].get("Warp Stall Sampling (All Samples)")))[:40]:
        print(f"{f}:{l:4d} samples {num(d.get('Warp Stall Sampling (All Samples)')) / ts * 100:5.1f}% "
              f"inst {num(d.get('Instructions Executed')) / ti * 100:5.1f}% "
              f"lanes {num(d.get('Thread Instructions Executed')) / max(num(d.get('Instructions Executed')), 1):4.1f}  {s}")
    print("\n== per source line (top 30 by executed warp instructions)")
    for f, l, s, d in sorted(lines, key=lambda t: -num(t[3].get("Instructions Executed")))[:30]:
        print(f"{f}:{l:4d} inst {num(d.get('Instructions Executed')) / ti * 100:5.1f}% "
              f"samples {num(d.get('Warp Stall Sampling (All Samples)')) / ts * 100:5.1f}% "
              f"lanes {num(d.get('Thread Instructions Executed')) / max(num(d.get('Instructions Executed')), 1):4.1f}  {s}")
    print("\n== per source file")
    agg = collections.defaultdict(lambda: [0.0, 0.0])
    for f, l, s, d in lines:
        agg[f][0] += num(d.get("Instructions Executed"))
        agg[f][1] += num(d.get("Warp Stall Sampling (All Samples)"))
    for f, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{f:28s} inst {i / ti * 100:5.1f}%  samples {s / ts * 100:5.1f}%")
    # opcode mix from the plain SASS page
    rows = ncu_csv(rep, "--page", "source")
    head = rows[1]
    ix, isrc, ith, ism = (head.index(k) for k in ("Instructions Executed", "Source",
                                                  "Thread Instructions Executed", "# Samples"))
    ops = collections.Counter()
    thr = collections.Counter()
    smp = collections.Counter()
    for r in rows[2:]:
        if len(r) <= ix:
            continue
        parts = r[isrc].split()
        if not parts:
            continue
        op = (parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]).split(".")[0]
        ops[op] += num(r[ix])
        thr[op] += num(r[ith])
        smp[op] += num(r[ism])
    tot = sum(ops.values()) or 1.0
    st = sum(smp.values()) or 1.0
    print(f"\n== SASS opcode mix (static instructions {len(rows) - 2})")
    for op, n in ops.most_common(24):
        print(f"{op:10s} inst {n / tot * 100:5.1f}%  lanes/inst {thr[op] / max(n, 1):5.1f}  samples {smp[op] / st * 100:5.1f}%")


if __name__ == "__main__":
    main()
