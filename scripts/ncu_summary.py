#!/usr/bin/env python
"""Summarises an .ncu-rep (ncu --set full --import-source on) into text: headline metrics, stall reasons, hottest source lines.

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.summary.txt
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.avg.per_second"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    h = rows[0]
    for k_row in rows[2:]:
        name = k_row[h.index("Kernel Name")] if "Kernel Name" in h else "?"
        print(f"== {name}")
        for i, k in enumerate(h):
            if k in WANT:
                print(f"  {k:70s} {k_row[i]:>16s} {rows[1][i]}")
    src = ncu(rep, "--page", "source", "--csv", "--print-source", "cuda,sass")
    stalls, lines, cur, hdr = defaultdict(int), [], None, None
    for r in csv.reader(io.StringIO(src)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]; continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r; continue
        if r[0] == "":
            continue
        try:
            line = int(r[0])
        except ValueError:
            continue
        d = dict(zip(hdr[4:], r[4:]))

        def num(k):
            try:
                return int(d.get(k, "0"))
            except ValueError:
                return 0
        for k in d:
            if k.startswith("stall_") and "Not Issued" not in k:
                stalls[k] += num(k)
        lines.append((cur, line, r[1].strip()[:100], num("# Samples"), num("Instructions Executed")))
    ts, ti = sum(stalls.values()) or 1, sum(x[4] for x in lines) or 1
    tsamp = sum(x[3] for x in lines) or 1
    print("\n== warp stall reasons (share of samples)")
    for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:10]:
        print(f"  {k:28s} {v / ts * 100:5.1f}%")
    per_file = defaultdict(lambda: [0, 0])
    for f, l, s, sm, wi in lines:
        per_file[f][0] += wi; per_file[f][1] += sm
    print("\n== warp instructions / samples per source file")
    for f, (wi, sm) in sorted(per_file.items(), key=lambda kv: -kv[1][0])[:8]:
        print(f"  {f:28s} instr {wi / ti * 100:5.1f}%  samples {sm / tsamp * 100:5.1f}%")
    print(f"\n== hottest source lines (of {ti} warp instructions)")
    for f, l, s, sm, wi in sorted(lines, key=lambda x: -x[4])[:25]:
        print(f"  {f}:{l:<4d} instr {wi / ti * 100:5.2f}% samples {sm / tsamp * 100:5.2f}% | {s}")


if __name__ == "__main__":
    main()
