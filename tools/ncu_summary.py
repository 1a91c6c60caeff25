#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, without a GPU) into a small JSON + per-source-line table for profiles/.

    python tools/ncu_summary.py gpurun_out/r01_jsfs.ncu-rep profiles/r01_jsfs

writes <out>.json (selected raw metrics, stall breakdown, instruction mix) and <out>_lines.csv (top source lines).
"""
import csv
import io
import json
import subprocess
import sys
from collections import Counter

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True, check=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    summary = {"report": rep, "kernel": vals[hdr.index("Kernel Name")], "metrics": {}}
    for i, h in enumerate(hdr):
        if h in KEYS:
            try:
                summary["metrics"][h] = {"value": float(vals[i]), "unit": units[i]}
            except ValueError:
                pass
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    cur, ix, lines, sass = None, None, [], []
    for r in src:
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            ix = {h: i for i, h in enumerate(r)}
            continue
        if not r or r[0] == "Function Name" or ix is None:
            continue
        try:
            inst = float(r[ix["Instructions Executed"]] or 0)
            smp = float(r[ix["# Samples"]] or 0)
        except (ValueError, IndexError):
            continue
        if r[0] != "" and r[2] == "-":
            lines.append((cur, int(r[0]), r[1].strip(), inst, smp))
        elif r[0] == "":
            sass.append((r[3].split()[0] if r[3].split() else "?", r, inst, smp))
    tot_i = sum(l[3] for l in lines) or 1.0
    tot_s = sum(l[4] for l in lines) or 1.0
    mix = Counter()
    for op, r, inst, smp in sass:
        if op.startswith("@") and len(r[3].split()) > 1:
            op = r[3].split()[1]
        mix[op.split(".")[0]] += inst
    tot_m = sum(mix.values()) or 1.0
    summary["instruction_mix_pct"] = {k: round(100 * v / tot_m, 2) for k, v in mix.most_common(16)}
    stall_cols = [h for h in ix if h.startswith("stall_") and "Not Issued" not in h]
    stalls = Counter()
    for op, r, inst, smp in sass:
        for h in stall_cols:
            try:
                stalls[h] += float(r[ix[h]] or 0)
            except (ValueError, IndexError):
                pass
    tot_st = sum(stalls.values()) or 1.0
    summary["stall_samples_pct"] = {k: round(100 * v / tot_st, 1) for k, v in stalls.most_common(8)}
    with open(out + ".json", "w") as f:
        json.dump(summary, f, indent=1)
    with open(out + "_lines.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["file", "line", "pct_instructions", "pct_stall_samples", "source"])
        for l in sorted(lines, key=lambda x: -x[3])[:40]:
            w.writerow([l[0], l[1], round(100 * l[3] / tot_i, 2), round(100 * l[4] / tot_s, 2), l[2][:110]])
    if len(sys.argv) > 3:  # optional: every source line, for aggregation by function
        with open(sys.argv[3], "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["file", "line", "instructions", "stall_samples"])
            for l in lines:
                w.writerow([l[0], l[1], l[3], l[4]])
    print(json.dumps(summary["metrics"], indent=1))
    print(summary["stall_samples_pct"])
    print(summary["instruction_mix_pct"])


if __name__ == "__main__":
    main()
