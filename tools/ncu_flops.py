#!/usr/bin/env python3
"""FP64 operations a kernel actually executed, from the SASS page of an .ncu-rep (read here, without a GPU):
thread-level 2*DFMA + DMUL + DADD, per launch and per item.

    python tools/ncu_flops.py <items per launch> <name>=<file.ncu-rep> ... > profiles/r01_executed_flops.json

bench.py uses the per-item figure of the correction kernel (whose work is not counted by the kernel itself; the JSFS
kernel reports its mat-vec count) to state the executed FLOP rate of a step.
"""
import csv
import io
import json
import subprocess
import sys
from collections import Counter


def count(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True,
                         check=True).stdout
    ix, warp, thr = None, Counter(), Counter()
    for r in csv.reader(io.StringIO(out)):
        if r and "Source" in r and "Instructions Executed" in r:
            ix = {h: i for i, h in enumerate(r)}
            continue
        if ix is None or not r:
            continue
        try:
            toks = r[ix["Source"]].split()
            wi = float(r[ix["Instructions Executed"]] or 0)
            ti = float(r[ix["Thread Instructions Executed"]] or 0)
        except (ValueError, IndexError):
            continue
        if not toks:
            continue
        op = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).split(".")[0]
        warp[op] += wi
        thr[op] += ti
    return warp, thr


def main():
    items = int(sys.argv[1])
    out = {"items_per_launch": items, "definition": "thread-level 2*DFMA + DMUL + DADD from the SASS page of the ncu capture"}
    for arg in sys.argv[2:]:
        name, rep = arg.split("=", 1)
        warp, thr = count(rep)
        flops = 2 * thr["DFMA"] + thr["DMUL"] + thr["DADD"]
        out[name] = {"report": rep, "warp_instructions": sum(warp.values()), "dfma_thread": thr["DFMA"], "dmul_thread": thr["DMUL"],
                     "dadd_thread": thr["DADD"], "flops_per_launch": flops, "flops_per_item": flops / items}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
