#!/usr/bin/env python3
"""A few evaluations of a SMALL batch (config 2, B items) -- the shape of an optimiser step -- for ncu captures of the
cooperative correction kernel and the JSFS kernel at low occupancy.  Usage: small_batch_probe.py [B] [calls]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import misti_b200  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    calls = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    with open(os.path.join(ROOT, "tests", "golden", "datasets.json")) as f:
        ds = json.load(f)["datasets"]["synthetic"]
    eng = misti_b200.Engine(0)
    gid = eng.add_grid(ds["times"], ds["lambdas"])
    mid = eng.add_model(gid, 40, 0, bands=[(1, 5, 12, 0.8, 0)])
    eng.set_data([ds["sfs"]], True)
    flags = misti_b200.FLAG_CORRECT | misti_b200.FLAG_CPFIT | misti_b200.FLAG_SMOOTH | misti_b200.FLAG_UNFOLDED
    p = np.random.default_rng(5).uniform(0, 3, (B, 1))
    for _ in range(calls):
        out = eng.evaluate(p, model=mid, flags=flags, want=("status",))
    print(json.dumps({"B": B, "ok": int((out["status"] == 0).sum()), "kernel_ms": eng.last_kernel_ms()}))


if __name__ == "__main__":
    main()
