#!/usr/bin/env python3
"""config 5b (9 009 Nelder-Mead fits) on the device: wall time per call (run under gpurun; MISTI_FIT_TRACE=file for per-round times)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import misti_b200
from misti_b200 import io as mio
from misti_b200.sweep import Sweep
DATA = os.path.join(ROOT, "data", "synthetic")
units = mio.Units.from_file(os.path.join(DATA, "setunits.txt"))
inp = mio.read_psmc(os.path.join(DATA, "m1.psmc"), os.path.join(DATA, "m2.psmc"), 0, -1, units)
bs = mio.read_jafs(os.path.join(DATA, "bs.sfs")).jafs
eng = misti_b200.Engine(0)
sw = Sweep(inp.times, inp.lambdas, bs, unfolded=True, cpfit=True, smooth=True, engine=eng)
for st in range(36, 45):
    sw.add_model(st, [[1, 4, st, 3, 1]])
out = []
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    t0 = time.perf_counter()
    r = sw.solve(tol=1e-4)
    out.append({"s": time.perf_counter() - t0, "rounds": r["launches"], "points": r["evaluations"], "nfev": int(r["nfev"].sum()),
                "checksum": float(r["llh"].sum())})
print(json.dumps(out))
