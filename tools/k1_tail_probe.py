#!/usr/bin/env python3
"""Is the correction kernel bound by its slowest items?  K1 time for 65 536 items with the bench's parameters (m ~ U(0, 5)),
with the same parameter for every item, and the distribution of the solver's evaluation counts."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import misti_b200
ds = json.load(open(os.path.join(ROOT, "tests", "golden", "datasets.json")))["datasets"]["synthetic"]
eng = misti_b200.Engine(0)
gid = eng.add_grid(ds["times"], ds["lambdas"])
m1 = eng.add_model(gid, 40, 0, bands=[(1, 5, 12, 0.8, 0)])
eng.set_data([ds["sfs"]], True)
B = 65536
rng = np.random.default_rng(1234)
out = {}


def t(p, name):
    ts = []
    for _ in range(6):
        o = eng.evaluate(p, model=m1, flags=15, want=("status", "nfev", "terms"))
        ts.append(eng.last_kernel_ms())
    nf = o["nfev"]
    out[name] = {"k1_ms": float(np.median([a for a, _ in ts])), "k2_ms": float(np.median([b for _, b in ts])),
                 "nfev_mean": float(nf.mean()), "nfev_max": int(nf.max()), "nfev_p99": float(np.percentile(nf, 99)),
                 "nfev_hist": np.bincount(nf)[:80].tolist()}


p = rng.uniform(0, 5, (B, 1))
t(p, "uniform_0_5")
t(np.sort(p, axis=0), "uniform_0_5_sorted")
for m in (0.1, 0.8, 2.0, 4.0, 5.0):
    t(np.full((B, 1), m), "all_%g" % m)
t(rng.uniform(0, 1, (B, 1)), "uniform_0_1")
t(rng.uniform(4, 5, (B, 1)), "uniform_4_5")
print(json.dumps(out))
