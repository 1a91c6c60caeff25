#!/usr/bin/env python3
"""Do interrupted correction chains change any result?  The same walkers with the time slice at 50 us / 350 us / never."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import misti_b200

W = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
niter = int(sys.argv[2]) if len(sys.argv) > 2 else 2
with open(os.path.join(ROOT, "tests", "golden", "datasets.json")) as f:
    ds = json.load(f)["datasets"]["synthetic"]
rng = np.random.default_rng(2024)
x0 = np.column_stack([rng.uniform(0, 5, W), rng.uniform(0, 5, W), rng.uniform(0, 0.5, W)])
res = {}
for us in ("100000000", "50", "350"):
    os.environ["MISTI_FIT_SLICE_US"] = us
    eng = misti_b200.Engine(0)
    gid = eng.add_grid(ds["times"], ds["lambdas"])
    mid = eng.add_model(gid, 40, 0, bands=[(0, 2, 10, 0.3, 0), (1, 5, 12, 0.8, 1)], pulses=[(0, 7, 0.05, 2)])
    eng.set_data([ds["sfs"]], True)
    if niter < 0:
        r = eng.nelder_mead(x0, np.full(W, mid, dtype=np.int32), flags=15, maxiter=600, maxfev=600)
    else:
        r = eng.basinhopping(x0, np.full(W, mid, dtype=np.int32), seeds=list(range(W)), flags=15, niter=niter)
    res[us] = r
    eng.close()
a = res["100000000"]
for us in ("50", "350"):
    b = res[us]
    bad = [w for w in range(W) if not (np.array_equal(a["x"][w], b["x"][w]) and a["fun"][w] == b["fun"][w] and a["nfev"][w] == b["nfev"][w])]
    print(us, "rounds", b["launches"], "vs", a["launches"], "differing fits:", len(bad), bad[:10])
    for w in bad[:3]:
        print("  ", w, a["x"][w], b["x"][w], a["fun"][w], b["fun"][w], a["nfev"][w], b["nfev"][w])
