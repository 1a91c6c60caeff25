#!/usr/bin/env python3
"""Run-away golden cases (rates ~1e7: dense scaling-and-squaring step) on the device, JSFS stage given the reference's rates,
against the 50-digit values of tests/golden/stiff_exact.json (tools/exact_jsfs.py) and against the reference's float64 result."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import misti_b200
from _cases import bands_pulses, flags_of, grid_of, sfs_of
ds = json.load(open(os.path.join(ROOT, "tests", "golden", "datasets.json")))["datasets"]
cases = {c["name"]: c for c in json.load(open(os.path.join(ROOT, "tests", "golden", "evals.json")))["cases"]}
exact = json.load(open(os.path.join(ROOT, "tests", "golden", "stiff_exact.json")))["cases"]
eng = misti_b200.Engine(0)
out = []
for ex in exact:
    case = cases[ex["name"]]
    times, lam, st, sd = grid_of(ds, case)
    bands, pulses = bands_pulses(case)
    eng.clear_models()
    gid = eng.add_grid(times, lam)
    mid = eng.add_model(gid, st, sd, bands, pulses)
    eng.set_data([sfs_of(ds, case)], case["flags"]["unfolded"])
    inj = np.zeros((1, eng.numT_max, 2)); inj[0, :len(lam)] = np.array(case["expect"]["lc"])
    r = eng.evaluate(np.array([case["params"]]), model=mid, flags=flags_of(case), lc_inject=inj, want=("jafs", "status", "terms"))
    je = np.array([float(v) for v in ex["jafs_exact"]])
    out.append({"name": ex["name"], "status": int(r["status"][0]), "terms": int(r["terms"][0]),
                "device_jafs_relerr_vs_exact": float(np.max(np.abs(r["jafs"][0] - je) / je)),
                "device_llh_relerr_vs_exact": abs(r["llh"][0, 0] - ex["llh_exact"]) / abs(ex["llh_exact"]),
                "reference_jafs_relerr_vs_exact": ex["reference_jafs_relerr_vs_exact"], "per_entry": (np.abs(r["jafs"][0] - je) / je).tolist()})
print(json.dumps(out, indent=1))
