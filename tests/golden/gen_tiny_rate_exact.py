#!/usr/bin/env python3
"""Tiny positive migration rates (where a fitted rate walks to zero): the reference's result, and the 50-digit value of the
same JSFS stage given the reference's own corrected rates (build container only; minutes).

MigrationInference.SolveDifEq (MigrationInference.py:530-540) integrates with inv(M)(P1 - P0); for m -> 0 the generator
becomes singular (the stationary states), and the float64 result loses ~1e-16 / m.  Output (committed):
tests/golden/tiny_rate_exact.json -- per rate the reference's spectrum / likelihood / rates and the exact spectrum /
likelihood for those rates, so that a test can hold the device to the EXACT value where the reference is off."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, ROOT)
import ref_shim  # noqa: E402

R = ref_shim.load()
from gen_golden import MI, quiet  # noqa: E402
import exact_jsfs  # noqa: E402
from oracle import misti_oracle as mo  # noqa: E402


def main():
    with open(os.path.join(HERE, "datasets.json")) as f:
        d = json.load(f)["datasets"]["synthetic"]
    band = [[2, 5, 12, 0.8, 1]]
    out = {"how": "tests/golden/gen_tiny_rate_exact.py (reference via ref_shim; exact = tools/exact_jsfs.py, mpmath 50 digits)",
           "case": {"dataset": "synthetic", "splitT": 40, "mi": band, "pu": [], "flags": {"smooth": True, "unfolded": True, "trueEPS": False, "cpfit": True}},
           "points": []}
    for m in (1e-14, 1e-12, 1e-10, 1e-8, 1e-6, 1e-4):
        M = quiet(MI, list(d["times"]), [list(v) for v in d["lambdas"]], list(d["sfs"]), 40, [list(map(str, b)) for b in band], [],
                  smooth=True, unfolded=True, trueEPS=False, cpfit=True, sampleDate=0, mixtureTH=0.0)
        llh = quiet(M.JAFSLikelihood, [m])
        lc = [[float(v[0]), float(v[1])] for v in M.lc]
        om = mo.OracleModel(d["times"], d["lambdas"], d["sfs"], 40, band, [], cpfit=True, smooth=True, unfolded=True)
        om.map_parameters([m])
        om.lc = [list(v) for v in lc]
        ex = exact_jsfs.exact_spectrum(om)
        exf = [float(v) for v in ex]
        llh_exact = float(om.score(exf))
        ref_err = float(max(abs(a - b) / b for a, b in zip(M.JAFS, exf)))
        out["points"].append({"m": m, "reference_llh": float(llh), "reference_jafs": [float(v) for v in M.JAFS], "lc": lc,
                              "jafs_exact": [exact_jsfs.mp.nstr(v, 25) for v in ex], "llh_exact": llh_exact,
                              "reference_jafs_relerr_vs_exact": ref_err, "reference_llh_relerr_vs_exact": abs(llh - llh_exact) / abs(llh_exact)})
        print(m, "reference vs exact: jafs", ref_err, "llh", out["points"][-1]["reference_llh_relerr_vs_exact"], flush=True)
    with open(os.path.join(HERE, "tiny_rate_exact.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
