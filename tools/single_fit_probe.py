#!/usr/bin/env python3
"""One Nelder-Mead fit of BASELINE config 2 stepped on the device (Engine.nelder_mead), for a launch list under ncu:
per round misti_nm_propose_kernel, misti_correct_kernel (cooperative), misti_jsfs_kernel, misti_stiff_kernel,
misti_nm_apply_kernel.  Prints the fit as JSON."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import misti_b200  # noqa: E402


def main():
    with open(os.path.join(ROOT, "tests", "golden", "datasets.json")) as f:
        ds = json.load(f)["datasets"]["synthetic"]
    eng = misti_b200.Engine(0)
    gid = eng.add_grid(ds["times"], ds["lambdas"])
    mid = eng.add_model(gid, 40, 0, bands=[(1, 5, 12, 0.8, 0)])
    eng.set_data([ds["sfs"]], True)
    flags = misti_b200.FLAG_CORRECT | misti_b200.FLAG_CPFIT | misti_b200.FLAG_SMOOTH | misti_b200.FLAG_UNFOLDED
    x0, one, zero = np.array([[0.8]]), np.array([mid], dtype=np.int32), np.zeros(1, dtype=np.int32)
    eng.nelder_mead(x0, one, zero, flags=flags, maxiter=1000)
    t = time.perf_counter()
    r = eng.nelder_mead(x0, one, zero, flags=flags, maxiter=1000)
    dt = time.perf_counter() - t
    print(json.dumps({"seconds": dt, "x": r["x"][0].tolist(), "llh": float(-r["fun"][0]), "nfev": int(r["nfev"][0]),
                      "rounds": r["launches"], "graph": r["graph"]}))


if __name__ == "__main__":
    main()
