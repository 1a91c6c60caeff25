#!/usr/bin/env python3
"""Basin-hopping walkers on the device (config 3 layout): wall time, rounds and points per call -- a probe for the latency of a
round of the on-device optimiser (run under gpurun, optionally under `ncu --metrics gpu__time_duration.sum`)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import misti_b200  # noqa: E402


def main():
    W = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    niter = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    with open(os.path.join(ROOT, "tests", "golden", "datasets.json")) as f:
        ds = json.load(f)["datasets"]["synthetic"]
    eng = misti_b200.Engine(0)
    gid = eng.add_grid(ds["times"], ds["lambdas"])
    mid = eng.add_model(gid, 40, 0, bands=[(0, 2, 10, 0.3, 0), (1, 5, 12, 0.8, 1)], pulses=[(0, 7, 0.05, 2)])
    eng.set_data([ds["sfs"]], True)
    rng = np.random.default_rng(2024)
    x0 = np.column_stack([rng.uniform(0, 5, W), rng.uniform(0, 5, W), rng.uniform(0, 0.5, W)])
    mids = np.full(W, mid, dtype=np.int32)
    out = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = eng.basinhopping(x0, mids, seeds=list(range(W)), flags=1 | 2 | 4 | 8, niter=niter, T=0.5, stepsize=0.5)
        dt = time.perf_counter() - t0
        out.append({"walkers": W, "niter": niter, "s": dt, "rounds": r["launches"], "points": r["evaluations"], "graph": r["graph"],
                    "ms_per_round": 1e3 * dt / max(1, r["launches"]), "scipy_nfev": int(r["nfev"].sum()), "best": float(-r["fun"].min())})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
