#!/usr/bin/env python3
"""Latency of one batched evaluation as a function of the batch size (what an optimiser iteration costs):
wall time per Engine.evaluate call and the device time of the two kernels.  Run under gpurun; prints JSON."""
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
    models = {"config2": eng.add_model(gid, 40, 0, bands=[(1, 5, 12, 0.8, 0)]),
              "config3": eng.add_model(gid, 40, 0, bands=[(0, 2, 10, 0.3, 0), (1, 5, 12, 0.8, 1)], pulses=[(0, 7, 0.05, 2)]),
              "config5b": eng.add_model(gid, 40, 0, bands=[(0, 4, 40, 3.0, 0)])}
    eng.set_data([ds["sfs"]], True)
    flags = misti_b200.FLAG_CORRECT | misti_b200.FLAG_CPFIT | misti_b200.FLAG_SMOOTH | misti_b200.FLAG_UNFOLDED
    rng = np.random.default_rng(5)
    out = {}
    for name, mid in models.items():
        for B in (1, 64, 1024, 4096, 16384, 65536):
            p = np.column_stack([rng.uniform(0, 3, B), rng.uniform(0, 3, B), rng.uniform(0, 0.3, B)])
            for _ in range(3):
                eng.evaluate(p, model=mid, flags=flags, want=("status",))
            n = 20
            t = time.perf_counter()
            k1 = k2 = 0.0
            for _ in range(n):
                eng.evaluate(p, model=mid, flags=flags, want=("status",))
                a, b = eng.last_kernel_ms()
                k1 += a
                k2 += b
            dt = (time.perf_counter() - t) / n
            out["%s/B=%d" % (name, B)] = {"wall_ms": 1e3 * dt, "k1_ms": k1 / n, "k2_ms": k2 / n}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
