#!/usr/bin/env python3
"""Per-round anatomy of the lock-step fits (run under gpurun): for config 5b (9 009 Nelder-Mead fits, `-mi 1 4 st 3 1
--cpfit`) and config 3 (basin-hopping walkers) every round of the host driver is recorded -- items submitted, device time
of the correction kernel and of the JSFS + stiff kernels, wall time, and the spread of the per-item solver work (nfev of
the trust-region solves, mat-vec terms) -- to show where a round's time goes and how uneven the items of a round are.
Prints one JSON object."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import misti_b200  # noqa: E402
from misti_b200 import io as mio  # noqa: E402
from misti_b200.sweep import Sweep  # noqa: E402

DATA = os.path.join(ROOT, "data", "synthetic")


def instrument(sw, log):
    eng = sw.engine

    def evaluate(model_idx, params, row_idx):
        model_idx = np.asarray(model_idx, dtype=np.int64).reshape(-1)
        K = model_idx.shape[0]
        P = max([m["n_params"] for m in sw.models] + [0])
        X = np.zeros((K, P))
        params = np.asarray(params, dtype=np.float64).reshape(K, -1) if K else np.zeros((0, P))
        X[:, :params.shape[1]] = params
        mids = sw._model_ids()[model_idx]
        t0 = time.perf_counter()
        out = eng.evaluate(X, model_ids=mids, flags=sw.flags, mixtureTH=sw.mixtureTH, want=("status", "nfev", "terms"),
                           row_ids=np.asarray(row_idx, dtype=np.int32))
        wall = time.perf_counter() - t0
        k1, k2 = eng.last_kernel_ms()
        nf, tm = out["nfev"], out["terms"]
        log.append({"items": int(K), "k1_ms": k1, "k2_stiff_ms": k2, "wall_ms": 1e3 * wall,
                    "nfev_q": [float(v) for v in np.percentile(nf, [50, 90, 99, 100])] if K else None,
                    "terms_q": [float(v) for v in np.percentile(tm, [50, 90, 99, 100])] if K else None,
                    "x_q": [float(v) for v in np.percentile(X[:, 0], [0, 50, 100])] if K and P else None})
        return out["llh"][:, 0], out["status"]
    sw.evaluate = evaluate


def main():
    out = {}
    units = mio.Units.from_file(os.path.join(DATA, "setunits.txt"))
    inp = mio.read_psmc(os.path.join(DATA, "m1.psmc"), os.path.join(DATA, "m2.psmc"), 0, -1, units)
    data = mio.column_sums(mio.read_jafs(os.path.join(DATA, "m.sfs")).jafs)
    bs = mio.read_jafs(os.path.join(DATA, "bs.sfs")).jafs
    eng = misti_b200.Engine(0)
    sts = list(range(36, 45))
    for rep in range(2):
        log = []
        sw = Sweep(inp.times, inp.lambdas, bs, unfolded=True, cpfit=True, smooth=True, engine=eng)
        for st in sts:
            sw.add_model(st, [[1, 4, st, 3, 1]])
        instrument(sw, log)
        t = time.perf_counter()
        r = sw.solve(tol=1e-4, on_device=False)
        dt = time.perf_counter() - t
    out["config5b"] = {"s": dt, "fits": len(r["llh"]), "rounds": len(log), "nit_q": [float(v) for v in np.percentile(r["nit"], [0, 10, 50, 90, 100])],
                       "nfev_scipy_total": int(r["nfev"].sum()), "device_evaluations": int(sum(e["items"] for e in log)),
                       "sum_k1_ms": sum(e["k1_ms"] for e in log), "sum_k2_stiff_ms": sum(e["k2_stiff_ms"] for e in log),
                       "sum_wall_ms": sum(e["wall_ms"] for e in log), "x_final_q": [float(v) for v in np.percentile(r["x"][:, 0], [0, 50, 100])],
                       "log": log}
    # config 3: one local search (Nelder-Mead, N = 3) of W walkers from random starts
    W = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    for rep in range(2):
        log = []
        sw = Sweep(inp.times, inp.lambdas, [data], unfolded=True, cpfit=True, engine=eng)
        m = sw.add_model(40, [[1, 2, 10, 0.3, 1], [2, 5, 12, 0.8, 1]], [[1, 7, 0.05, 1]])
        instrument(sw, log)
        rng = np.random.default_rng(2024)
        x0 = np.column_stack([rng.uniform(0, 5, W), rng.uniform(0, 5, W), rng.uniform(0, 0.5, W)])
        from misti_b200.optim import nelder_mead_batch

        def fun(X, who):
            llh, _ = sw.evaluate(np.zeros(len(X), dtype=int), X, np.zeros(len(X), dtype=int))
            return -llh
        t = time.perf_counter()
        r = nelder_mead_batch(fun, x0, xatol=1e-4, fatol=1e-4, maxiter=600, maxfev=600)
        dt = time.perf_counter() - t
    out["config3_local_search"] = {"s": dt, "walkers": W, "rounds": len(log), "nit_q": [float(v) for v in np.percentile(r["nit"], [0, 10, 50, 90, 100])],
                                   "nfev_scipy_total": int(r["nfev"].sum()), "device_evaluations": int(sum(e["items"] for e in log)),
                                   "sum_k1_ms": sum(e["k1_ms"] for e in log), "sum_k2_stiff_ms": sum(e["k2_stiff_ms"] for e in log),
                                   "sum_wall_ms": sum(e["wall_ms"] for e in log), "log": log}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
