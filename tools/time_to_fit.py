#!/usr/bin/env python3
"""Time-to-fit of the BASELINE optimiser configs on one GPU (run under gpurun); prints one JSON object.

config 2: one Nelder-Mead fit (-uf -mi 2 5 12 0.8 1 --cpfit)            -- latency of a single serial fit
config 3: basin-hopping, W walkers in lock step (two bands + pulse, --cpfit)
config 5: 1001 data rows x split times 36..44
          (a) no migration: 9 chains + 9009 likelihood contractions in ONE launch pair
          (b) -uf -mi 1 4 st 3 1 --cpfit: 9009 Nelder-Mead fits in lock step
The CPU column is the reference's own measured time for the same fit (tests/golden/fits.json, recorded when the
fixtures were generated in the build container: 1 core, scipy 1.18.1) or, where marked, an extrapolation from it.
"""
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


def main():
    walkers = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    niter = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    out = {}
    fits = {f["name"]: f for f in json.load(open(os.path.join(ROOT, "tests", "golden", "fits.json")))["fits"]}
    units = mio.Units.from_file(os.path.join(DATA, "setunits.txt"))
    inp = mio.read_psmc(os.path.join(DATA, "m1.psmc"), os.path.join(DATA, "m2.psmc"), 0, -1, units)
    data = mio.column_sums(mio.read_jafs(os.path.join(DATA, "m.sfs")).jafs)
    bs = mio.read_jafs(os.path.join(DATA, "bs.sfs")).jafs
    eng = misti_b200.Engine(0)

    # warm-up (context, module load)
    sw = Sweep(inp.times, inp.lambdas, [data], unfolded=True, cpfit=True, engine=eng)
    sw.add_model(40, [[2, 5, 12, 0.8, 1]])
    sw.solve()
    for cp in (False, True):  # one-time costs of the other kernel variants / host imports (lgamma) outside the timed regions
        sw = Sweep(inp.times, inp.lambdas, bs[:2], unfolded=cp, cpfit=cp, smooth=True, engine=eng)
        sw.add_model(40)
        sw.evaluate_grid()

    t = time.perf_counter()
    sw = Sweep(inp.times, inp.lambdas, [data], unfolded=True, cpfit=True, engine=eng)
    sw.add_model(40, [[2, 5, 12, 0.8, 1]])
    r = sw.solve(tol=1e-4)
    dt = time.perf_counter() - t
    ref = fits["fit_c2_cpfit"]["expect"]
    out["config2_single_fit"] = {"gpu_s": dt, "x": r["x"][0][:1].tolist(), "llh": float(r["llh"][0]), "nfev": int(r["nfev"][0]),
                                 "launches": r["launches"], "reference_s_1core": ref["seconds"], "reference_x": ref["x"],
                                 "reference_llh": ref["llh"], "reference_nfev": len(ref["calls"])}

    t = time.perf_counter()
    sw = Sweep(inp.times, inp.lambdas, [data], unfolded=True, cpfit=True, engine=eng)
    m = sw.add_model(40, [[1, 2, 10, 0.3, 1], [2, 5, 12, 0.8, 1]], [[1, 7, 0.05, 1]])
    rng = np.random.default_rng(2024)
    # walkers start from random draws m ~ U(0,5), pulse ~ U(0,0.5)
    sw.models[m]["init"] = None
    x0 = np.column_stack([rng.uniform(0, 5, walkers), rng.uniform(0, 5, walkers), rng.uniform(0, 0.5, walkers)])
    from misti_b200.optim import basinhopping_batch

    def fun(X, who):
        llh, _ = sw.evaluate(np.zeros(len(X), dtype=int), X, np.zeros(len(X), dtype=int))
        return -llh
    mids = np.full(walkers, sw.models[m]["id"], dtype=np.int32)

    def on_device(xs):  # the local searches as streams of device launches (Engine.nelder_mead); same decisions as the host driver
        return eng.nelder_mead(xs, mids, np.zeros(walkers, dtype=np.int32), flags=sw.flags)
    mode = os.environ.get("MISTI_TTF_MODE", "device_walkers")  # device_walkers | lockstep_hops | host_driver
    seeds = [2024 + w for w in range(walkers)]
    if mode == "device_walkers":  # round 2: walkers advance independently on the device (misti_fit)
        r = eng.basinhopping(x0, mids, np.zeros(walkers, dtype=np.int32), seeds=seeds, flags=sw.flags, niter=niter, T=0.5, stepsize=0.5)
    else:  # round 1: hops in lock step on the host, local searches on the device or host-driven
        r = basinhopping_batch(fun, x0, niter=niter, T=0.5, seeds=seeds, local_solver=None if mode == "host_driver" else on_device)
    host_driver = mode
    dt = time.perf_counter() - t
    best = int(np.argmin(r["fun"]))
    ref = fits["fit_c3_cpfit"]["expect"]
    out["config3_basinhopping"] = {"gpu_s": dt, "mode": host_driver, "walkers": walkers, "niter": niter, "best_x": r["x"][best].tolist(),
                                   "best_llh": float(-r["fun"][best]), "scipy_nfev_total": int(r["nfev"].sum()),
                                   "device_evaluations": r["evaluations"], "launches": r["launches"],
                                   "reference_single_nelder_mead_s_1core": ref["seconds"], "reference_single_nfev": len(ref["calls"]),
                                   "reference_extrapolated_s_1core": ref["seconds"] / len(ref["calls"]) * int(r["nfev"].sum()),
                                   "reference_local_optimum_llh": ref["llh"]}

    sts = list(range(36, 45))
    first = None
    for _ in range(2):  # the first pass pays one-time costs (lazy loading of a kernel variant, buffer growth); both are reported
        t = time.perf_counter()
        sw = Sweep(inp.times, inp.lambdas, bs, unfolded=False, cpfit=False, smooth=True, engine=eng)
        for st in sts:
            sw.add_model(st)
        llh = sw.evaluate_grid()
        dt = time.perf_counter() - t
        first = dt if first is None else first
    best_st = [sts[i] for i in np.argmax(llh, axis=0)]
    out["config5a_no_migration_grid"] = {"gpu_s": dt, "first_call_s": first, "rows": len(bs), "split_times": sts, "pairs": int(llh.size),
                                         "argmax_split_row0": best_st[0],
                                         "argmax_split_histogram": {str(s): int(best_st.count(s)) for s in sts},
                                         "reference_extrapolated_s_1core": 0.28 * llh.size,
                                         "reference_note": "0.28 s per MiSTI.py evaluation (BASELINE.md 2) x 9009 processes, start-up excluded"}

    t = time.perf_counter()
    sw = Sweep(inp.times, inp.lambdas, bs, unfolded=True, cpfit=True, smooth=True, engine=eng)
    for st in sts:
        sw.add_model(st, [[1, 4, st, 3, 1]])
    r = sw.solve(tol=1e-4)
    dt = time.perf_counter() - t
    ref = fits["fit_c5_band_to_split"]["expect"]
    k0 = [k for k in range(len(r["llh"])) if r["row"][k] == 0 and sw.models[int(r["model"][k])]["splitT"] == 40][0]
    out["config5b_band_to_split_fits"] = {"gpu_s": dt, "fits": len(r["llh"]), "converged": int(r["success"].sum()),
                                          "scipy_nfev_total": int(r["nfev"].sum()), "device_evaluations": r["evaluations"],
                                          "launches": r["launches"], "row0_st40_x": r["x"][k0][:1].tolist(),
                                          "row0_st40_llh": float(r["llh"][k0]), "row0_st40_nfev": int(r["nfev"][k0]),
                                          "reference_row0_st40": {"x": ref["x"], "llh": ref["llh"], "nfev": len(ref["calls"]), "s_1core": ref["seconds"]},
                                          "reference_extrapolated_s_1core": ref["seconds"] / len(ref["calls"]) * int(r["nfev"].sum())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
