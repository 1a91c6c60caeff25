#!/usr/bin/env python3
"""Random models through the PAIR-of-lanes JSFS kernel against the CPU oracle (run under gpurun; prints JSON).

tests/_cases.random_jsfs_cases draws grids / models with every segment type and event combination (bands, pulses, sampling
dates, splits at odd places, folded and unfolded spectra); the oracle's corrected rates are injected, so this is the JSFS
stage alone.  Every model is evaluated as a batch of 16 identical items with MISTI_JSFS_PAIR = 1 (one warp of the pair
kernel, all items in lock step) and as one item with MISTI_JSFS_PAIR = 0 (the 16-lane kernel); models with a stiff segment
or an infinite last interval go to the 16-lane kernel through the redo list either way and are counted."""
import json, multiprocessing as mp, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
for _k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
    os.environ.setdefault(_k, "1")


def relerr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def main():
    import misti_b200
    from _cases import bands_pulses, random_jsfs_cases
    from fuzz_parity import oracle_jsfs
    pool = mp.get_context("spawn").Pool(os.cpu_count() or 1)
    engines = {}
    for knob in ("1", "0"):
        os.environ["MISTI_JSFS_PAIR"] = knob
        os.environ["MISTI_DEFER_POST"] = "2"
        engines[knob] = misti_b200.Engine(0)
    del os.environ["MISTI_JSFS_PAIR"], os.environ["MISTI_DEFER_POST"]
    worst = {"1": 0.0, "0": 0.0}
    n, bad, cross = 0, [], 0.0
    for seed in range(1, 11):
        cases = random_jsfs_cases(60, seed=seed)
        refs = pool.map(oracle_jsfs, cases)
        for c, r in zip(cases, refs):
            if r is None:
                continue
            times, lam, st, sd = c["grid"]
            bands, pulses = bands_pulses(c)
            got = {}
            for knob, eng in engines.items():
                eng.clear_models()
                mid = eng.add_model(eng.add_grid(times, lam), st, sd, bands, pulses)
                eng.set_data([c["sfs"]], c["flags"]["unfolded"])
                B = 16 if knob == "1" else 1
                inj = np.zeros((B, eng.numT_max, 2))
                inj[:, :len(r[2])] = np.array(r[2])
                got[knob] = eng.evaluate(np.zeros((B, 0)), model=mid, flags=8 if c["flags"]["unfolded"] else 0, lc_inject=inj,
                                         want=("jafs", "status", "terms"))
            n += 1
            for knob, o in got.items():
                if not (o["status"] == 0).all():
                    bad.append((seed, c.get("name", n), knob, o["status"].tolist()[:2]))
                    continue
                d = max(relerr(o["jafs"][b], r[1]) for b in range(len(o["status"])))
                d = max(d, max(relerr(o["llh"][b, 0], r[0]) for b in range(len(o["status"]))))
                worst[knob] = max(worst[knob], d)
                if d > 1e-9:
                    bad.append((seed, n, knob, d))
            if (got["1"]["status"] == 0).all() and (got["0"]["status"] == 0).all():
                assert (got["1"]["terms"] == got["0"]["terms"][0]).all()
                cross = max(cross, relerr(got["1"]["jafs"][0], got["0"]["jafs"][0]))
                assert all(np.array_equal(got["1"]["jafs"][0], got["1"]["jafs"][b]) for b in range(16))  # lock step leaks nothing
    print(json.dumps({"models": n, "worst_rel_vs_oracle": {"pair_kernel": worst["1"], "16_lane_kernel": worst["0"]},
                      "worst_rel_pair_vs_16_lane": cross, "outside_1e-9_or_failed": bad}, indent=1))


if __name__ == "__main__":
    main()
