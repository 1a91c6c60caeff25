#!/usr/bin/env python3
"""A small run through every kernel of the library, twice -- once with small launches (MISTI_MAX_CHUNK = 3000, time slices of
30 us so that chains ARE interrupted) and once with the defaults -- and the two compared bit for bit: chunk boundaries, the
per-row reduction across chunks, the scoring kernel, the post-split kernel in both placements, stiff items, interruptible
chains in fits and walkers.  (compute-sanitizer is closed on this pool: this is the bounds check that can be had.)"""
import json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run():
    import misti_b200
    from misti_b200 import io as mio
    ds = json.load(open(os.path.join(ROOT, "tests", "golden", "datasets.json")))["datasets"]["synthetic"]
    bs = mio.read_jafs(os.path.join(ROOT, "data", "synthetic", "bs.sfs")).jafs
    eng = misti_b200.Engine(0)
    gid = eng.add_grid(ds["times"], ds["lambdas"])
    m1 = eng.add_model(gid, 40, 0, bands=[(1, 5, 12, 0.8, 0)])
    m2 = eng.add_model(gid, 38, 0, bands=[(0, 4, 38, 3.0, 0)])
    m3 = eng.add_model(gid, 40, 0, bands=[(0, 2, 10, 0.3, 0), (1, 5, 12, 0.8, 1)], pulses=[(0, 7, 0.05, 2)])
    rng = np.random.default_rng(3)
    res = {}
    for R in (1, 70, 333):
        eng.set_data(bs[:R], True)
        p = np.zeros((5000, 3)); p[:, 0] = rng.uniform(0, 3, 5000)
        out = eng.evaluate(p, model=m1, flags=15, want=("jafs", "status", "nfev", "terms"), row_best=True)
        mids = np.where(np.arange(777) % 3 == 0, m2, m1).astype(np.int32)
        q = np.zeros((777, 3)); q[:, 0] = rng.uniform(0, 3, 777)
        out2 = eng.evaluate(q, model_ids=mids, flags=15, want=("status", "lc"), row_best=True)
        for k, v in out.items():
            res["a%d_%s" % (R, k)] = v
        for k, v in out2.items():
            res["b%d_%s" % (R, k)] = v
        assert np.array_equal(out["row_best_llh"], np.max(np.where(np.isnan(out["llh"]), -np.inf, out["llh"]), axis=0)), R
    eng.set_data(bs[:5], True)
    x0 = np.zeros((96, 3)); x0[:, 0] = rng.uniform(0, 4, 96)
    r = eng.nelder_mead(x0[:, :1], np.where(np.arange(96) % 2 == 0, m1, m2).astype(np.int32), np.arange(96, dtype=np.int32) % 5, flags=15, maxiter=60)
    for k in ("x", "fun", "nfev", "nit"):
        res["nm_" + k] = r[k]
    x3 = np.column_stack([rng.uniform(0, 5, 40), rng.uniform(0, 5, 40), rng.uniform(0, 0.5, 40)])
    b = eng.basinhopping(x3, np.full(40, m3, dtype=np.int32), seeds=list(range(40)), flags=15, niter=2)
    for k in ("x", "fun", "nfev", "accepted"):
        res["bh_" + k] = b[k]
    eng.close()
    return res


if __name__ == "__main__":
    if len(sys.argv) > 1:
        np.savez(sys.argv[1], **run())
        sys.exit(0)
    env = dict(os.environ, MISTI_MAX_CHUNK="3000", MISTI_FIT_SLICE_US="30")
    subprocess.run([sys.executable, __file__, "/tmp/sc_small.npz"], check=True, env=env)
    subprocess.run([sys.executable, __file__, "/tmp/sc_default.npz"], check=True)
    a, b = np.load("/tmp/sc_small.npz"), np.load("/tmp/sc_default.npz")
    bad = [k for k in a.files if not np.array_equal(a[k], b[k], equal_nan=True)]
    print("arrays compared:", len(a.files), "differing:", bad)
    sys.exit(1 if bad else 0)
