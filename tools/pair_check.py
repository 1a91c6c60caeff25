#!/usr/bin/env python3
"""The pair-of-lanes JSFS kernel (MISTI_JSFS_PAIR = 1) against the 16-lane kernel (= 0) on the same items: every layout the
two share (bands, pulses, sampling date, default and cpfit mode, folded and unfolded, mixed models incl. stiff items and
an infinite last interval, few and many data rows).  Status and term counts must be equal, spectra and likelihoods agree to
rounding (the two sum in different orders).  Prints the worst relative deviation per case and the kernel times."""
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
    numT = len(ds["lambdas"])
    m1 = eng.add_model(gid, 40, 0, bands=[(1, 5, 12, 0.8, 0)])
    m2 = eng.add_model(gid, 38, 0, bands=[(0, 4, 38, 3.0, 0)])
    m3 = eng.add_model(gid, 40, 0, bands=[(0, 2, 10, 0.3, 0), (1, 5, 12, 0.8, 1)], pulses=[(0, 7, 0.05, 2)])
    m4 = eng.add_model(gid, 44, 12, bands=[(1, 5, 20, 0.8, 0)], pulses=[(1, 12, 0.1, -1)])
    m5 = eng.add_model(gid, 30, 30, bands=[(0, 3, 9, 0.5, 0)])
    m6 = eng.add_model(gid, numT, 0, bands=[(0, 100, numT, 0.5, 0), (1, 100, numT, 0.7, -1)])
    m7 = eng.add_model(gid, 36, 0)
    rng = np.random.default_rng(11)
    res, ms = {}, {}

    def case(name, n, model=None, mids=None, flags=15, unfolded=True, R=1, cols=3, hi=3.0):
        eng.set_data(bs[:R], unfolded)
        p = np.zeros((n, cols))
        p[:, 0] = rng.uniform(0, hi, n)
        if cols > 1:
            p[:, 1] = rng.uniform(0, hi, n)
            p[:, 2] = rng.uniform(0, 0.5, n)
        p[::97, 0] = -0.1  # negative parameters
        kw = dict(model=model) if mids is None else dict(model_ids=mids)
        for rep in range(3):
            out = eng.evaluate(p, flags=flags, want=("jafs", "status", "terms"), **kw)
        ms[name] = eng.last_kernel_ms()
        for k, v in out.items():
            res[name + "_" + k] = v

    case("c2", 65536, model=m1, cols=1)
    case("c2_rows5", 20000, model=m1, cols=1, R=5)
    case("c2_rows70", 20000, model=m1, cols=1, R=70)
    case("c2_folded", 20000, model=m1, cols=1, unfolded=False, flags=7)
    case("c2_default", 20000, model=m1, cols=1, flags=13)
    case("c3", 30000, model=m3)
    case("c4_sdate", 20000, model=m4, cols=1)
    case("c4_split_at_sdate", 20000, model=m5, cols=1)
    case("plain", 17000, model=m7, cols=1)
    mids = np.array([m1, m2, m3, m4, m5, m6, m7], dtype=np.int32)[np.arange(40000) % 7]
    case("mixed", 40000, mids=mids)
    case("mixed_default", 20000, mids=mids[:20000], flags=13)
    # a split-time grid, interleaved item by item: the segment lists agree in type, so the warps stay in the pair kernel
    grid = np.array([eng.add_model(gid, st, 0, bands=[(1, 5, 12, 0.8, 0)]) for st in range(36, 45)], dtype=np.int32)
    case("split_grid", 45000, mids=grid[np.arange(45000) % 9])
    eng.close()
    return res, ms


if __name__ == "__main__":
    if len(sys.argv) > 1:
        res, ms = run()
        np.savez(sys.argv[1], **res)
        json.dump(ms, open(sys.argv[1] + ".json", "w"))
        sys.exit(0)
    for v in ("0", "1"):
        subprocess.run([sys.executable, __file__, "/tmp/pc_%s.npz" % v], check=True, env=dict(os.environ, MISTI_JSFS_PAIR=v))
    a, b = np.load("/tmp/pc_0.npz"), np.load("/tmp/pc_1.npz")
    ma, mb = json.load(open("/tmp/pc_0.npz.json")), json.load(open("/tmp/pc_1.npz.json"))
    ok = True
    rep = {}
    for k in a.files:
        x, y = a[k], b[k]
        if k.endswith("_status") or k.endswith("_terms"):
            same = np.array_equal(x, y)
            rep[k] = {"equal": bool(same), "nonzero": int((x != 0).sum()) if k.endswith("_status") else None}
            if not same:
                bad = np.nonzero(x != y)[0]
                rep[k]["first_bad"] = [int(bad[0]), int(x[bad[0]]), int(y[bad[0]]), int(len(bad))]
            ok &= same
        else:
            nan_same = np.array_equal(np.isnan(x), np.isnan(y))
            inf_same = np.array_equal(np.isinf(x), np.isinf(y))
            m = np.isfinite(x) & np.isfinite(y)
            d = float(np.max(np.abs(x[m] - y[m]) / np.abs(x[m]))) if m.any() else 0.0
            rep[k] = {"worst_rel": d, "nan_same": bool(nan_same), "inf_same": bool(inf_same), "finite": int(m.sum())}
            ok &= nan_same and inf_same and d < 1e-10  # (llh: the rounding of the logs times counts of 1e6)
    print(json.dumps({"ok": bool(ok), "cases": rep, "kernel_ms_16lane": ma, "kernel_ms_pair": mb}, indent=1))
    sys.exit(0 if ok else 1)
