#!/usr/bin/env python3
"""A/B of one tuning knob at the bench batch: results must be bit-identical, kernel times are printed.
    python tools/knob_ab.py MISTI_SPLIT_SEGMENTS 0 1"""
import json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(out):
    import misti_b200
    ds = json.load(open(os.path.join(ROOT, "tests", "golden", "datasets.json")))["datasets"]["synthetic"]
    eng = misti_b200.Engine(0)
    gid = eng.add_grid(ds["times"], ds["lambdas"])
    m1 = eng.add_model(gid, 40, 0, bands=[(1, 5, 12, 0.8, 0)])
    m3 = eng.add_model(gid, 40, 0, bands=[(0, 2, 10, 0.3, 0), (1, 5, 12, 0.8, 1)], pulses=[(0, 7, 0.05, 2)])
    eng.set_data([ds["sfs"]], True)
    rng = np.random.default_rng(1234)
    res, ms = {}, {}
    for name, mid, cols, flags in (("c2", m1, 1, 15), ("c3", m3, 3, 15), ("c2_default", m1, 1, 13)):
        p = rng.uniform(0, 5, (65536, cols))
        if cols == 3:
            p[:, 2] *= 0.1
        t = []
        for _ in range(8):
            o = eng.evaluate(p, model=mid, flags=flags, want=("jafs", "status", "terms", "nfev"))
            t.append(eng.last_kernel_ms())
        ms[name] = [float(np.median([a for a, _ in t])), float(np.median([b for _, b in t]))]
        for k, v in o.items():
            res[name + "_" + k] = v
    eng.close()
    np.savez(out, **res)
    json.dump(ms, open(out + ".json", "w"))


if __name__ == "__main__":
    if sys.argv[1] == "--run":
        run(sys.argv[2])
        sys.exit(0)
    knob, vals = sys.argv[1], sys.argv[2:]
    outs = []
    for v in vals:
        f = "/tmp/knob_%s.npz" % v
        subprocess.run([sys.executable, __file__, "--run", f], check=True, env=dict(os.environ, **{knob: v}))
        outs.append((np.load(f), json.load(open(f + ".json"))))
    a = outs[0][0]
    for v, (b, ms) in zip(vals, outs):
        bad = {}
        for k in a.files:
            if not np.array_equal(a[k], b[k], equal_nan=True):
                x, y = a[k].astype(float), b[k].astype(float)
                m = np.isfinite(x) & np.isfinite(y)
                bad[k] = {"max_rel": float(np.max(np.abs(x[m] - y[m]) / np.maximum(np.abs(x[m]), 1e-300))) if m.any() else None,
                          "nonfinite_pattern_equal": bool(np.array_equal(np.isfinite(x), np.isfinite(y)))}
        print(knob, "=", v, "kernel ms (K1, K2):", ms, "differs from the first in:", bad)
