#!/usr/bin/env python3
"""MISTI_CORRECT_ALIGN (a barrier at every interval of the correction chain in one-block-per-SM launches) on batches that
exercise the make-up arrivals: negative parameters, corrections that fail, models of different length in one block, batches
that are not a multiple of the block.  Results must be bit-identical with and without; prints the kernel times."""
import json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "--run":
    import misti_b200
    ds = json.load(open(os.path.join(ROOT, "tests", "golden", "datasets.json")))["datasets"]["synthetic"]
    eng = misti_b200.Engine(0)
    gid = eng.add_grid(ds["times"], ds["lambdas"])
    numT = len(ds["lambdas"])
    ms = [eng.add_model(gid, 40, 0, bands=[(1, 5, 12, 0.8, 0)]), eng.add_model(gid, 20, 0, bands=[(1, 5, 12, 0.8, 0)]),
          eng.add_model(gid, 90, 0, bands=[(0, 4, 38, 3.0, 0)]), eng.add_model(gid, numT, 0, bands=[(0, 100, numT, 0.5, 0)]),
          eng.add_model(gid, 36, 0)]
    eng.set_data([ds["sfs"]], True)
    rng = np.random.default_rng(9)
    res, t = {}, {}
    for name, B, flags, mixed in (("c2", 65536, 15, False), ("c2_ragged", 60001, 15, False), ("mixed", 70001, 15, True),
                                  ("mixed_default", 65536, 13, True), ("big", 200000, 15, True)):
        p = rng.uniform(0, 5, (B, 1))
        p[::53, 0] = -1.0
        kw = dict(model_ids=np.array(ms, dtype=np.int32)[rng.integers(0, len(ms), B)]) if mixed else dict(model=ms[0])
        ts = []
        for _ in range(4):
            o = eng.evaluate(p, flags=flags, want=("jafs", "status", "nfev", "terms"), **kw)
            ts.append(eng.last_kernel_ms()[0])
        t[name] = float(np.median(ts))
        for k, v in o.items():
            res[name + "_" + k] = v
    eng.close()
    np.savez(sys.argv[2], **res)
    json.dump(t, open(sys.argv[2] + ".json", "w"))
    sys.exit(0)
outs = []
for v in ("0", "1"):
    f = "/tmp/align_%s.npz" % v
    r = subprocess.run(["timeout", "100", sys.executable, __file__, "--run", f], env=dict(os.environ, MISTI_CORRECT_ALIGN=v))
    if r.returncode != 0:
        print("MISTI_CORRECT_ALIGN =", v, "FAILED, rc", r.returncode)
        sys.exit(1)
    outs.append((np.load(f), json.load(open(f + ".json"))))
a, b = outs[0][0], outs[1][0]
bad = [k for k in a.files if not np.array_equal(a[k], b[k], equal_nan=True)]
print("K1 ms without:", outs[0][1]); print("K1 ms with:   ", outs[1][1]); print("arrays", len(a.files), "differing:", bad)
sys.exit(1 if bad else 0)
