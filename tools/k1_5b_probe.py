#!/usr/bin/env python3
"""Config-5b layout (`-mi 1 4 st 3 1 --cpfit`: 36 trust-region intervals per item): the correction kernel at 65 536 items with
blocks of two warps, one block per SM, and one block per SM with interval barriers."""
import json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "--run":
    import misti_b200
    ds = json.load(open(os.path.join(ROOT, "tests", "golden", "datasets.json")))["datasets"]["synthetic"]
    eng = misti_b200.Engine(0)
    gid = eng.add_grid(ds["times"], ds["lambdas"])
    m = eng.add_model(gid, 40, 0, bands=[(0, 4, 40, 3.0, 0)])
    eng.set_data([ds["sfs"]], True)
    rng = np.random.default_rng(2)
    out = {}
    for lo, hi in ((0.0, 1.0), (0.0, 3.0)):
        p = rng.uniform(lo, hi, (65536, 1))
        ts = []
        for _ in range(5):
            eng.evaluate(p, model=m, flags=15, want=("status",))
            ts.append(eng.last_kernel_ms())
        out["m in [%g, %g)" % (lo, hi)] = [round(float(np.median([a for a, _ in ts])), 3), round(float(np.median([b for _, b in ts])), 3)]
    print(json.dumps(out))
    sys.exit(0)
for env in ({"MISTI_CORRECT_BIG_BLOCKS": "0"}, {"MISTI_CORRECT_ALIGN": "0"}, {}):
    r = subprocess.run(["timeout", "120", sys.executable, __file__, "--run"], env=dict(os.environ, **env), capture_output=True, text=True)
    print(env, r.stdout.strip(), r.stderr[-200:])
