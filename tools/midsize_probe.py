#!/usr/bin/env python3
"""Plain batches of 1 024 ... 32 768 items (config 2): kernel times with the default choice of kernels and with the large-batch
kernels forced (MISTI_JSFS_PAIR=1 MISTI_DEFER_POST=2) -- where should the switch be?"""
import json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "--run":
    import misti_b200
    ds = json.load(open(os.path.join(ROOT, "tests", "golden", "datasets.json")))["datasets"]["synthetic"]
    eng = misti_b200.Engine(0)
    gid = eng.add_grid(ds["times"], ds["lambdas"])
    m1 = eng.add_model(gid, 40, 0, bands=[(1, 5, 12, 0.8, 0)])
    eng.set_data([ds["sfs"]], True)
    rng = np.random.default_rng(1)
    out = {}
    for B in [int(v) for v in os.environ.get('PROBE_SIZES', '1024,2048,4096,6144,8192,12288,16384,24576,32768').split(',')]:
        p = rng.uniform(0, 5, (B, 1))
        ts = []
        for _ in range(9):
            eng.evaluate(p, model=m1, flags=15, want=("status",))
            ts.append(eng.last_kernel_ms())
        out[B] = [round(float(np.median([a for a, _ in ts])), 4), round(float(np.median([b for _, b in ts])), 4)]
    print(json.dumps(out))
    sys.exit(0)
for env in ([{}, {"MISTI_CORRECT_BIG_BLOCKS": "0"}] if os.environ.get("PROBE_BIG") else [{}, {"MISTI_JSFS_PAIR": "1", "MISTI_DEFER_POST": "2"}]):
    r = subprocess.run([sys.executable, __file__, "--run"], env=dict(os.environ, **env), capture_output=True, text=True)
    print(env, r.stdout.strip(), r.stderr[-300:])
