#!/usr/bin/env python3
"""The run-away regime of BASELINE config 5b (`-mi 1 4 st 3 1 --cpfit`, band up to the split): for m above ~1.5 the
reference's own least-squares correction runs away to rates ~1e5..1e8 on the last band intervals, which the JSFS stage
handles with the dense scaling-and-squaring step (misti_stiff_kernel, FP64 MMA).  Prints, per range of m, the share of
such items, series lengths, solver evaluations and the device times.  Run under gpurun; `--one LO HI` evaluates a single
range once (for ncu)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import misti_b200  # noqa: E402


def main():
    with open(os.path.join(ROOT, "tests", "golden", "datasets.json")) as f:
        ds = json.load(f)["datasets"]["synthetic"]
    eng = misti_b200.Engine(0)
    gid = eng.add_grid(ds["times"], ds["lambdas"])
    mid = eng.add_model(gid, 40, 0, bands=[(0, 4, 40, 3.0, 0)])
    eng.set_data([ds["sfs"]], True)
    flags = misti_b200.FLAG_CORRECT | misti_b200.FLAG_CPFIT | misti_b200.FLAG_SMOOTH | misti_b200.FLAG_UNFOLDED
    B = 8192
    ranges = [(0.0, 0.5), (0.5, 1.0), (1.0, 2.0), (2.0, 3.0), (0.0, 3.0)]
    once = "--one" in sys.argv
    if once:
        i = sys.argv.index("--one")
        ranges = [(float(sys.argv[i + 1]), float(sys.argv[i + 2]))]
    out = {}
    for lo, hi in ranges:
        p = np.random.default_rng(1).uniform(lo, hi, (B, 1))
        r = eng.evaluate(p, model=mid, flags=flags, want=("status", "terms", "nfev", "lc"))
        if not once:
            eng.evaluate(p, model=mid, flags=flags, want=("status",))
        k1, k2 = eng.last_kernel_ms()
        mx = r["lc"][:, :40].max(axis=(1, 2))
        out["m in [%.1f, %.1f)" % (lo, hi)] = {
            "items": B, "ok": int((r["status"] == 0).sum()), "share_with_rates_above_1e4": float((mx > 1e4).mean()),
            "terms_mean": float(r["terms"].mean()), "terms_max": int(r["terms"].max()), "nfev_mean": float(r["nfev"].mean()),
            "nfev_max": int(r["nfev"].max()), "correct_kernel_ms": k1, "jsfs_and_stiff_kernels_ms": k2}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
