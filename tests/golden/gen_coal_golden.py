#!/usr/bin/env python3
"""Golden vectors for the forward map MigrationInference.CoalescentRates (MigrationInference.py:542-564 ->
CorrectLambda.CoalRates, CorrectLambda.py:112-122): true model rates -> PSMC-apparent rates and the trajectory of the
3-state chains.  Run in the build container only (needs /root/reference through ref_shim); writes coal.json next to
this script.  The grid is the merged synthetic grid of datasets.json, the rates are those of the truth model of
make_synthetic.py (so the first case reproduces the rates the synthetic PSMC files were written from).  The reference's
method never sets the migration rates of its CorrectLambda helper: every interval is propagated with the rates the
preceding JAFSLikelihood call left there (those of interval splitT - 1) -- the fixtures pin exactly that behaviour."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
import make_synthetic as ms  # noqa: E402


def fl(x):
    return [float(v) for v in x]


def main():
    R = ref_shim.load()
    ref_shim.reset_units(R["migrationIO"])
    g = ms.psmc_grid()
    grids = [[t * s for t in g] for s in ms.THETA_SCALE]
    merged = sorted(grids[0] + grids[1][1:])
    times = [b - a for a, b in zip(merged[:-1], merged[1:])]
    cases = []
    specs = [("truth_band", 40, [[2, 5, 12, 0.8, 0]], []),
             ("no_migration", 40, [], []),
             ("two_bands_pulse", 38, [[1, 2, 10, 0.3, 0], [2, 5, 12, 0.8, 0]], [[1, 7, 0.05, 0]]),
             ("pulse_at_zero", 41, [[1, 0, 6, 1.5, 0]], [[2, 0, 0.1, 0]]),
             ("band_to_split", 36, [[1, 4, 36, 3.0, 0]], []),
             ("bands_at_split_pulse", 40, [[1, 30, 40, 0.7, 0], [2, 35, 40, 1.2, 0]], [[2, 3, 0.2, 0]])]
    for name, st, mi, pu in specs:
        lam = []
        for i, t in enumerate(merged):
            lam.append([1.0 / ms.ne1(t), 1.0 / ms.ne2(t)] if i < st else [1.0 / ms.nea(t), 1.0 / ms.nea(t)])
        M = R["MigrationInference"](list(times), [list(v) for v in lam], [1] * 8, st, [list(m) for m in mi], [list(p) for p in pu],
                                    unfolded=True, trueEPS=True)
        M.JAFSLikelihood([])  # as TestModel.py:96 does first; leaves cl.mu = mi[splitT - 1], which CoalescentRates then uses throughout
        M.CoalescentRates()
        cases.append({"name": name, "splitT": st, "mi": mi, "pu": pu, "times": fl(times), "lambdas": [fl(v) for v in lam],
                      "expect": {"lh": [fl(v) for v in M.lh], "Pr": [[fl(r) for r in P] for P in M.Pr]}})
    import numpy
    import scipy
    with open(os.path.join(HERE, "coal.json"), "w") as f:
        json.dump({"meta": {"generator": "tests/golden/gen_coal_golden.py", "numpy": numpy.__version__, "scipy": scipy.__version__},
                   "cases": cases}, f)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
