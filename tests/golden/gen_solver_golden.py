#!/usr/bin/env python3
"""Iterate-level known answers of the coalescence-rate correction (build container only).

The reference solves one small non-linear system per interval with scipy.optimize.least_squares
(CorrectLambda.py:85, 260, 303, 305).  This script runs the UNMODIFIED reference on golden evaluation cases with that
call wrapped, and records for every call, in order: the start vector, the solution, scipy's `nfev` and `status`.
Where the correction is ulp-chaotic (default mode with migration) it also runs the reference with every
parameter moved by one ulp up and down, and four times with the entries of every 3x3 `scipy.linalg.expm` result of the
correction (CorrectLambda.py:62) moved by at most one ulp at random -- no two correct implementations of expm agree more
closely than that -- and records `stable_calls` = the number of leading calls whose (nfev, status)
are the same in all runs and whose solutions agree to 1e-9: up to there the reference determines its own iterates, and
an iterate-faithful port must reproduce the counts exactly.  Per call: `probe_agree` (the three runs took the same number
of evaluations and stopped for the same reason) and `probe_spread` (largest relative difference of the solutions) -- the
FIRST call past the stable prefix still starts from identical inputs, so where the probes agree on its counts a port
must too, and its solution can be held to a small multiple of the spread.

Output (committed): tests/golden/solver.json."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

R = ref_shim.load()
import CorrectLambda as CL  # noqa: E402  (the reference's module, on sys.path after ref_shim.load())
from gen_golden import MI, quiet  # noqa: E402

CASES = ["c1_st40_uf", "c1_st20_fo", "c1_st40_nosmooth", "c1_st40.5_frac", "c4_st40", "c2_default_m0", "c2_default_m0.3",
         "c2_default_m0.8", "c2_default_fail_m2", "c3_default", "c2_cpfit_m0.8", "c2_cpfit_m5", "c3_cpfit_0.3_0.8_0.05",
         "c3_band_to_split", "c4_cpfit_band"]


class Recorder:
    """stands in for `scipy.optimize` inside CorrectLambda.py: least_squares is wrapped, the rest passes through"""

    def __init__(self, real):
        self._real, self.calls = real, []

    def __getattr__(self, name):
        return getattr(self._real, name)

    def least_squares(self, fun, x0, **kw):
        x0v = [float(v) for v in np.atleast_1d(np.asarray(x0, dtype=float))]
        res = self._real.least_squares(fun, x0, **kw)
        lb = kw.get("bounds", (-np.inf, np.inf))[0]
        self.calls.append({"x0": x0v, "x": [float(v) for v in res.x], "nfev": int(res.nfev), "status": int(res.status),
                           "bounded": bool(np.isfinite(lb)), "lb": float(lb) if np.isfinite(lb) else None})
        return res


class NoisyLinalg:
    """stands in for `scipy.linalg` inside CorrectLambda.py: expm's result is moved by at most one ulp per entry"""

    def __init__(self, real, seed):
        self._real, self._rng = real, np.random.default_rng(seed)

    def __getattr__(self, name):
        return getattr(self._real, name)

    def expm(self, A):
        E = self._real.expm(A)
        return E * (1.0 + self._rng.integers(-1, 2, E.shape) * 2.0 ** -52)


def run(ds, case, params, expm_seed=None):
    d = ds[case["dataset"]]
    flags = case["flags"]
    sfs = list(d["sfs"]) if case.get("bs", -1) < 0 else list(d["bs_rows"][case["bs"]])
    rec = Recorder(CL.optimize._real if isinstance(CL.optimize, Recorder) else CL.optimize)
    CL.optimize = rec
    real_linalg = CL.linalg
    if expm_seed is not None:
        CL.linalg = NoisyLinalg(real_linalg, expm_seed)
    try:
        M = quiet(MI, list(d["times"]), [list(v) for v in d["lambdas"]], sfs, case["splitT"], [list(map(str, m)) for m in case["mi"]],
                  [list(map(str, p)) for p in case["pu"]], smooth=flags["smooth"], unfolded=flags["unfolded"],
                  trueEPS=flags["trueEPS"], cpfit=flags["cpfit"], sampleDate=d["sampleDate"], mixtureTH=0.0)
        llh = quiet(M.JAFSLikelihood, list(params))
    finally:
        CL.optimize = rec._real
        CL.linalg = real_linalg
    return rec.calls, float(llh)


def main():
    with open(os.path.join(HERE, "datasets.json")) as f:
        ds = json.load(f)["datasets"]
    with open(os.path.join(HERE, "evals.json")) as f:
        cases = {c["name"]: c for c in json.load(f)["cases"]}
    out = []
    for name in CASES:
        case = cases[name]
        calls, llh = run(ds, case, case["params"])
        assert llh == case["expect"]["llh"] or (np.isinf(llh) and not case["expect"]["ok"]), name
        stable = len(calls)
        llh_probe = []
        for c in calls:
            c["probe_agree"], c["probe_spread"] = True, 0.0
        has_mig = any(float(m[3]) != 0.0 or int(m[4]) == 1 for m in case["mi"]) and any(v != 0.0 for v in case["params"])
        probes = [("param", +1), ("param", -1)] + ([("expm", 1), ("expm", 2), ("expm", 3), ("expm", 4)] if has_mig else [])
        for kind, arg in probes:
            if not any(v != 0.0 for v in case["params"]):
                continue  # nothing to move by one ulp (one ulp off zero is a denormal or a negative rate)
            if kind == "param":
                p = [float(np.nextafter(v, np.inf * arg)) if v != 0.0 else 0.0 for v in case["params"]]
                other, l2 = run(ds, case, p)
            else:  # the same parameters, scipy.linalg.expm's 3x3 results moved by one ulp at random
                other, l2 = run(ds, case, case["params"], expm_seed=arg)
            llh_probe.append(l2)
            k, prefix = 0, True
            while k < min(len(calls), len(other)):
                a, b = calls[k], other[k]
                counts = a["nfev"] == b["nfev"] and a["status"] == b["status"]
                spread = float(np.max(np.abs(np.array(a["x"]) - np.array(b["x"])) / np.abs(np.array(a["x"]))))
                a["probe_agree"] = a["probe_agree"] and counts
                a["probe_spread"] = max(a["probe_spread"], spread)
                if prefix and not (counts and spread <= 1e-9):
                    stable, prefix = min(stable, k), False
                k += 1
            for c in calls[k:]:
                c["probe_agree"] = False
            if prefix:
                stable = min(stable, k)
        out.append({"name": name, "calls": calls, "stable_calls": stable, "llh": llh, "llh_one_ulp_probe": llh_probe})
        print(name, len(calls), "calls, stable", stable, "nfev total", sum(c["nfev"] for c in calls), llh, llh_probe, flush=True)
    # An ill-conditioned --cpfit chain found by the fuzz sweep (tools/fuzz_parity.py part D, profiles/r01_fuzz_parity.json):
    # the corrected rates run away to ~200 before the split and the reference's result is BISTABLE -- one trust-region solve
    # stops one evaluation earlier or later depending on the last bit of its 3x3 expm results, and the likelihood lands on
    # one of two values 1.5e-8 apart.  Recorded: the reference's likelihood as is and under eight one-ulp expm probes.
    branch_case = {"name": "fuzz_st52_two_bands", "dataset": "synthetic", "splitT": 52, "mi": [[1, 3, 10, 0.607, 1], [2, 42, 49, 0.802, 1]],
                   "pu": [], "flags": {"smooth": True, "unfolded": True, "trueEPS": False, "cpfit": True}}
    branch_points = [[3.705137947124301, 8.757866362592685], [3.7058955164970597, 8.758171298915583],
                     [3.7067050415996903, 8.758509405044837], [3.707060261654498, 8.758613320009076],
                     [3.706956953503096, 8.758538988322861], [3.706831580522753, 8.758501444193307]]
    branches = []
    for x in branch_points:
        calls, llh = run(ds, branch_case, x)
        probes = [run(ds, branch_case, x, expm_seed=k)[1] for k in range(1, 9)]
        branches.append({"x": x, "llh": llh, "llh_expm_one_ulp_probes": probes, "nfev": sum(c["nfev"] for c in calls)})
        print("branch", x, llh, sorted(set(round(v, 6) for v in probes)), flush=True)
    import scipy
    meta = {"numpy": np.__version__, "scipy": scipy.__version__, "generated_by": "tests/golden/gen_solver_golden.py",
            "reference": "Genomics-HSE/MiSTI (unmodified, via ref_shim; scipy.optimize.least_squares wrapped, not changed)"}
    with open(os.path.join(HERE, "solver.json"), "w") as f:
        json.dump({"meta": meta, "cases": out, "bistable": {"case": branch_case, "points": branches}}, f)


if __name__ == "__main__":
    main()
