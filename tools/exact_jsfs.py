#!/usr/bin/env python3
"""Expected JSFS of golden cases in 50-digit arithmetic (mpmath), GIVEN the reference's corrected rates: the JSFS stage of
MigrationInference.JAFSpectrum / SolveDifEq (MigrationInference.py:467-540) restated with exact-to-50-digits expm and linear
solves.  Used to judge the run-away cases (rates ~1e7, generator condition ~1e8): there the reference's own float64 result
(expm + inv of a 44x44 matrix) is only good to ~1e-9, so "1e-9 from the reference" is not a meaningful bar for an
implementation -- "1e-9 from the exact value" is.

    python tools/exact_jsfs.py c3_band_to_split c5_bs3_band > tests/golden/stiff_exact.json      (build container, minutes)
"""
import json
import os
import sys

import mpmath as mp
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import misti_oracle as mo  # noqa: E402
from _cases import grid_of  # noqa: E402

mp.mp.dps = 50


def exact_spectrum(om):
    """om: OracleModel with .lc filled in; follows OracleModel.jaf_spectrum"""
    J = mp.matrix(7, 1)
    P0 = mp.matrix(mo.NSTATE2, 1)
    P0[2] = 1
    for it in range(om.numT):
        two = it < om.splitT
        m1, m2 = om.mi[it]
        if it == om.sampleDate:
            P0 = mp.matrix(mo.ancient_sample_reset(np.array([float(v) for v in P0])).tolist())  # sums of a few entries: exact enough
        pu = om.pu[it][0] + om.pu[it][1]
        if two and pu > 0:
            P0 = mp.matrix(mo.pulse_matrix(pu, 0 if om.pu[it][0] > 0 else 1).tolist()) * P0
        if it == om.splitT:
            C = np.zeros((mo.NSTATE1, mo.NSTATE2))
            for s in range(mo.NSTATE2):
                e = np.zeros(mo.NSTATE2)
                e[s] = 1.0
                C[:, s] = mo.collapse_pops(e)
            P0 = mp.matrix(C.tolist()) * P0
        if two:
            M = np.asarray(mo.generator_two_pop(om.lc[it][0], om.lc[it][1], m1, m2), dtype=float)
            W = mo.W44
            drop = list(mo.STATIONARY) if m1 + m2 == 0 else []
        else:
            M = np.asarray(mo.generator_one_pop(om.lc[it][0]), dtype=float)
            W = mo.W8
            drop = []
        n = M.shape[0]
        keep = [i for i in range(n) if i not in drop]
        Mk = mp.matrix(M[np.ix_(keep, keep)].tolist())
        full0 = P0
        Ps = mp.matrix([full0[i] for i in keep])
        last = it == om.numT - 1
        if not last:
            T = mp.mpf(om.times[it])
            P1 = mp.expm(Mk * T) * Ps
        else:
            T = None
            P1 = mp.matrix(len(keep), 1)
        integ = mp.lu_solve(Mk, P1 - Ps)
        if drop:
            P1f, If = mp.matrix(n, 1), mp.matrix(n, 1)
            for j, i in enumerate(keep):
                P1f[i], If[i] = P1[j], integ[j]
            # restored by conservation of the class mass (TwoPopulations.py:264-309; the oracle's _restore):
            # res[ind] = sum over the class of (full0[i] - out[i]), resp. (T full0[i] - out[i]), with out[drop] = 0
            for ind in drop:
                c = mo.STATE_CLASS[ind]
                sP, sI = mp.mpf(0), mp.mpf(0)
                for i in range(n):
                    if mo.STATE_CLASS[i] == c:
                        outP = P1f[i] if (i not in drop) else mp.mpf(0)
                        outI = If[i] if (i not in drop) else mp.mpf(0)
                        sP += full0[i] - outP
                        sI += (T * full0[i] if T is not None else mp.mpf(0)) - outI
                P1f[ind], If[ind] = sP, sI
            P1, integ = P1f, If
        P0 = P1
        Wi = np.array(W, dtype=float)
        if it < om.sampleDate:
            Wi = Wi.copy()
            Wi[2:, :] = 0
        J = J + mp.matrix(Wi[:, :len(integ)].tolist()) * integ
    tot = sum(J)
    return [J[i] / tot for i in range(7)]


def main():
    names = sys.argv[1:] or ["c3_band_to_split", "c5_bs3_band"]
    with open(os.path.join(ROOT, "tests", "golden", "datasets.json")) as f:
        ds = json.load(f)["datasets"]
    with open(os.path.join(ROOT, "tests", "golden", "evals.json")) as f:
        cases = {c["name"]: c for c in json.load(f)["cases"]}
    out = {"how": "tools/exact_jsfs.py: mpmath, %d digits, JSFS stage given the reference's corrected rates" % mp.mp.dps, "cases": []}
    for name in names:
        case = cases[name]
        times, lam, st, sd = grid_of(ds, case)
        d = ds[case["dataset"]]
        sfs = list(d["sfs"]) if case.get("bs", -1) < 0 else list(d["bs_rows"][case["bs"]])
        f = case["flags"]
        om = mo.OracleModel(times, lam, sfs, st, case["mi"], case["pu"], cpfit=f["cpfit"], smooth=f["smooth"], unfolded=f["unfolded"],
                            trueEPS=f["trueEPS"], sampleDate=sd)
        om.map_parameters(case["params"])
        om.lc = [list(v) for v in case["expect"]["lc"]]
        ex = exact_spectrum(om)
        ref = case["expect"]["JAFS"]
        err_ref = max(abs(float((mp.mpf(r) - e) / e)) for r, e in zip(ref, ex))
        exf = [float(v) for v in ex]
        llh_exact = float(om.score(exf))
        out["cases"].append({"name": name, "jafs_exact": [mp.nstr(v, 25) for v in ex], "reference_jafs_relerr_vs_exact": err_ref,
                             "llh_exact": llh_exact, "reference_llh": case["expect"]["llh"],
                             "reference_llh_relerr_vs_exact": abs(case["expect"]["llh"] - llh_exact) / abs(llh_exact),
                             "max_rate": float(np.max(np.array(case["expect"]["lc"])))})
        print(name, "reference vs exact: jafs", err_ref, "llh", out["cases"][-1]["reference_llh_relerr_vs_exact"], file=sys.stderr, flush=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
