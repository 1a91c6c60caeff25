#!/usr/bin/env python3
"""Generate golden vectors by running the UNMODIFIED reference (build container only).

Outputs (committed): tests/golden/tables.json, datasets.json, evals.json, fits.json.
The reference has no tests or golden files of its own, so these outputs of
the reference run here -- with the numpy/scipy versions recorded -- are the parity pin for both
the oracle restatement and the CUDA path.
"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

DATA = os.path.join(HERE, "..", "..", "data", "synthetic")
R = ref_shim.load()
mio, MI = R["migrationIO"], R["MigrationInference"]
TwoPop, OnePop = R["TwoPopulations"], R["OnePopulation"]


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        return fn(*a, **k)


def fl(x):
    return [float(v) for v in x]


def fl2(x):
    return [[float(v) for v in row] for row in x]


# ------------------------------------------------------------------ tables
def gen_tables():
    rng = np.random.default_rng(20240611)
    out = {"generator": [], "onepop": [], "pulse": [], "ancient": [], "jaf44": [], "jaf8": [], "zero_mig": []}
    tp = TwoPop(1, 1, 1, 1)
    out["jaf44"] = [[int(v) for v in tp.StateToJAF(i)] for i in range(44)]
    op = OnePop(1)
    out["jaf8"] = [[int(v) for v in op.StateToJAF(i)] for i in range(8)]
    out["stationary"] = [int(v) for v in TwoPop(1, 1, 0, 0).stationary]
    for _ in range(4):
        l1, l2, m1, m2 = [float(v) for v in rng.uniform(0.1, 3.0, 4)]
        M = np.asarray(TwoPop(l1, l2, m1, m2).SetMatrix())
        out["generator"].append({"args": [l1, l2, m1, m2], "M": fl2(M)})
    for args in ([1.3, 0.7, 0.0, 0.9], [0.4, 2.2, 1.7, 0.0]):
        M = np.asarray(TwoPop(*args).SetMatrix())
        out["generator"].append({"args": args, "M": fl2(M)})
    for lam in (1.0, 0.37):
        out["onepop"].append({"lam": lam, "M": fl2(np.asarray(OnePop(lam).SetMatrix()))})
    for src in (0, 1):
        for r in (0.05, 0.6):
            P0 = rng.uniform(0, 1, 44)
            out["pulse"].append({"rate": r, "src": src, "P0": fl(P0), "P1": fl(tp.PulseMigration(list(P0), r, src))})
    P0 = rng.uniform(0, 1, 44)
    out["ancient"].append({"P0": fl(P0), "P1": fl(TwoPop(1, 1, 0, 0).AncientSampleP0(list(P0)))})
    # zero-migration bookkeeping (SetInitialConditions / UpdateInitialConditions / UpdateIntegral)
    from scipy import linalg
    for _ in range(2):
        l1, l2 = [float(v) for v in rng.uniform(0.2, 2.0, 2)]
        T = float(rng.uniform(0.01, 0.5))
        model = TwoPop(l1, l2, 0.0, 0.0)
        M = model.SetMatrix()
        P0 = rng.uniform(0, 1, 44)
        P0 /= P0.sum()
        Pr = model.SetInitialConditions(np.array(P0))
        P1 = np.dot(linalg.expm(np.dot(M, T)), Pr)
        integ = np.dot(linalg.inv(M), [x - y for x, y in zip(np.asarray(P1).ravel(), Pr)])
        P1f = model.UpdateInitialConditions(np.asarray(P1).ravel())
        If = model.UpdateIntegral(np.asarray(integ).ravel(), T)
        out["zero_mig"].append({"l": [l1, l2], "T": T, "P0": fl(P0), "P1": fl(P1f), "integral": fl(If)})
    return out


# ------------------------------------------------------------------ datasets
def sum_rows(J):
    s = [0.0] * 8
    for r in J.jafs:
        s = [a + b for a, b in zip(s, r)]
    return s


def gen_datasets():
    ds = {}
    ref_shim.reset_units(mio)
    inp = quiet(mio.ReadPSMC, os.path.join(DATA, "m1.psmc"), os.path.join(DATA, "m2.psmc"), 0, -1)
    J = ref_shim.read_jafs(mio, os.path.join(DATA, "m.sfs"))
    B = ref_shim.read_jafs(mio, os.path.join(DATA, "bs.sfs"))
    bs_rows = [fl(r) for r in B.jafs[:6]]
    ds["synthetic"] = {"times": fl(inp.times), "lambdas": fl2(inp.lambdas), "sampleDate": int(inp.sampleDateDiscr),
                       "theta": float(inp.theta), "rho": float(inp.rho), "scaleTime": float(inp.scaleTime),
                       "sfs": fl(sum_rows(ref_shim.read_jafs(mio, os.path.join(DATA, "m.sfs")))), "bs_rows": bs_rows,
                       "psmc": ["m1.psmc", "m2.psmc"], "sdate": 0, "hetloss": None}
    ref_shim.reset_units(mio)
    mio.Units().SetHetLoss([0.05, 0.2])
    inp2 = quiet(mio.ReadPSMC, os.path.join(DATA, "m1.psmc"), os.path.join(DATA, "m2.psmc"), 3000.0, -1)
    ds["synthetic_ancient"] = {"times": fl(inp2.times), "lambdas": fl2(inp2.lambdas), "sampleDate": int(inp2.sampleDateDiscr),
                               "theta": float(inp2.theta), "rho": float(inp2.rho), "scaleTime": float(inp2.scaleTime),
                               "sfs": ds["synthetic"]["sfs"], "bs_rows": bs_rows,
                               "psmc": ["m1.psmc", "m2.psmc"], "sdate": 3000.0, "hetloss": [0.05, 0.2]}
    ref_shim.reset_units(mio)
    ms1 = ("4 100 -t 15000 -r 1920 30000000 -l -I 2 2 2 -n 1 10 -n 2 4.5 -eN 0.025 0.2 -ej 0.045 2 1 "
           "-eN 0.175 3 -eN 0.625 1.8 -eN 3 3.2 -eN 8 5.5")  # README.md:102
    ms2 = ("-n 2 3.0 -em 0.0 1 2 2.0 -em 0.05 2 1 3.0 -en 0.01 1 0.5 -en 0.02 2 0.05 -en 0.0375 1 0.5 "
           "-en 0.0375 2 0.5 -ej 1.25 2 1 -eM 1.25 0.0 -eN 1.25 1.0 -eN 2.0 5.0")  # migrationIO.py:661
    for name, ms in (("ms_readme", ms1), ("ms_two_bands", ms2)):
        i3 = quiet(mio.ReadMS, ms)
        ds[name] = {"times": fl(i3.times), "lambdas": fl2(i3.lambdas), "sampleDate": 0, "splitT": int(i3.divergenceTime),
                    "mi": [[int(m[0]), int(m[1]), int(m[2]), float(m[3]), int(m[4])] for m in i3.mi],
                    "pu": [[int(p[0]), int(p[1]), float(p[2]), int(p[3])] for p in i3.pu],
                    "sfs": [7.0] + [1.0] * 7, "ms": ms}
    return ds


# ------------------------------------------------------------------ evaluations
class Tracing(MI):
    def SolveDifEq(self, interval):
        MI.SolveDifEq(self, interval)
        self._trace.append({"interval": interval, "P0": fl(np.asarray(self.P0).ravel()),
                            "P1": fl(np.asarray(self.P1).ravel()), "integral": fl(np.asarray(self.integralP).ravel())})


def run_eval(ds, case, trace=False):
    d = ds[case["dataset"]]
    sfs = list(d["sfs"]) if case.get("bs", -1) < 0 else list(d["bs_rows"][case["bs"]])
    cls = Tracing if trace else MI
    flags = case["flags"]
    M = quiet(cls, list(d["times"]), [list(v) for v in d["lambdas"]], sfs, case["splitT"],
              [list(map(str, m)) for m in case["mi"]], [list(map(str, p)) for p in case["pu"]],
              smooth=flags["smooth"], unfolded=flags["unfolded"], trueEPS=flags["trueEPS"], cpfit=flags["cpfit"],
              sampleDate=d["sampleDate"], mixtureTH=0.0)
    M._trace = []
    t0 = time.perf_counter()
    llh = quiet(M.JAFSLikelihood, list(case["params"]))
    dt = time.perf_counter() - t0
    out = {"llh": float(llh), "ok": bool(np.isfinite(llh)), "seconds": dt, "numT": M.numT, "splitT_int": M.splitT}
    if np.isfinite(llh):
        out.update(JAFS=fl(M.JAFS), lc=fl2(M.lc), llh_const=float(M.llh_const),
                   Pr=[[fl(r) for r in blk] for blk in M.Pr], max_llh=float(M.MaximumLLHFunction()))
        if trace:
            out["trace"] = M._trace
    return out


def F(smooth=True, unfolded=True, trueEPS=False, cpfit=False):
    return {"smooth": smooth, "unfolded": unfolded, "trueEPS": trueEPS, "cpfit": cpfit}


def eval_cases(ds):
    cases = []

    def add(name, dataset, splitT, mi=(), pu=(), params=(), flags=None, stable=True, bs=-1, trace=False):
        cases.append({"name": name, "dataset": dataset, "splitT": splitT, "mi": [list(m) for m in mi],
                      "pu": [list(p) for p in pu], "params": list(params), "flags": flags or F(), "stable": stable,
                      "bs": bs, "trace": trace})
    # ms-string known answers (trueEPS forward model)
    for name in ("ms_readme", "ms_two_bands"):
        d = ds[name]
        add(name, name, d["splitT"], d["mi"], d["pu"], flags=F(smooth=False, trueEPS=True), trace=True)
    # config 1: no migration, fixed split, folded + unfolded, default + cpfit
    for st in (20, 36, 40, 44, 60):
        for uf in (False, True):
            add("c1_st%d_%s" % (st, "uf" if uf else "fo"), "synthetic", st, flags=F(unfolded=uf), trace=(st == 40 and uf))
    add("c1_st40_cpfit", "synthetic", 40, flags=F(cpfit=True))
    add("c1_st40_nosmooth", "synthetic", 40, flags=F(smooth=False))
    add("c1_st40.5_frac", "synthetic", 40.5, flags=F())
    add("c1_st40_trueEPS", "synthetic", 40, flags=F(trueEPS=True))
    # config 2: one band, objective at fixed parameters
    band = [[2, 5, 12, 0.8, 1]]
    for m in (0.0, 1e-4, 0.05, 0.3, 0.8, 2.0, 5.0, 20.0):
        add("c2_cpfit_m%g" % m, "synthetic", 40, band, params=[m], flags=F(cpfit=True), trace=(m == 0.8))
    for m in (0.0, 0.3, 0.8):
        add("c2_default_m%g" % m, "synthetic", 40, band, params=[m], flags=F(), stable=(m == 0.0))
    add("c2_default_fail_m2", "synthetic", 40, band, params=[2.0], flags=F(), stable=False)
    add("c2_negative", "synthetic", 40, band, params=[-0.1], flags=F(cpfit=True))
    for st in (36, 44):
        add("c2_cpfit_st%d" % st, "synthetic", st, band, params=[1.1], flags=F(cpfit=True))
    add("c2_cpfit_folded", "synthetic", 40, band, params=[0.6], flags=F(cpfit=True, unfolded=False))
    # config 3: two bands + pulse
    bands = [[1, 2, 10, 0.3, 1], [2, 5, 12, 0.8, 1]]
    pulse = [[1, 7, 0.05, 1]]
    for p in ([0.3, 0.8, 0.05], [1.7, 0.2, 0.31], [0.0, 3.1, 0.0], [4.2, 4.9, 0.45]):
        add("c3_cpfit_%g_%g_%g" % tuple(p), "synthetic", 40, bands, pulse, params=p, flags=F(cpfit=True), trace=(p[0] == 0.3))
    add("c3_default", "synthetic", 40, bands, pulse, params=[0.3, 0.8, 0.05], flags=F(), stable=False)
    add("c3_pulse_pop2", "synthetic", 40, bands, [[2, 9, 0.2, 1]], params=[0.5, 0.4, 0.15], flags=F(cpfit=True))
    add("c3_fixed_band", "synthetic", 40, [[1, 2, 10, 0.3, 0], [2, 5, 12, 0.8, 1]], pulse, params=[0.9, 0.1], flags=F(cpfit=True))
    add("c3_band_to_split", "synthetic", 40, [[1, 4, 40, 3, 1]], [], params=[2.5], flags=F(cpfit=True))
    add("c3_two_sided_to_split", "synthetic", 38, [[1, 4, 38, 3, 1], [2, 4, 38, 3, 1]], [], params=[1.5, 0.7], flags=F(cpfit=True))
    # config 4: ancient second genome with hetloss, unfolded, split grid
    for st in (30, 40, 50):
        add("c4_st%d" % st, "synthetic_ancient", st, flags=F(), trace=(st == 40))
    add("c4_cpfit_band", "synthetic_ancient", 40, [[2, 14, 20, 0.5, 1]], params=[0.7], flags=F(cpfit=True))
    add("c4_split_at_sampledate", "synthetic_ancient", 12, flags=F())
    # config 5: bootstrap rows
    for b in (0, 1, 5):
        add("c5_bs%d_st40" % b, "synthetic", 40, flags=F(unfolded=False), bs=b)
    add("c5_bs3_band", "synthetic", 40, [[1, 4, 40, 3, 1]], params=[1.9], flags=F(cpfit=True), bs=3)
    return cases


# ------------------------------------------------------------------ fits (Nelder-Mead through the reference)
class Recording(MI):
    def ObjectiveFunction(self, mu):
        res = -self.JAFSLikelihood(mu)
        self._calls.append([fl(mu), float(res)])
        return res


def run_fit(ds, case):
    d = ds[case["dataset"]]
    flags = case["flags"]
    M = quiet(Recording, list(d["times"]), [list(v) for v in d["lambdas"]], list(d["sfs"]), case["splitT"],
              [list(map(str, m)) for m in case["mi"]], [list(map(str, p)) for p in case["pu"]],
              smooth=flags["smooth"], unfolded=flags["unfolded"], trueEPS=flags["trueEPS"], cpfit=flags["cpfit"],
              sampleDate=d["sampleDate"], mixtureTH=0.0)
    M._calls = []
    t0 = time.perf_counter()
    sol = quiet(M.Solve, case["tol"])
    dt = time.perf_counter() - t0
    return {"x": fl(sol[0]), "llh": float(sol[1]), "calls": M._calls, "seconds": dt}


def main():
    import scipy
    meta = {"numpy": np.__version__, "scipy": scipy.__version__, "generated_by": "tests/golden/gen_golden.py",
            "reference": "Genomics-HSE/MiSTI (unmodified, via ref_shim)"}
    which = sys.argv[1:] or ["tables", "datasets", "evals", "fits"]
    if "tables" in which:
        with open(os.path.join(HERE, "tables.json"), "w") as f:
            json.dump({"meta": meta, **gen_tables()}, f)
    ds = gen_datasets()
    if "datasets" in which:
        with open(os.path.join(HERE, "datasets.json"), "w") as f:
            json.dump({"meta": meta, "datasets": ds}, f)
    if "evals" in which:
        cases = eval_cases(ds)
        for c in cases:
            c["expect"] = run_eval(ds, c, trace=c["trace"])
            print(c["name"], c["expect"]["llh"], round(c["expect"]["seconds"], 3), flush=True)
        with open(os.path.join(HERE, "evals.json"), "w") as f:
            json.dump({"meta": meta, "cases": cases}, f)
    if "fits" in which:
        fits = [
            {"name": "fit_c2_cpfit", "dataset": "synthetic", "splitT": 40, "mi": [[2, 5, 12, 0.8, 1]], "pu": [],
             "flags": F(cpfit=True), "tol": 1e-4},
            {"name": "fit_c5_band_to_split", "dataset": "synthetic", "splitT": 40, "mi": [[1, 4, 40, 3, 1]], "pu": [],
             "flags": F(cpfit=True), "tol": 1e-4},
            {"name": "fit_c3_cpfit", "dataset": "synthetic", "splitT": 40, "mi": [[1, 2, 10, 0.3, 1], [2, 5, 12, 0.8, 1]],
             "pu": [[1, 7, 0.05, 1]], "flags": F(cpfit=True), "tol": 1e-4},
        ]
        for c in fits:
            c["expect"] = run_fit(ds, c)
            print(c["name"], c["expect"]["x"], c["expect"]["llh"], len(c["expect"]["calls"]), round(c["expect"]["seconds"], 1), flush=True)
        with open(os.path.join(HERE, "fits.json"), "w") as f:
            json.dump({"meta": meta, "fits": fits}, f)


if __name__ == "__main__":
    main()
