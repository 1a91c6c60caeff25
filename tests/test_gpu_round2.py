"""GPU parity, second batch: the BASELINE configurations at FULL size (1 001 bootstrap rows x 9 split times, every
split time of the ancient-sample grid, the config-5b fit layout of the golden fits), the optimisers against scipy itself
around the CPU oracle (basin-hopping with a seed), and the one fuzz finding outside 1e-9 pinned on what the reference
itself does there.  Everything goes through the C ABI (misti_b200.Engine)."""
import json
import os

import numpy as np
import pytest

from _cases import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-9
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "data", "synthetic")
CPFIT_UF = 1 | 2 | 4 | 8  # correct, cpfit, smooth, unfolded


def _inputs():
    from misti_b200 import io as mio
    units = mio.Units.from_file(os.path.join(DATA, "setunits.txt"))
    inp = mio.read_psmc(os.path.join(DATA, "m1.psmc"), os.path.join(DATA, "m2.psmc"), 0, -1, units)
    bs = mio.read_jafs(os.path.join(DATA, "bs.sfs")).jafs
    return inp, np.asarray(bs, dtype=np.float64)


def test_config5b_fit_layout_matches_the_references_fit(engine, golden_datasets, golden_fits):
    """golden fit `fit_c5_band_to_split` (MiSTI.py ... 40 -uf -mi 1 4 40 3 1 --cpfit: the layout of BASELINE config 5b,
    test.bs/din_sar.bs.sh:29-38), fitted by the reference's own scipy Nelder-Mead: the on-device optimiser takes the same
    number of evaluations and ends on the same rate and likelihood -- although the fit starts in the run-away regime of the
    correction (m = 3: rates ~1e7, dense scaling-and-squaring step)."""
    from misti_b200.sweep import Sweep
    fit = [f for f in golden_fits if f["name"] == "fit_c5_band_to_split"][0]
    ds = golden_datasets[fit["dataset"]]
    sw = Sweep(ds["times"], ds["lambdas"], [ds["sfs"]], unfolded=True, cpfit=True, smooth=True, engine=engine)
    sw.add_model(fit["splitT"], fit["mi"], fit["pu"])
    exp = fit["expect"]
    for on_device in (True, False):
        res = sw.solve(tol=fit["tol"], on_device=on_device)
        assert res["success"][0]
        assert res["nfev"][0] == len(exp["calls"]), (on_device, res["nfev"][0], len(exp["calls"]))
        assert np.allclose(res["x"][0][:1], exp["x"], rtol=1e-6, atol=1e-9), (on_device, res["x"][0], exp["x"])
        assert relerr(res["llh"][0], exp["llh"]) < TOL, on_device
    # every point the reference's optimiser visited: same likelihood (1e-9 outside the run-away regime; there the
    # reference's own value moves by more than that under a one-ulp probe, tests/golden/solver.json: c3_band_to_split)
    xs = np.array([c[0] for c in exp["calls"]])
    out = engine.evaluate(xs, model=sw.models[0]["id"], flags=sw.flags, want=("status", "lc"))
    for k, (x, f) in enumerate(exp["calls"]):
        if not np.isfinite(f):
            assert out["status"][k] != 0
            continue
        runaway = float(np.nanmax(out["lc"][k])) > 1e3
        assert relerr(out["llh"][k, 0], -f) < (1e-6 if runaway else TOL), (x, runaway)


def test_full_bootstrap_sweep_1001_rows_times_9_split_times(engine):
    """BASELINE config 5 at full size: data/synthetic/bs.sfs (row 0 + 1 000 bootstrap rows, utils/generateJSFS_bs.py:39-48)
    x split times 36..44 (test.bs/san_sar.bs.no.mig.sh:29-36).  (a) no migration: ONE launch scores the 9 models against
    the 1 001 rows; all 9 009 likelihoods are checked against the spectrum the device returns (llh = const_r + sum_i d_ri log
    p_i in the reference's summation order, MigrationInference.py:583-613) and a sample of rows against the CPU oracle;
    the per-row arg-max over split times (test.bs/bs_conf_int.ipynb) comes back from the device.  (b) `-mi 1 4 st 3 1
    --cpfit`: 9 009 Nelder-Mead fits on the device; results do not depend on how the fits are packed into rounds (the
    whole sweep against the same fits taken alone), sampled fits against scipy around the oracle."""
    from misti_b200.engine import llh_constants
    from misti_b200.sweep import Sweep
    from oracle.misti_oracle import OracleModel
    inp, bs = _inputs()
    assert bs.shape == (1001, 8)
    sts = list(range(36, 45))
    # (a) folded, default mode, as the no-migration script runs it
    sw = Sweep(inp.times, inp.lambdas, bs, unfolded=False, cpfit=False, smooth=True, engine=engine)
    for st in sts:
        sw.add_model(st)
    llh = sw.evaluate_grid()
    assert llh.shape == (9, 1001) and np.isfinite(llh).all()
    mids = np.array([m["id"] for m in sw.models], dtype=np.int32)
    out = engine.evaluate(np.zeros((9, 0)), model_ids=mids, flags=sw.flags, want=("jafs", "status"))
    jn = out["jafs"]
    logs = np.stack([np.log(jn[:, 0] + jn[:, 6]), np.log(jn[:, 1] + jn[:, 5]), np.log(jn[:, 2] + jn[:, 4]), np.log(jn[:, 3])], axis=1)
    d = bs[:, 1:]
    folded = np.stack([d[:, 0] + d[:, 6], d[:, 1] + d[:, 5], d[:, 2] + d[:, 4], d[:, 3]], axis=1)
    want = llh_constants(bs, False)[None, :].copy()
    for c in range(4):  # the reference's order of summation
        want = want + folded[None, :, c] * logs[:, None, c]
    assert relerr(llh, want) < 1e-12  # the logs are taken on the device (one ulp) and the products are fused
    for r in (0, 1, 500, 1000):
        for k in (0, 4, 8):
            om = OracleModel(inp.times, inp.lambdas, list(bs[r]), sts[k], smooth=True, unfolded=False)
            assert relerr(llh[k, r], om.likelihood([])) < TOL, (r, sts[k])
    best = sw.argmax_split()
    assert np.array_equal(best["model"], np.argmax(llh, axis=0)) and np.array_equal(best["llh"], llh.max(axis=0))
    assert set(best["splitT"].tolist()) <= set(sts)
    ci = sw.split_time_confidence()  # the notebook's confidence interval over the 1 000 replicates, from the device's reduction
    assert ci["from_data"] == best["splitT"][0] and sum(ci["histogram"].values()) == 1000
    assert ci["interval"][0] <= np.mean(best["splitT"][1:]) <= ci["interval"][1]
    # (b) one migration band up to the split, optimised: 9 009 fits
    sw = Sweep(inp.times, inp.lambdas, bs, unfolded=True, cpfit=True, smooth=True, engine=engine)
    for st in sts:
        sw.add_model(st, [[1, 4, st, 3, 1]])
    import time
    t0 = time.perf_counter()
    res = sw.solve(tol=1e-4)
    print("9 009 fits on the device: %.3f s, %d rounds, %d points" % (time.perf_counter() - t0, res["launches"], res["evaluations"]))
    assert len(res["llh"]) == 9009 and res["success"].all() and np.isfinite(res["llh"]).all()
    sample = [0, 4, 9 * 500 + 2, 9 * 1000 + 8, 9 * 77 + 5]
    alone = sw.solve(pairs=[(int(res["model"][k]), int(res["row"][k])) for k in sample], tol=1e-4)
    for j, k in enumerate(sample):
        for key in ("x", "llh", "nfev", "nit"):
            assert np.array_equal(np.asarray(alone[key][j]), np.asarray(res[key][k])), (key, k)
    ci = sw.split_time_confidence(res)  # ... and from the fitted likelihoods of the 9 009 fits
    fitted_best = np.array([sts[int(np.argmax(res["llh"][9 * r:9 * r + 9]))] for r in range(1001)])
    assert ci["from_data"] == fitted_best[0] and ci["histogram"] == {float(k): int(v) for k, v in zip(*np.unique(fitted_best[1:], return_counts=True))}
    from scipy import optimize
    # two fits with an interior optimum against scipy around the oracle (where the rate is fitted to zero the simplex walks
    # down to m ~ 1e-8, where the REFERENCE's own likelihood is inaccurate -- inv(M) of a nearly singular generator, error ~
    # 1e-16 / m, test_tiny_migration_rates_are_continuous -- and its last decisions are not determined to the last bit)
    interior = [k for k in range(len(res["llh"])) if res["x"][k][0] > 0.05][:2]
    assert len(interior) == 2
    for k in interior:
        m, r = int(res["model"][k]), int(res["row"][k])
        om = OracleModel(inp.times, inp.lambdas, list(bs[r]), sts[m], [[1, 4, sts[m], 3, 1]], [], cpfit=True, smooth=True, unfolded=True)
        ref = optimize.minimize(lambda x: -om.likelihood(list(x)), [3.0], method="Nelder-Mead",
                                options={"xatol": 1e-4, "fatol": 1e-4, "maxiter": 1000})
        assert ref.nfev == res["nfev"][k] and ref.nit == res["nit"][k], (k, ref.nfev, res["nfev"][k])
        assert abs(ref.x[0] - res["x"][k][0]) <= 1e-6 * max(abs(ref.x[0]), 1e-3) and relerr(res["llh"][k], -ref.fun) < TOL


def test_config4_every_split_time_of_the_ancient_sample_grid(engine, golden_datasets):
    """BASELINE config 4: `st -uf --sdate 3000 --hetloss 0.05 0.2` for EVERY st in [30, 60] (numT = 128, sampling date at
    interval 12): one batch with one model per split time against the CPU oracle, default and --cpfit mode."""
    from oracle.misti_oracle import OracleModel
    ds = golden_datasets["synthetic_ancient"]
    sd = int(ds["sampleDate"])
    assert sd == 12 and len(ds["lambdas"]) == 128
    sts = list(range(30, 61))
    for cpfit in (False, True):
        engine.clear_models()
        gid = engine.add_grid(ds["times"], ds["lambdas"])
        engine.set_data([ds["sfs"]], True)
        mids = np.array([engine.add_model(gid, st, sd) for st in sts], dtype=np.int32)
        out = engine.evaluate(np.zeros((len(sts), 0)), model_ids=mids, flags=1 | 4 | 8 | (2 if cpfit else 0), want=("jafs", "status"))
        assert (out["status"] == 0).all()
        for k, st in enumerate(sts):
            om = OracleModel(ds["times"], ds["lambdas"], ds["sfs"], st, cpfit=cpfit, smooth=True, unfolded=True, sampleDate=sd)
            ref = om.likelihood([])
            assert relerr(out["llh"][k, 0], ref) < TOL and relerr(out["jafs"][k], om.JAFS) < TOL, (cpfit, st)
    # with a band that starts at the sampling date (all bands must start there or later, MigrationInference.py:229-289)
    engine.clear_models()
    gid = engine.add_grid(ds["times"], ds["lambdas"])
    engine.set_data([ds["sfs"]], True)
    mids = np.array([engine.add_model(gid, st, sd, bands=[(1, sd, sd + 8, 0.5, 0)]) for st in sts], dtype=np.int32)
    par = np.full((len(sts), 1), 0.7)
    out = engine.evaluate(par, model_ids=mids, flags=CPFIT_UF, want=("jafs", "status"))
    for k in (0, 10, 30):
        om = OracleModel(ds["times"], ds["lambdas"], ds["sfs"], sts[k], [[2, sd, sd + 8, 0.5, 1]], [], cpfit=True, smooth=True,
                         unfolded=True, sampleDate=sd)
        assert relerr(out["llh"][k, 0], om.likelihood([0.7])) < TOL, sts[k]


def test_basinhopping_on_the_device_equals_scipy_around_the_oracle(engine, golden_datasets):
    """MigrationInference.Solve(globalOpt=True) (MigrationInference.py:724: scipy.optimize.basinhopping, T = 0.5, local
    search Nelder-Mead) -- the reference passes no seed; with one (`rng=seed`), scipy around the CPU oracle and a walker of
    the on-device optimiser with the same seed visit the same minima: same best rate and likelihood, same evaluation
    count, same number of accepted hops.  Walkers of one call do not influence each other (a walker alone = the walker in
    a crowd of walkers with other seeds and start points)."""
    from scipy import optimize
    from oracle.misti_oracle import OracleModel
    ds = golden_datasets["synthetic"]
    mi = [[2, 5, 12, 0.8, 1]]
    engine.clear_models()
    gid = engine.add_grid(ds["times"], ds["lambdas"])
    mid = engine.add_model(gid, 40, 0, bands=[(1, 5, 12, 0.8, 0)])
    engine.set_data([ds["sfs"]], True)
    om = OracleModel(ds["times"], ds["lambdas"], ds["sfs"], 40, mi, [], cpfit=True, smooth=True, unfolded=True)
    niter, seeds, x0 = 4, [2024, 7], [[0.8], [1.9]]
    refs = [optimize.basinhopping(lambda x: -om.likelihood(list(x)), x0[w], niter=niter, T=0.5, stepsize=0.5, rng=seeds[w],
                                  minimizer_kwargs={"method": "Nelder-Mead"}) for w in range(2)]
    # the two walkers in a crowd: six more walkers with other seeds and starts run in the same call
    rng = np.random.default_rng(1)
    X0 = np.vstack([np.array(x0), rng.uniform(0.1, 4.0, (6, 1))])
    got = engine.basinhopping(X0, np.full(8, mid, dtype=np.int32), seeds=seeds + [100 + k for k in range(6)], flags=CPFIT_UF,
                              niter=niter, T=0.5, stepsize=0.5)
    assert (got["nit"] == niter).all() and got["launches"] > 0
    for w, ref in enumerate(refs):
        assert got["nfev"][w] == ref.nfev, (w, got["nfev"][w], ref.nfev)
        assert got["minimization_failures"][w] == ref.minimization_failures
        assert abs(got["x"][w, 0] - ref.x[0]) <= 1e-6 * abs(ref.x[0]) and relerr(got["fun"][w], ref.fun) < TOL, w
    alone = engine.basinhopping(X0[:1], [mid], seeds=seeds[:1], flags=CPFIT_UF, niter=niter, T=0.5, stepsize=0.5)
    for k in ("x", "fun", "nfev", "accepted"):
        assert np.array_equal(alone[k][0], got[k][0]), k
    # the drop-in class: SolveBatch(globalOpt=True) = Solve(globalOpt=True) for every data row, one walker per row
    from misti_b200 import MigrationInference
    M = MigrationInference(list(ds["times"]), [list(v) for v in ds["lambdas"]], list(ds["sfs"]), 40, [list(map(str, m)) for m in mi], [],
                           smooth=True, unfolded=True, cpfit=True, sampleDate=0)
    M.SetJAFSBatch([list(ds["sfs"])])
    xb, llhb, info = M.SolveBatch(globalOpt=True, niter=niter, seeds=[seeds[0]])
    assert np.array_equal(xb[0], got["x"][0]) and llhb[0] == -got["fun"][0] and info["nfev"][0] == refs[0].nfev
    # the host-driven lock-step walkers (misti_b200.optim.basinhopping_batch) take the same path
    from misti_b200.optim import basinhopping_batch

    def fun(X, who):
        return -engine.evaluate(X, model=mid, flags=CPFIT_UF, want=("status",))["llh"][:, 0]
    host = basinhopping_batch(fun, X0[:2], niter=niter, T=0.5, stepsize=0.5, seeds=seeds)
    for k in ("x", "fun", "nfev", "accepted"):
        assert np.array_equal(host[k], got[k][:2]), k


def test_bistable_chain_lands_on_one_of_the_references_branches(engine, golden_datasets):
    """The one end-to-end --cpfit finding of the fuzz sweep outside 1e-9 (tools/fuzz_parity.py part D: split 52, bands
    `-mi 1 3 10 .. 1 -mi 2 42 49 .. 1`, six points of one fit, 1.5e-8 from the oracle): the corrected rates run away to ~200
    before the split, and the REFERENCE is bistable there -- moving the entries of its 3x3 expm results by one ulp flips one
    trust-region solve to one evaluation more or less and the likelihood to a second value 1.5e-8 away (recorded from the
    unmodified reference, tests/golden/solver.json `bistable`).  Derived bound: the device must be within 1e-9 of ONE of the
    values the reference itself takes under that probe."""
    with open(os.path.join(ROOT, "tests", "golden", "solver.json")) as f:
        gold = json.load(f)["bistable"]
    case = gold["case"]
    ds = golden_datasets[case["dataset"]]
    engine.clear_models()
    gid = engine.add_grid(ds["times"], ds["lambdas"])
    mid = engine.add_model(gid, case["splitT"], 0, bands=[(m[0] - 1, m[1], m[2], m[3], k) for k, m in enumerate(case["mi"])])
    engine.set_data([ds["sfs"]], True)
    xs = np.array([p["x"] for p in gold["points"]])
    out = engine.evaluate(xs, model=mid, flags=CPFIT_UF, want=("status", "nfev"))
    for k, p in enumerate(gold["points"]):
        assert out["status"][k] == 0
        values = [p["llh"]] + p["llh_expm_one_ulp_probes"]
        spread = (max(values) - min(values)) / abs(p["llh"])
        assert spread > 5e-9  # the reference IS bistable at this point
        dist = min(relerr(out["llh"][k, 0], v) for v in values)
        print("bistable point", p["x"], "device", out["llh"][k, 0], "reference", p["llh"], "nearest reference branch at", dist,
              "nfev device / reference", out["nfev"][k], p["nfev"])
        assert dist < TOL, (p["x"], out["llh"][k, 0], values)


def test_walkers_finish_their_hops_and_do_not_depend_on_the_time_slice(golden_datasets):
    """300 walkers x 8 hops on the device: every walker takes all its hops (the host must not take a round in which the
    last running walkers are between two local searches for the end of the fit), and the results do not depend on where
    -- or whether -- the correction chains of a round were interrupted (ChainCkpt): time slices of 50 us, 350 us and none."""
    import misti_b200
    ds = golden_datasets["synthetic"]
    W, niter = 300, 8
    rng = np.random.default_rng(77)
    x0 = np.column_stack([rng.uniform(0, 5, W), rng.uniform(0, 5, W), rng.uniform(0, 0.5, W)])
    res = {}
    for us in ("0", "50", "350"):
        os.environ["MISTI_FIT_SLICE_US"] = us
        try:
            eng = misti_b200.Engine(0)
        finally:
            del os.environ["MISTI_FIT_SLICE_US"]
        gid = eng.add_grid(ds["times"], ds["lambdas"])
        mid = eng.add_model(gid, 40, 0, bands=[(0, 2, 10, 0.3, 0), (1, 5, 12, 0.8, 1)], pulses=[(0, 7, 0.05, 2)])
        eng.set_data([ds["sfs"]], True)
        res[us] = eng.basinhopping(x0, np.full(W, mid, dtype=np.int32), seeds=[500 + w for w in range(W)], flags=CPFIT_UF, niter=niter)
        eng.close()
        assert (res[us]["nit"] == niter).all(), us
    assert res["50"]["launches"] > res["0"]["launches"]  # chains WERE interrupted
    for us in ("50", "350"):
        for k in ("x", "fun", "nfev", "accepted", "minimization_failures"):
            assert np.array_equal(res[us][k], res["0"][k], equal_nan=True), (us, k)


def test_tiny_migration_rates_against_50_digit_values(engine, golden_datasets):
    """The device against 50-digit values of the JSFS stage at m = 1e-14 ... 1e-4 (tests/golden/tiny_rate_exact.json): within
    1e-12 of the exact value throughout, while the reference's own float64 result is off by ~5e-17 / m (5e-3 at m = 1e-14) --
    where a fit walks to m -> 0 the two optimisers see different objectives for a reason on the reference's side."""
    with open(os.path.join(ROOT, "tests", "golden", "tiny_rate_exact.json")) as f:
        gold = json.load(f)
    ds = golden_datasets["synthetic"]
    engine.clear_models()
    gid = engine.add_grid(ds["times"], ds["lambdas"])
    mid = engine.add_model(gid, 40, 0, bands=[(1, 5, 12, 0.8, 0)])
    engine.set_data([ds["sfs"]], True)
    pts = gold["points"]
    inj = np.zeros((len(pts), engine.numT_max, 2))
    for k, pt in enumerate(pts):
        inj[k, :len(pt["lc"])] = np.array(pt["lc"])
    out = engine.evaluate(np.array([[pt["m"]] for pt in pts]), model=mid, flags=CPFIT_UF, lc_inject=inj, want=("jafs", "status"))
    for k, pt in enumerate(pts):
        assert out["status"][k] == 0
        exact = [float(v) for v in pt["jafs_exact"]]
        print("m", pt["m"], "device vs exact", relerr(out["jafs"][k], exact), "reference vs exact", pt["reference_jafs_relerr_vs_exact"])
        assert relerr(out["jafs"][k], exact) < 1e-12 and relerr(out["llh"][k, 0], pt["llh_exact"]) < 1e-11, pt["m"]


def test_small_launches_and_short_time_slices_change_nothing():
    """tools/sanitize_case.py: every kernel of the library on a small case, once with launches of at most 3 000 items and
    time slices of 30 us (several chunks per call, the per-row reduction merged across chunks, chains interrupted many
    times, the look-ahead region squeezed) and once with the defaults -- all 44 result arrays bit for bit the same."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_case.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "differing: []" in r.stdout


def test_sharded_evaluator_zero_copy_equals_staged_copies(golden_datasets):
    """misti_b200.parallel.ShardedEvaluator (the host-buffer call bench.py times end to end): with pinned buffers the kernels
    read the parameters and write likelihoods, spectra and status across PCIe themselves (MISTI_FLAG_DEVICE_PTRS with
    device-accessible host memory); the results are those of the staged-copy path and of Engine.evaluate, bit for bit."""
    import torch
    import misti_b200
    from misti_b200.parallel import ShardedEvaluator
    ds = golden_datasets["synthetic"]
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(dev)
    with torch.cuda.stream(stream):
        eng = misti_b200.Engine(0, stream=stream.cuda_stream)
        gid = eng.add_grid(ds["times"], ds["lambdas"])
        mid = eng.add_model(gid, 40, 0, bands=[(1, 5, 12, 0.8, 0)])
        eng.set_data([ds["sfs"], ds["bs_rows"][1]], True)
        B = 3001
        p = torch.from_numpy(np.random.default_rng(5).uniform(-0.1, 5.0, (B, 1))).pin_memory()
        ref = eng.evaluate(p.numpy(), model=mid, flags=CPFIT_UF, want=("jafs", "status"))
        outs = []
        for zc in (True, False):
            sh = ShardedEvaluator(eng, dev, B, 1, want_jafs=True, zero_copy=zc)
            llh, status, jafs = sh.evaluate(p, mid, CPFIT_UF)
            outs.append((llh.numpy().copy(), status.numpy().copy(), jafs.numpy().copy()))
        eng.close()
    for llh, status, jafs in outs:
        assert np.array_equal(llh, ref["llh"], equal_nan=True) and np.array_equal(status, ref["status"])
        ok = status == 0
        assert ok.sum() > B // 2 and (~ok).sum() > 0
        assert np.array_equal(jafs[ok], ref["jafs"][ok])
