"""Device numerics, compiled for the HOST (tests/hostsim, test-only): the very headers the CUDA kernels
include -- trust-region least squares, 3x3 expm, correction chain, uniformised JSFS stage, likelihood
tail -- checked against the golden vectors of the reference without a GPU."""
import ctypes

import numpy as np

from _cases import RUNAWAY, bands_pulses, end_to_end_gated, flags_of, grid_of, relerr, sfs_of
from misti_b200.engine import llh_constants

dp = ctypes.POINTER(ctypes.c_double)


def _arr(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(dp)


def _chain(lib, ds, case, with_trace=False):
    times, lam, st, sd = grid_of(ds, case)
    bands, pulses = bands_pulses(case)
    numT = len(lam)
    T, L = _arr(times), _arr(lam)
    Bn = _arr([[b[0], b[1], b[2], b[3], b[4]] for b in bands] or [[0] * 5])
    Pu = _arr([[p[0], p[1], p[2], p[3]] for p in pulses] or [[0] * 4])
    par = _arr(case["params"] or [0.0])
    lc, Pr, nfev = np.zeros((numT, 2)), np.zeros((numT + 1, 3, 2)), ctypes.c_int(0)
    trace = np.zeros((numT, 2), dtype=np.int32)
    rc = lib.hs_correct_lambdas(numT, st, sd, _p(T), _p(L), len(bands), _p(Bn), len(pulses), _p(Pu), len(case["params"]), _p(par),
                                flags_of(case), ctypes.c_double(0.0), _p(lc), _p(Pr), ctypes.byref(nfev),
                                trace.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)) if with_trace else None)
    if with_trace:
        return rc, lc, Pr, nfev.value, trace
    return rc, lc, Pr, nfev.value


def _jsfs(lib, ds, case, lc):
    times, lam, st, sd = grid_of(ds, case)
    bands, pulses = bands_pulses(case)
    T = _arr(times)
    Bn = _arr([[b[0], b[1], b[2], b[3], b[4]] for b in bands] or [[0] * 5])
    Pu = _arr([[p[0], p[1], p[2], p[3]] for p in pulses] or [[0] * 4])
    par = _arr(case["params"] or [0.0])
    uf = case["flags"]["unfolded"]
    row = _arr(sfs_of(ds, case))
    d = row[1:]
    drow = _arr(list(d) + [0.0]) if uf else _arr([d[0] + d[6], d[1] + d[5], d[2] + d[4], d[3], 0, 0, 0, 0.0])
    drow[7] = llh_constants([row], uf)[0]
    raw, jn, llh, terms = np.zeros(7), np.zeros(7), ctypes.c_double(0), ctypes.c_int(0)
    lc = _arr(lc)
    rc = lib.hs_jsfs(len(lam), st, sd, _p(T), len(bands), _p(Bn), len(pulses), _p(Pu), len(case["params"]), _p(par), _p(lc), int(uf),
                     _p(drow), _p(raw), _p(jn), ctypes.byref(llh), ctypes.byref(terms))
    return rc, raw, jn, llh.value, terms.value


def test_expm3_and_inverse(hostsim):
    """3x3 expm / inverse on matrices of the shape the correction uses (CorrectLambda.SetMatrix, CorrectLambda.py:55-56)."""
    from scipy.linalg import expm, inv
    rng = np.random.default_rng(0)
    for T in (1e-3, 0.05, 0.5, 1.0, 4.0, 30.0):
        for _ in range(20):
            l0, l1, m0, m1 = rng.uniform(0.0, 5.0, 4)
            M = _arr([[-2 * m0 - l0, 0, m1], [0, -2 * m1 - l1, m0], [2 * m0, 2 * m1, -m0 - m1]])
            A = _arr(M * T)
            E = np.zeros((3, 3))
            hostsim.hs_expm3(_p(A), _p(E))
            assert np.max(np.abs(E - expm(A))) <= 5e-15 * max(1.0, np.abs(A).sum(axis=0).max())
            Mi = np.zeros((3, 3))
            assert hostsim.hs_inv3(_p(M), _p(Mi)) == 1
            ref = inv(M)
            assert np.max(np.abs(Mi - ref)) <= 1e-13 * np.linalg.cond(M) * np.max(np.abs(ref))


def test_correction_chain_matches_reference(hostsim, golden_datasets, golden_cases):
    exact = 0
    for case in golden_cases:
        rc, lc, Pr, nfev = _chain(hostsim, golden_datasets, case)
        exp = case["expect"]
        if not exp["ok"]:
            assert rc in (1, 2), case["name"]
            continue
        if not end_to_end_gated(case):
            continue
        assert rc == 0, case["name"]
        err = relerr(lc, exp["lc"])
        assert err < 1e-9, (case["name"], err)
        exact += err == 0.0
        n = len(exp["Pr"])
        assert relerr(Pr[:n] + 1.0, np.array(exp["Pr"]) + 1.0) < 1e-10, case["name"]
    # the trust-region port is iterate-faithful: every no-migration case reproduces the rates bit for bit
    assert exact >= 20


def test_jsfs_stage_matches_reference_given_rates(hostsim, golden_datasets, golden_cases):
    for case in golden_cases:
        exp = case["expect"]
        if not exp["ok"] or case["name"] in RUNAWAY:
            continue
        rc, raw, jn, llh, terms = _jsfs(hostsim, golden_datasets, case, exp["lc"])
        assert rc == 0, case["name"]
        assert relerr(jn, exp["JAFS"]) < 1e-12, case["name"]
        assert relerr(llh, exp["llh"]) < 1e-10, case["name"]
        assert abs(raw.sum() / jn.sum() - raw.sum()) < 1e-9 * raw.sum()


def test_end_to_end_host_build(hostsim, golden_datasets, golden_cases):
    for case in golden_cases:
        exp = case["expect"]
        if not exp["ok"] or not end_to_end_gated(case):
            continue
        rc, lc, _, _ = _chain(hostsim, golden_datasets, case)
        assert rc == 0
        rc, raw, jn, llh, terms = _jsfs(hostsim, golden_datasets, case, lc)
        assert rc == 0
        assert relerr(jn, exp["JAFS"]) < 1e-9, case["name"]
        assert relerr(llh, exp["llh"]) < 1e-9, case["name"]


def test_runaway_rates_are_flagged_stiff(hostsim, golden_datasets, golden_cases):
    """where the reference's own least-squares solve runs away to rates ~1e7 the uniformisation would need
    millions of sweeps; the kernel reports MISTI_STIFF (5) instead."""
    for case in golden_cases:
        if case["name"] in RUNAWAY:
            rc, *_ = _jsfs(hostsim, golden_datasets, case, case["expect"]["lc"])
            assert rc == 5


def test_no_split_inside_grid(hostsim, golden_datasets):
    """splitT == numT: the last two-population interval is infinite; needs migration there."""
    from oracle.misti_oracle import OracleModel
    d = golden_datasets["ms_two_bands"]
    numT = len(d["lambdas"])
    mi = [[1, 0, numT, 0.7, 0], [2, 0, numT, 0.4, 0]]
    case = {"dataset": "ms_two_bands", "splitT": numT, "mi": mi, "pu": [], "params": [],
            "flags": dict(trueEPS=True, cpfit=False, smooth=False, unfolded=True), "bs": -1}
    om = OracleModel(d["times"], d["lambdas"], d["sfs"], numT, mi, [], trueEPS=True, unfolded=True)
    ref = om.likelihood([])
    rc, raw, jn, llh, terms = _jsfs(hostsim, golden_datasets, case, d["lambdas"])
    assert rc == 0
    assert relerr(jn, om.JAFS) < 1e-10
    assert relerr(llh, ref) < 1e-10
    case["mi"] = []
    rc, *_ = _jsfs(hostsim, golden_datasets, case, d["lambdas"])
    assert rc == 4  # infinite coalescent time, no migration


def test_segment_lists(hostsim, golden_datasets, golden_cases):
    """the per-item scalar pre-pass: zero-migration intervals merge into closed-form runs (type 2), broken by migration
    intervals (type 1, one segment each), the sampling date and pulses"""
    def types_of(case):
        rc, *_ = _jsfs(hostsim, golden_datasets, case, case["expect"]["lc"])
        assert rc == 0
        buf = (ctypes.c_int * 256)()
        n = hostsim.hs_segment_types(buf, 256)
        return list(buf[:n])
    by_name = {c["name"]: c for c in golden_cases}
    assert types_of(by_name["c1_st36_uf"]) == [2]                       # no migration at all: one run up to the split
    assert types_of(by_name["c2_cpfit_m0"]) == [2]                      # a band with rate 0 is no migration either
    assert types_of(by_name["c2_cpfit_m0.8"]) == [2] + [1] * 7 + [2]    # band over intervals 5..11 of 40
    assert types_of(by_name["c4_cpfit_band"]) == [2, 2] + [1] * 6 + [2]  # sampling date at 12, band 14..19
    for case in golden_cases:
        if not case["expect"]["ok"] or case["name"] in RUNAWAY:
            continue
        ty = types_of(case)
        times, lam, st, sd = grid_of(golden_datasets, case)
        n_mig = sum(1 for t in ty if t == 1)
        assert all(t in (1, 2, 4) for t in ty) and n_mig <= min(st, len(lam))
        if not case["mi"] and not case["flags"].get("trueEPS"):
            assert n_mig == 0
            # runs are broken only by the sampling date and by pulses
            assert len(ty) <= 1 + (1 if 0 < sd < st else 0) + len(case["pu"])


def _jsfs_random(lib, case, lam):
    times, _, st, sd = case["grid"]
    bands, pulses = bands_pulses(case)
    T, L = _arr(times), _arr(lam)
    Bn = _arr([[b[0], b[1], b[2], b[3], b[4]] for b in bands] or [[0] * 5])
    Pu = _arr([[p[0], p[1], p[2], p[3]] for p in pulses] or [[0] * 4])
    par = _arr([0.0])
    uf = case["flags"]["unfolded"]
    row = _arr(case["sfs"])
    d = row[1:]
    drow = _arr(list(d) + [0.0]) if uf else _arr([d[0] + d[6], d[1] + d[5], d[2] + d[4], d[3], 0, 0, 0, 0.0])
    drow[7] = llh_constants([row], uf)[0]
    raw, jn, llh, terms = np.zeros(7), np.zeros(7), ctypes.c_double(0), ctypes.c_int(0)
    rc = lib.hs_jsfs(len(lam), st, sd, _p(T), len(bands), _p(Bn), len(pulses), _p(Pu), 0, _p(par), _p(L), int(uf), _p(drow), _p(raw),
                     _p(jn), ctypes.byref(llh), ctypes.byref(terms))
    return rc, raw, jn, llh.value


def test_random_models_against_oracle(hostsim):
    """JSFS stage on random grids / rates / splits / sampling dates / bands / pulses vs the CPU oracle"""
    from _cases import random_jsfs_cases
    from oracle.misti_oracle import OracleModel
    seen = set()
    for case in random_jsfs_cases(40):
        times, lam, st, sd = case["grid"]
        om = OracleModel(times, lam, case["sfs"], st, case["mi"], case["pu"], trueEPS=True, unfolded=case["flags"]["unfolded"],
                         sampleDate=sd)
        ref = om.likelihood([])
        rc, raw, jn, llh = _jsfs_random(hostsim, case, om.lc)  # the oracle's rates (its post-split part is always re-fitted)
        assert rc == 0, case
        assert relerr(jn, om.JAFS) < 1e-11, (case["name"], relerr(jn, om.JAFS))
        assert relerr(llh, ref) < 1e-10, case["name"]
        buf = (ctypes.c_int * 256)()
        seen |= set(buf[:hostsim.hs_segment_types(buf, 256)])
    assert {1, 2, 4} <= seen  # swept intervals, closed-form runs and an infinite last interval all occurred


def test_split_at_the_ends_of_the_grid(hostsim, golden_datasets):
    """split index 0 / 1 (no or one two-population interval), last finite interval, and no split inside the grid"""
    from oracle.misti_oracle import OracleModel
    ds = golden_datasets["synthetic"]
    times, lam = ds["times"], ds["lambdas"]
    n = len(lam)
    for st, mi in ((0, []), (1, []), (1, [[1, 0, 1, 0.7, 0]]), (2, [[2, 1, 2, 1.3, 0]]), (n - 1, [[1, 100, n - 1, 0.2, 0]]),
                   (n, [[1, 120, n, 0.2, 0]])):
        case = {"grid": (times, lam, st, 0), "mi": mi, "pu": [], "params": [],
                "flags": dict(trueEPS=True, cpfit=False, smooth=False, unfolded=True), "sfs": ds["sfs"]}
        om = OracleModel(times, lam, ds["sfs"], st, mi, [], trueEPS=True, unfolded=True)
        ref = om.likelihood([])
        rc, raw, jn, llh = _jsfs_random(hostsim, case, om.lc)
        assert rc == 0, (st, mi)
        assert relerr(jn, om.JAFS) < 1e-12 and relerr(llh, ref) < 1e-12, (st, mi)


def test_post_split_pass_by_lane_groups(hostsim, golden_datasets):
    """cpfit mode: the post-split coefficients summed slice by slice over the 16-lane table of a model (the JSFS kernel's
    pass) equal the sequential pass of the correction chain (MigrationInference.py:356-374), and the rate computed on
    request equals the rate that pass writes."""
    ds = golden_datasets["synthetic"]
    T0, L = _arr(ds["times"]), _arr(ds["lambdas"])
    numT = len(L)
    rng = np.random.default_rng(5)
    for trial in range(40):
        T = T0.copy()
        splitT = int(rng.integers(1, numT))
        if trial % 4 == 0 and splitT < numT - 1:
            T[int(rng.integers(splitT, numT - 1))] = 0.0  # a zero-length interval is skipped (:358-360)
        ed = float(np.exp(rng.uniform(-3, 3)))
        for lanes in (1, 16):
            seq, grp, lc, req = np.zeros(3), np.zeros(3), np.zeros((numT, 2)), np.zeros(numT)
            hostsim.hs_post_split(numT, splitT, _p(T), _p(L), ctypes.c_double(ed), lanes, _p(seq), _p(grp), _p(lc), _p(req))
            # (the sequential pass gets ed as exp(log(ed)): one rounding apart)
            assert relerr(grp, seq) < 1e-13, (splitT, lanes, grp, seq)
            assert relerr(req[splitT:], lc[splitT:, 0]) < 1e-13


def _coal_cases():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "coal.json")) as f:
        return json.load(f)["cases"]


def _mu_last(case):
    """migration rates of the last interval before the split (what the reference's helper is left with, MigrationInference.py:324)"""
    t = case["splitT"] - 1
    mu = [0.0, 0.0]
    for pop, a, b, val, _ in case["mi"]:
        if a <= t < b:
            mu[int(pop) - 1] = float(val)
    return mu


def test_forward_map_matches_reference(hostsim):
    """the device's CoalescentRates (coalescent_rates_item, host build) against outputs of the unmodified reference"""
    for case in _coal_cases():
        T, L = _arr(case["times"]), _arr(case["lambdas"])
        numT, st = len(L), case["splitT"]
        Bn = _arr([[m[0] - 1, m[1], m[2], m[3], -1] for m in case["mi"]] or [[0] * 5])
        Pu = _arr([[p[0] - 1, p[1], p[2], -1] for p in case["pu"]] or [[0] * 4])
        mu = _mu_last(case)
        lh, Pr, par = np.zeros((numT, 2)), np.zeros((st + 1, 3, 2)), _arr([0.0])
        hostsim.hs_coal_rates(numT, st, _p(T), _p(L), len(case["mi"]), _p(Bn), len(case["pu"]), _p(Pu), 0, _p(par),
                              ctypes.c_double(mu[0]), ctypes.c_double(mu[1]), _p(lh), _p(Pr))
        assert relerr(lh, case["expect"]["lh"]) < 1e-11, case["name"]
        assert relerr(Pr + 1.0, np.array(case["expect"]["Pr"]) + 1.0) < 1e-13, case["name"]


def test_tiny_migration_rates_host_build(hostsim, golden_datasets):
    """CPU counterpart of tests/test_gpu_parity.py::test_tiny_migration_rates_are_continuous: for a tiny positive rate the
    reference's SolveDifEq (inv(M) of a nearly singular generator, MigrationInference.py:530-540) is off by about 1e-16 / m,
    which the oracle -- same scipy calls -- reproduces; the device numerics (uniformisation, no inverse) stay Lipschitz in m
    down to zero and agree with the oracle at m = 0 and where the oracle is accurate (m >= 1e-5)."""
    from oracle.misti_oracle import OracleModel
    ds = golden_datasets["synthetic"]
    mi, pu = [[2, 15, 23, 1.418, 1], [1, 8, 13, 1.358, 1]], [[1, 35, 0.05, 1]]
    x0, x2 = 1.7209222496044632, 0.4466402770983831

    def host(m):
        case = {"dataset": "synthetic", "splitT": 51, "mi": mi, "pu": pu, "params": [x0, m, x2],
                "flags": dict(trueEPS=False, cpfit=True, smooth=True, unfolded=True)}
        rc, lc, _, _ = _chain(hostsim, golden_datasets, case)
        assert rc == 0
        rc, _, _, llh, _ = _jsfs(hostsim, golden_datasets, case, lc)
        assert rc == 0
        return llh

    def oracle(m):
        om = OracleModel(ds["times"], ds["lambdas"], ds["sfs"], 51, mi, pu, cpfit=True, smooth=True, unfolded=True)
        return om.likelihood([x0, m, x2])

    ref0, h0 = oracle(0.0), host(0.0)
    assert relerr(h0, ref0) < 1e-9
    for m in (1e-5, 1e-3):
        assert relerr(host(m), oracle(m)) < 1e-9, m
    s = (host(1e-5) - h0) / 1e-5
    for m in (1e-14, 7.84e-12, 1e-9, 1e-7):
        assert abs(host(m) - h0) <= 1e-10 * abs(ref0) + 2.0 * abs(s) * m, m
    # the reference's algorithm at the same points: 14 % off at 1e-14, 7e-5 at 7.84e-12
    assert abs(oracle(1e-14) - ref0) > 1e-2 * abs(ref0) and abs(oracle(7.84e-12) - ref0) > 1e-6 * abs(ref0)


def test_random_layouts_end_to_end_host_build(hostsim, golden_datasets):
    """--cpfit mode end to end on random layouts (split time, one or two optimised bands, sometimes a pulse; plain and
    ancient-sample PSMC pair) with random parameters: correction chain + segment pre-pass + sweep + likelihood of the device
    numerics (host build) against the oracle; an item fails in both or in neither.  (The GPU-sized version of this sweep is
    tools/fuzz_parity.py part B.)"""
    from oracle.misti_oracle import OracleModel
    rng = np.random.default_rng(77)
    compared = 0
    for dsn in ("synthetic", "synthetic_ancient"):
        ds = golden_datasets[dsn]
        sd = int(ds["sampleDate"])
        for _ in range(8):
            st = int(rng.integers(max(sd + 8, 25), 61))
            a = int(rng.integers(sd, st - 3))
            b = int(rng.integers(a + 1, min(st, a + 12) + 1))
            mi = [[int(rng.integers(1, 3)), a, b, 0.5, 1]]
            pu = [[int(rng.integers(1, 3)), int(rng.integers(sd, st)), 0.05, 1]] if rng.random() < 0.3 else []
            par = [float(rng.uniform(0.0, 2.0))] + ([float(rng.uniform(0.0, 0.5))] if pu else [])
            case = {"dataset": dsn, "splitT": st, "mi": mi, "pu": pu, "params": par,
                    "flags": dict(trueEPS=False, cpfit=True, smooth=True, unfolded=True)}
            om = OracleModel(ds["times"], ds["lambdas"], ds["sfs"], st, mi, pu, cpfit=True, smooth=True, unfolded=True, sampleDate=sd)
            ref = om.likelihood(par)
            rc, lc, _, _ = _chain(hostsim, golden_datasets, case)
            if not np.isfinite(ref):
                assert rc != 0, (dsn, st, mi, pu, par)
                continue
            assert rc == 0, (dsn, st, mi, pu, par)
            rc, raw, jn, llh, terms = _jsfs(hostsim, golden_datasets, case, lc)
            assert rc == 0
            assert relerr(jn, om.JAFS) < 1e-9 and relerr(llh, ref) < 1e-9, (dsn, st, mi, pu, par)
            compared += 1
    assert compared >= 10


def test_solver_iterates_match_the_reference_call_by_call(hostsim, golden_datasets, golden_cases, golden_solver):
    """Every scipy.optimize.least_squares call of the reference's correction (CorrectLambda.py:85, 260, 303, 305) against
    the port's solve of the same interval: evaluation count (`nfev`) and termination status, call by call, in every mode
    (bounded TRF n = 2 and n = 1 of the default mode, unbounded n = 2 with migration, default and cpfit residuals).
    Default mode WITH migration is ulp-chaotic in the reference itself, so there the comparison runs up
    to and including the first call whose result the reference does not determine (see _cases.check_solver_trace)."""
    from _cases import check_solver_trace
    by_name = {c["name"]: c for c in golden_cases}
    total_equal = total_calls = 0
    for name, gold in golden_solver.items():
        case = by_name[name]
        rc, lc, Pr, nfev, trace = _chain(hostsim, golden_datasets, case, with_trace=True)
        checked, equal, n = check_solver_trace(gold, trace, lc, name)
        total_equal += equal
        total_calls += n
        if gold["stable_calls"] == n:  # the reference determines the whole chain: every count is the reference's
            assert equal == n, (name, equal, n)
            assert nfev == sum(c["nfev"] for c in gold["calls"]), name
    assert total_equal >= 0.97 * total_calls, (total_equal, total_calls)


def test_tiny_migration_rates_against_50_digit_values(hostsim, golden_datasets):
    """Where a fitted rate walks to zero (m = 1e-14 ... 1e-4) the REFERENCE loses ~5e-17 / m: SolveDifEq integrates with
    inv(M)(P1 - P0) of a generator that is singular at m = 0 (MigrationInference.py:530-540).  tests/golden/tiny_rate_exact.json
    holds, per rate, the reference's result and the 50-digit value of the same stage for the reference's own rates (mpmath,
    tests/golden/gen_tiny_rate_exact.py): the reference is off by 5e-3 at m = 1e-14 and 1.6e-9 at m = 1e-8; the
    uniformised sweep has no inverse and must stay at 1e-12 from the exact value throughout."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_rate_exact.json")) as f:
        gold = json.load(f)
    for pt in gold["points"]:
        case = dict(gold["case"], params=[pt["m"]])
        rc, raw, jn, llh, terms = _jsfs(hostsim, golden_datasets, case, pt["lc"])
        assert rc == 0
        exact = [float(v) for v in pt["jafs_exact"]]
        assert relerr(jn, exact) < 1e-12, (pt["m"], relerr(jn, exact))
        assert relerr(llh, pt["llh_exact"]) < 1e-11, pt["m"]
        if pt["m"] <= 1e-10:  # ... where the reference itself is outside 1e-9 of the exact value
            assert pt["reference_jafs_relerr_vs_exact"] > 1e-9 and relerr(jn, pt["reference_jafs"]) > 1e-9
