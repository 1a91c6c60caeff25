"""Helpers shared by the parity tests: golden case -> model arguments."""
import numpy as np

# golden cases whose reference result itself is not reproducible to 1e-9: default-mode
# correction with migration ("stable": false in evals.json) and the two cases where the reference's
# least-squares solve runs away to rates ~1e7 on an unidentifiable interval
RUNAWAY = ("c3_band_to_split", "c5_bs3_band")


def grid_of(ds, case):
    """times, lambdas, integer split after the constructor's fractional-split surgery (MigrationInference.py:89-99)."""
    d = ds[case["dataset"]]
    times = [float(v) for v in d["times"]]
    lam = [[float(v[0]), float(v[1])] for v in d["lambdas"]]
    st = case["splitT"]
    frac = st % 1
    st = int(st)
    if frac != 0.0:
        t1 = frac * times[st]
        t2 = times[st] - t1
        times[st] = t1
        times.insert(st + 1, t2)
        lam.insert(st + 1, list(lam[st]))
        st += 1
    return times, lam, st, int(d["sampleDate"])


def bands_pulses(case):
    """0-based (pop, start, end, value, opt) / (pop, time, value, opt) with optimiser indices in MapParameters order."""
    bands, pulses, k = [], [], 0
    for m in case["mi"]:
        opt = k if int(m[4]) == 1 else -1
        k += int(m[4]) == 1
        bands.append((int(m[0]) - 1, int(m[1]), int(m[2]), float(m[3]), opt))
    for p in case["pu"]:
        opt = k if int(p[3]) == 1 else -1
        k += int(p[3]) == 1
        pulses.append((int(p[0]) - 1, int(p[1]), float(p[2]), opt))
    return bands, pulses


def flags_of(case):
    f = case["flags"]
    return (0 if f["trueEPS"] else 1) | (2 if f["cpfit"] else 0) | (4 if f["smooth"] else 0) | (8 if f["unfolded"] else 0)


def sfs_of(ds, case):
    d = ds[case["dataset"]]
    return list(d["sfs"]) if case.get("bs", -1) < 0 else list(d["bs_rows"][case["bs"]])


def relerr(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def end_to_end_gated(case):
    """True where the reference is reproducible to 1e-9 end to end (SURVEY.md 7.3 / BASELINE.md 3.7)."""
    return bool(case["stable"]) and case["name"] not in RUNAWAY


def random_jsfs_cases(n, seed=2026):
    """Random models for the JSFS stage alone (trueEPS: the given rates ARE the model rates): short random grids, random
    rates over two decades, random split (sometimes at the end of the grid = infinite last interval with migration),
    sampling date, bands (fixed rates, both demes, possibly up to the split) and pulses -- every segment type and event
    combination of the segment pre-pass.  Returned in the layout of tests/golden/evals.json cases (+ "grid")."""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        numT = int(rng.integers(6, 16))
        times = (10 ** rng.uniform(-2.5, -0.3, numT - 1)).tolist()
        lam = (10 ** rng.uniform(-0.7, 1.0, (numT, 2))).tolist()
        no_split = rng.random() < 0.2
        st = numT if no_split else int(rng.integers(1, numT))
        sd = int(rng.integers(0, st + 1)) if rng.random() < 0.4 else 0
        mi, pu = [], []
        used = [[False] * numT, [False] * numT]
        for _ in range(int(rng.integers(0, 4))):
            pop = int(rng.integers(0, 2))
            a = int(rng.integers(sd, max(sd + 1, st)))
            b = int(rng.integers(a + 1, st + 1)) if a + 1 <= st else a + 1
            if b > numT or any(used[pop][a:b]):
                continue
            for i in range(a, b):
                used[pop][i] = True
            mi.append([pop + 1, a, b, float(10 ** rng.uniform(-2, 0.7)), 0])
        if no_split and not (used[0][numT - 1] or used[1][numT - 1]):
            a = max(sd, numT - 2)
            if any(used[0][a:numT]):
                continue
            for i in range(a, numT):
                used[0][i] = True
            mi.append([1, a, numT, float(10 ** rng.uniform(-1, 0.5)), 0])
        ptimes = set()
        for _ in range(int(rng.integers(0, 3))):
            t = int(rng.integers(sd, st)) if st > sd else None
            if t is None or t in ptimes:
                continue
            ptimes.add(t)
            pu.append([int(rng.integers(1, 3)), t, float(rng.uniform(0.01, 0.6)), 0])
        out.append({"name": "rnd%d" % len(out), "grid": (times, lam, st, sd), "mi": mi, "pu": pu, "params": [],
                    "flags": dict(trueEPS=True, cpfit=False, smooth=False, unfolded=bool(rng.integers(0, 2))),
                    "sfs": [float(1000 + 7 * 300)] + rng.integers(50, 600, 7).astype(float).tolist()})
    return out


def check_solver_trace(gold, trace, lc_raw, name):
    """Iterate-level parity of the correction chain with the reference's scipy.optimize.least_squares calls
    (tests/golden/solver.json, CorrectLambda.py:85, 260, 303, 305).  trace[numT][2] = (nfev, status) per interval from the
    device (misti_eval_io.solve_trace; (0, -9) = closed form, no solver); lc_raw (nullable) = per-interval solutions
    BEFORE smoothing, in the reference's units.  Asserted: over the prefix of calls on which the reference determines its
    own iterates (`stable_calls`: same counts and solutions to 1e-9 when every parameter moves by one ulp) the evaluation
    counts and termination reasons are EQUAL, call by call; the first call past the prefix still starts from identical
    inputs, so where the reference's three probe runs agree on its counts the port must too.  Returns the number of
    calls compared and how many of ALL calls (also past the prefix, where the inputs have already drifted) have equal
    counts -- reported by the callers."""
    calls = gold["calls"]
    got = [(int(n), int(s)) for n, s in trace if int(s) != -9]
    stable = gold["stable_calls"]
    assert len(got) >= min(stable + 1, len(calls)), (name, len(got), len(calls))
    for k in range(stable):
        assert got[k] == (calls[k]["nfev"], calls[k]["status"]), (name, "call", k, got[k], calls[k]["nfev"], calls[k]["status"])
    checked = stable
    if stable < len(calls) and calls[stable]["probe_agree"]:
        assert got[stable] == (calls[stable]["nfev"], calls[stable]["status"]), (name, "first unstable call", stable, got[stable])
        checked += 1
    equal = sum(1 for k in range(min(len(got), len(calls))) if got[k] == (calls[k]["nfev"], calls[k]["status"]))
    return checked, equal, len(calls)
