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
