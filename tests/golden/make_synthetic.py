#!/usr/bin/env python3
"""Generate the model-consistent synthetic inputs of SURVEY.md section 8(d).

Run in the build container only (needs /root/reference through ref_shim); the outputs under
data/synthetic/ are committed so that bench.py, smoke() and the GPU tests never need the reference.

  m1.psmc, m2.psmc : two 64-interval PSMC outputs whose apparent coalescence rates are the
                     reference's own forward map (MigrationInference.CoalescentRates,
                     MigrationInference.py:542-564) of a ground-truth two-population model
                     (split index 40, band "-mi 2 5 12 0.8").
  m.sfs            : 200 chunk rows [12.5 Mb, multinomial(20000, truth JSFS)], default_rng(7).
  bs.sfs           : row 0 = column sums, rows 1..1000 = BootstrapJAFS semantics
                     (migrationIO.py:506-524) with random.Random(12345).
"""
import os
import sys
import math
import random
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(HERE, "..", "..", "data", "synthetic")
THETA0 = 0.05
THETA_SCALE = (1.0, 1.37)
NPSMC = 64
TRUTH_SPLIT = 40
TRUTH_MI = [[2, 5, 12, 0.8, 0]]


def psmc_grid():
    return [0.1 * (math.exp(k / 63.0 * math.log(1 + 10 * 15)) - 1) for k in range(NPSMC)]


def ne1(t): return 1 + 0.5 * math.sin(2 * math.log(t + 0.02))
def ne2(t): return 0.6 + 0.3 * math.cos(1.5 * math.log(t + 0.02))
def nea(t): return 1.5 + 0.8 * math.sin(math.log(t + 0.05))


def main():
    R = ref_shim.load()
    mio = R["migrationIO"]
    ref_shim.reset_units(mio)
    os.makedirs(OUT, exist_ok=True)
    g = psmc_grid()
    grids = [[t * s for t in g] for s in THETA_SCALE]
    merged = sorted(grids[0] + grids[1][1:])
    assert len(merged) == 2 * NPSMC - 1 and len(set(merged)) == len(merged)
    numT = len(merged)
    times = [b - a for a, b in zip(merged[:-1], merged[1:])]
    lam = []
    for i, t in enumerate(merged):
        if i < TRUTH_SPLIT:
            lam.append([1.0 / ne1(t), 1.0 / ne2(t)])
        else:
            lam.append([1.0 / nea(t), 1.0 / nea(t)])
    MI = R["MigrationInference"]
    M = MI(list(times), [list(v) for v in lam], [1] * 8, TRUTH_SPLIT, TRUTH_MI, [], unfolded=True, trueEPS=True)
    M.JAFSLikelihood([])
    truth = [float(v) for v in M.JAFS]
    M.CoalescentRates()
    lh = [[float(v[0]), float(v[1])] for v in M.lh]
    # write PSMC files
    for gi in (0, 1):
        own = grids[gi]
        sizes = []
        for k in range(NPSMC):
            lo = own[k]
            hi = own[k + 1] if k + 1 < NPSMC else None
            num, den = 0.0, 0.0
            for j in range(numT - 1):
                if merged[j] >= lo and (hi is None or merged[j] < hi):
                    num += lh[j][gi] * times[j]
                    den += times[j]
            if den > 0 and hi is not None:
                rate = num / den
            else:
                rate = lh[numT - 1][gi]
            sizes.append(1.0 / rate)
        sc = THETA_SCALE[gi]
        with open(os.path.join(OUT, "m%d.psmc" % (gi + 1)), "w") as f:
            f.write("MM\tsynthetic PSMC output for misti-b200 (SURVEY 8d)\n")
            f.write("RD\t0\n")
            f.write("TR\t%.17g\t%.17g\n" % (THETA0 * sc, 0.01))
            for k in range(NPSMC):
                f.write("RS\t%d\t%.17g\t%.17g\t0\t0\t0\n" % (k, own[k] / sc, sizes[k] / sc))
            f.write("PA\t4+25*2+4+6 %.17g 0.01 15\n" % (THETA0 * sc))
            f.write("//\n")
    # chunked JSFS
    rng = np.random.default_rng(7)
    p = np.array(truth) / sum(truth)
    rows = []
    for _ in range(200):
        c = rng.multinomial(20000, p)
        rows.append([12500000.0] + [float(v) for v in c])
    hdr = "#MiSTI_JSFS version 1.0\n#pop1\tA\n#pop2\tB\n" + "\t".join(
        ["total", "0100", "1100", "0001", "0101", "1101", "0011", "0111"]) + "\n"
    with open(os.path.join(OUT, "m.sfs"), "w") as f:
        f.write(hdr)
        for r in rows:
            f.write("\t".join(repr(v) for v in r) + "\n")
    # bootstrap rows (BootstrapJAFS semantics with an explicit seed)
    r = random.Random(12345)
    glen = sum(x[0] for x in rows)
    bs = [[sum(x[i] for x in rows) for i in range(8)]]
    for _ in range(1000):
        s = [0.0] * 8
        while s[0] < glen:
            k = r.randint(0, len(rows) - 1)
            for i in range(8):
                s[i] += rows[k][i]
        bs.append(s)
    with open(os.path.join(OUT, "bs.sfs"), "w") as f:
        f.write(hdr)
        for row in bs:
            f.write("\t".join(repr(v) for v in row) + "\n")
    with open(os.path.join(OUT, "setunits.txt"), "w") as f:
        f.write("mutRate=1.25e-8\nbinsize=100\nN0=10000\ngenTime=1\n")
    with open(os.path.join(OUT, "truth.txt"), "w") as f:
        f.write("truth_jsfs\t" + "\t".join(repr(v) for v in truth) + "\n")
    print("truth JSFS", truth)


if __name__ == "__main__":
    main()
