#!/usr/bin/env python3
"""Golden vectors for the reference's second PSMC input mode (MiSTI.py -pm 1: migrationIO.ReadPSMC1, migrationIO.py:297-344,
on top of psmc.PSMC, psmc.py:25-163): the two trajectories re-estimated on the average of their collapsed (pattern) grids
with the split time, given in years, inserted.  The synthetic PSMC files carry no `MM pattern:` line, which this mode needs,
so the cases prepend one (and, for the second case, a thinner pattern and an earlier round).  Also stores what the reference's
MigrationInference makes of that input (expected JSFS, likelihood), so that the command line can be checked end to end.
Run in the build container only (needs /root/reference through ref_shim); writes psmc1.json next to this script."""
import contextlib
import io
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

CASES = [
    {"name": "pattern_4_25x2_4_6_split_8000y", "pattern": "4+25*2+4+6", "st_years": 8000.0, "RD": -1, "unfolded": True},
    {"name": "pattern_16x4_split_30000y", "pattern": "16*4", "st_years": 30000.0, "RD": -1, "unfolded": False},
    {"name": "pattern_1_2_61_no_split", "pattern": "1+2+61", "st_years": -1, "RD": -1, "unfolded": True},
]


def with_pattern(src, dst, pattern):
    with open(src) as f:
        body = f.read()
    with open(dst, "w") as f:
        f.write("MM\tpattern:%s, n:63, n_free_lambdas:%d\n" % (pattern, len(pattern.split("+"))))
        f.write(body)


def main():
    R = ref_shim.load()
    mio = R["migrationIO"]
    syn = os.path.join(ROOT, "data", "synthetic")
    u = mio.Units()
    u.SetUnitsFromFile(os.path.join(syn, "setunits.txt"))
    with open(os.path.join(HERE, "datasets.json")) as f:
        sfs = json.load(f)["datasets"]["synthetic"]["sfs"]  # column sums of data/synthetic/m.sfs (MiSTI.py:172-178)
    out = []
    with tempfile.TemporaryDirectory() as tmp:
        for c in CASES:
            f1, f2 = os.path.join(tmp, "p1.psmc"), os.path.join(tmp, "p2.psmc")
            with_pattern(os.path.join(syn, "m1.psmc"), f1, c["pattern"])
            with_pattern(os.path.join(syn, "m2.psmc"), f2, c["pattern"])
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                d = mio.ReadPSMC1(f1, f2, c["RD"], divergenceTime=c["st_years"])
            rec = dict(c)
            rec.update({"times": [float(v) for v in d.times], "lambdas": [[float(a), float(b)] for a, b in d.lambdas],
                        "divTime": int(d.divergenceTime), "scaleTime": float(d.scaleTime), "theta": float(d.theta)})
            if d.divergenceTime != -1:
                with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                    M = R["MigrationInference"](d.times, d.lambdas, sfs, d.divergenceTime, [], [], thrh=[d.theta, d.rho],
                                                Tpsmc=d.Tpsmc, enableOutput=False, smooth=True, unfolded=c["unfolded"],
                                                sampleDate=d.sampleDateDiscr, cpfit=True)
                    sol = M.Solve(1e-4)
                rec["llh_cpfit"] = float(sol[1])
                rec["jafs_cpfit"] = [float(v) for v in M.JAFS]
            out.append(rec)
    import numpy
    import scipy
    with open(os.path.join(HERE, "psmc1.json"), "w") as f:
        json.dump({"meta": {"generator": "tests/golden/gen_psmc1_golden.py", "numpy": numpy.__version__, "scipy": scipy.__version__},
                   "cases": out}, f, indent=1)
    for r in out:
        print(r["name"], len(r["times"]), r["divTime"], r.get("llh_cpfit"))


if __name__ == "__main__":
    main()
