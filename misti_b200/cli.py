"""Command line: the reference's MiSTI.py arguments (MiSTI.py:43-140) on the B200 path, plus in-process sweeps.

    python -m misti_b200.cli m1.psmc m2.psmc m.sfs 40 -uf -mi 2 5 12 0.8 1 --cpfit            # as MiSTI.py
    python -m misti_b200.cli m1.psmc m2.psmc bs.sfs 40 --st-grid 36 44 --bs-rows 0 1000 ...   # replaces the bash loops

Single run: prints the result line of MiSTI.py:240 ("bs_id = ... splitT = ... time = ... migration rates ...
llh = ...") and writes the `.mi` file under the same condition as the reference (only with -bs 0, MiSTI.py:248).
Sweep (--st-grid and/or --bs-rows): one result line per (bootstrap row, split time), all fits advanced in lock
step on the device (misti_b200.sweep); `-mi` start / end may contain the token `st` (e.g. `-mi 1 4 st 3 1`, the
test.bs/din_sar.bs.sh layout) which is replaced by each split time of the grid.
"""
import argparse
import os
import sys
import time
from math import ceil


def build_parser():
    p = argparse.ArgumentParser(description="Migration inference from PSMC (B200 evaluation path).")
    p.add_argument("fpsmc1")
    p.add_argument("fpsmc2")
    p.add_argument("fjafs")
    p.add_argument("st", type=float, help="split time")
    p.add_argument("-o", "--fout", default="")
    p.add_argument("-wd", default="")
    p.add_argument("-tol", type=float, default=1e-4)
    p.add_argument("-mth", type=float, default=0.0)
    p.add_argument("-mi", nargs=5, action="append", default=[])
    p.add_argument("-pu", nargs=4, action="append", default=[])
    p.add_argument("--sdate", type=float, default=0)
    p.add_argument("--hetloss", "-hl", nargs=2, type=float)
    p.add_argument("--discr", "-d", type=int, default=1, help="accepted and ignored, as in the reference")
    p.add_argument("-rd", type=int, default=-1)
    p.add_argument("--funits", default="setunits.txt")
    p.add_argument("-uf", action="store_true")
    p.add_argument("--nosmooth", action="store_true")
    p.add_argument("--trueEPS", action="store_true")
    p.add_argument("--cpfit", action="store_true")
    p.add_argument("--bsMode", "-bs", type=int, default=-1)
    p.add_argument("--psmcMode", "-pm", type=int, default=0, help="1: re-estimate both trajectories on the average collapsed grid "
                   "with the split time (st, then in years) inserted (migrationIO.ReadPSMC1)")
    p.add_argument("--debug", action="store_true")
    # additions
    p.add_argument("--st-grid", nargs=2, type=int, metavar=("FIRST", "LAST"), help="sweep the integer split times FIRST..LAST")
    p.add_argument("--bs-rows", nargs=2, type=int, metavar=("FIRST", "LAST"), help="sweep the rows FIRST..LAST of the JSFS file")
    p.add_argument("--globalOpt", action="store_true", help="basin-hopping instead of a single Nelder-Mead (Solve(globalOpt=True))")
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--device", type=int, default=0)
    return p


def _subst(mi, st):
    """-mi / -pu descriptors with the token `st` (the split time of the model, as in test.bs/din_sar.bs.sh: `-mi 1 4 ${st} 3 1`)
    replaced.  The token stands for an interval INDEX (SetModel takes int() of it): a whole split time is written as an
    integer ("40", not "40.0" -- argparse delivers -st as a float), a fractional one has no index to stand for."""
    if not any(str(v) == "st" for el in mi for v in el):
        return [list(el) for el in mi]
    if float(st) != int(float(st)):
        print("The `st` token in -mi / -pu needs a whole split time (got %s)." % st)
        sys.exit(0)
    return [[str(int(float(st))) if str(v) == "st" else v for v in el] for el in mi]


def main(argv=None):
    from . import io as mio
    from .inference import MigrationInference
    from .sweep import Sweep
    t0 = time.time()
    a = build_parser().parse_args(argv)
    units = mio.Units.from_file(a.funits, hetloss=a.hetloss or (0.0, 0.0))
    units.PrintUnits()
    print(" ".join(sys.argv if argv is None else ["misti_b200.cli"] + list(argv)))
    f1, f2, fj = (os.path.join(a.wd, f) for f in (a.fpsmc1, a.fpsmc2, a.fjafs))
    print("Reading from files:")
    print("pop1\t", f1)
    print("pop2\t", f2)
    print("jafs\t", fj)
    jafs = mio.read_jafs(fj, silent_mode=False)
    if a.psmcMode == 0:
        inp = mio.read_psmc(f1, f2, a.sdate, a.rd, units)
    else:  # MiSTI.py:190-195: st is a time in years and becomes the index of that time in the re-estimated grid
        if a.st_grid:
            print("--st-grid sweeps split indices of one grid; with -pm 1 the grid depends on the split time.")
            sys.exit(0)
        inp = mio.read_psmc1(f1, f2, a.rd, divergenceTime=a.st, units=units)
        if inp.divergenceTime != -1:
            a.st = inp.divergenceTime
    kw = dict(smooth=not a.nosmooth, unfolded=a.uf, trueEPS=a.trueEPS, cpfit=a.cpfit, sampleDate=inp.sampleDateDiscr, mixtureTH=a.mth)
    t1 = time.time()
    if a.st_grid or a.bs_rows:
        rows_idx = list(range(a.bs_rows[0], a.bs_rows[1] + 1)) if a.bs_rows else None
        rows = [jafs.jafs[r] for r in rows_idx] if rows_idx else [mio.column_sums(jafs.jafs) if a.bsMode == -1 else jafs.jafs[a.bsMode]]
        sts = list(range(a.st_grid[0], a.st_grid[1] + 1)) if a.st_grid else [a.st]
        sw = Sweep(inp.times, inp.lambdas, rows, device=a.device, **kw)
        for st in sts:
            sw.add_model(st, _subst(a.mi, st), _subst(a.pu, st))
        res = sw.solve(tol=a.tol, globalOpt=a.globalOpt, seed=a.seed)
        for k in range(len(res["llh"])):
            bs_id = rows_idx[int(res["row"][k])] if rows_idx else a.bsMode
            print(sw.result_line(res, k, inp.scaleTime, bs_id=bs_id))
        t2 = time.time()
        print("Sweep: %d fits, %d objective evaluations in %d launches" % (len(res["llh"]), res["evaluations"], res["launches"]))
    else:
        sfs = mio.column_sums(jafs.jafs) if a.bsMode == -1 else jafs.jafs[a.bsMode]
        M = MigrationInference(inp.times, inp.lambdas, sfs, a.st, a.mi, a.pu, thrh=[inp.theta, inp.rho], Tpsmc=inp.Tpsmc,
                               enableOutput=False, device=a.device, **kw)
        sol = M.Solve(a.tol, globalOpt=a.globalOpt)
        print(sol)
        print("\nParameter estimates:")
        fixed = [float(el[3]) for el in a.mi if int(el[4]) == 0]
        fs = "fixed = [" + ", ".join(str(v) for v in fixed) + "]" if fixed else ""
        os_ = "optim = [" + ", ".join(str(v) for v in sol[0]) + "]" if len(sol[0]) > 0 else ""
        mig = fs + "\t" + os_ if fs and os_ else fs + os_
        print("bs_id =", a.bsMode, "\tsplitT =", a.st, "\ttime =", sum(inp.times[0:ceil(a.st)]) * inp.scaleTime,
              "\tmigration rates", mig, "\tllh =", sol[1])
        print("\n")
        t2 = time.time()
        if sol[1] == -10 ** 9:
            print("Failed to fit such a model.")
        elif a.bsMode == 0:
            mio.output_migration(os.path.join(a.wd, a.fout) if a.fout else "", sol[0], M, inp.scaleTime, inp.scaleEPS)
    MigrationInference.Report()
    print("Runtime:   optimisation", t2 - t1)
    print("           total       ", time.time() - t0)
    return 0


if __name__ == "__main__":
    sys.exit(main())
