"""Command line of the reference's TestModel.py (TestModel.py:40-125) on the B200 path: the expected joint SFS of a model
given as an ms command line, its likelihood against a JSFS file, and the forward map to PSMC-apparent rates.

    python -m misti_b200.testmodel "<ms command>" [jsfs file] [-uf] [-bs N] [-o out.mi] [--funits setunits.txt]

Prints what TestModel.py prints ("Expected SFS ...", and with a JSFS file "Data SFS", "data llh under the model is",
"maximum of the llh function is", the bootstrap intervals), then runs CoalescentRates and writes the `.mi` file.  The
bootstrap of the reference refers to an undefined variable (TestModel.py:112) and cannot run there; here the -bs
replicates (BootstrapJAFS semantics, seeded) are scored in ONE batched evaluation against the model's spectrum.
"""
import argparse
import math
import random
import sys


def build_parser():
    p = argparse.ArgumentParser(description="Expected JSFS and likelihood of an ms-style model (B200 evaluation path).")
    p.add_argument("msstring", help="ms style command")
    p.add_argument("fjafs", nargs="?", default="", help="joint allele frequency spectrum file")
    p.add_argument("--funits", default="setunits.txt")
    p.add_argument("-uf", action="store_true", help="unfolded spectrum")
    p.add_argument("--bsSize", "-bs", type=int, default=0, help="number of bootstrap repetitions")
    p.add_argument("-o", "--fout", default="")
    p.add_argument("--seed", type=int, default=0, help="seed of the bootstrap (addition)")
    p.add_argument("--debug", action="store_true")
    p.add_argument("--device", type=int, default=0)
    return p


def main(argv=None):
    from . import io as mio
    from .inference import MigrationInference
    a = build_parser().parse_args(argv)
    units = mio.Units.from_file(a.funits)
    units.PrintUnits()
    rows = None
    if a.fjafs == "":
        sfs = [1 for _ in range(8)]
    else:
        rows = mio.read_jafs(a.fjafs, silent_mode=False).jafs
        sfs = mio.column_sums(rows)
    d = mio.read_ms(a.msstring)
    M = MigrationInference(d.times, d.lambdas, sfs, d.divergenceTime, d.mi, d.pu, unfolded=a.uf, trueEPS=True, device=a.device)
    llh = M.JAFSLikelihood([])
    print("Expected SFS", M.JAFS)
    if rows is not None:
        tot = sum(sfs[1:])
        print("Data     SFS", [v / tot for v in sfs[1:]])
        print("data llh under the model is", llh)
        print("maximum of the llh function is", M.MaximumLLHFunction())
        if a.bsSize > 1:
            rng = random.Random(a.seed)
            M.SetJAFSBatch([sfs] + [mio.bootstrap_jafs(rows, rng) for _ in range(a.bsSize)])
            bs = sorted(float(v) for v in M.JAFSLikelihoodBatch([[]]).reshape(-1)[1:])
            M.SetJAFS(sfs)
            for pct, frac in (("10%", 0.05), ("5%", 0.025)):
                cut = math.ceil(frac * a.bsSize)
                print(pct, "confidence interval", bs[cut], bs[-cut])
    M.CoalescentRates()
    if a.fout != "":
        mio.output_migration(a.fout, [], M, 2 * units.N0)
    return 0


if __name__ == "__main__":
    sys.exit(main())
