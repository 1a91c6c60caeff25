#!/usr/bin/env python3
"""Golden vectors for the `.mi` result file (migrationIO.OutputMigration / ReadMigration, migrationIO.py:346-505): the
reference evaluates BASELINE config 2 (split 40, -uf, -mi 2 5 12 0.8 1 --cpfit) at m = 0.8 and config 4 (ancient sample)
on the synthetic data, writes the file, and reads it back; stored are the file's text and what the reference parsed from it.
Run in the build container only (needs /root/reference through ref_shim); writes mi.json next to this script."""
import contextlib
import io
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402


def main():
    R = ref_shim.load()
    mio = R["migrationIO"]
    with open(os.path.join(HERE, "datasets.json")) as f:
        dss = json.load(f)["datasets"]
    runs = [("config2_cpfit", "synthetic", 40, [["2", "5", "12", "0.8", "1"]], [0.8], dict(cpfit=True)),
            ("config4_ancient", "synthetic_ancient", 45, [], [], dict())]
    out = []
    with tempfile.TemporaryDirectory() as tmp:
        for name, dsn, st, mi, mu, kw in runs:
            ds = dss[dsn]
            fn = os.path.join(tmp, name + ".mi")
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                M = R["MigrationInference"]([v for v in ds["times"]], [list(v) for v in ds["lambdas"]], list(ds["sfs"]), st, mi, [],
                                            thrh=[ds["theta"], ds["rho"]], enableOutput=False, smooth=True, unfolded=True,
                                            sampleDate=ds.get("sampleDate", 0), **kw)
                M.JAFSLikelihood(mu)
                mio.OutputMigration(fn, mu, M, ds["scaleTime"], 1)
                d = mio.ReadMigration(fn)
            with open(fn) as f:
                text = f.read()
            out.append({"name": name, "text": text, "llh": d.llh, "splitT": d.splitT, "sampleDate": d.sampleDate, "thrh": d.thrh,
                        "jaf": [float(v) for v in d.jaf], "times": d.times, "lambda1": d.lambda1, "lambda2": d.lambda2,
                        "lambdah1": d.lambdah1, "lambdah2": d.lambdah2})
    with open(os.path.join(HERE, "mi.json"), "w") as f:
        json.dump({"meta": {"generator": "tests/golden/gen_mi_golden.py"}, "cases": out}, f, indent=1)
    for c in out:
        print(c["name"], c["llh"], c["splitT"], c["sampleDate"], len(c["times"]), len(c["text"]))


if __name__ == "__main__":
    main()
