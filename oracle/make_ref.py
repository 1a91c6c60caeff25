#!/usr/bin/env python3
"""Recipe for oracle/_ref/ -- the UNMODIFIED reference, staged so that it can be timed next to the CUDA path.

TEST / BENCH INFRASTRUCTURE ONLY.  Genomics-HSE/MiSTI is pure Python (no build step), so "building" the reference is
packing the modules the path imports -- MigrationInference.py, TwoPopulations.py, OnePopulation.py, CorrectLambda.py,
migrationIO.py (+ psmc.py, which migrationIO imports, MiSTI.py for the command line and setunits.txt) -- byte for byte,
from where they lie under /root/reference, into ONE binary artefact, oracle/_ref/misti_reference.zip (Python imports
straight from a zip archive), in the git-ignored directory oracle/_ref/ (never into history: the directory is listed in
.gitignore but not in .gpurunignore, so it travels to the GPU box like a built .so).  A manifest with the sha256 of every
packed file is written beside it.  `bench.py --impl reference` and bench.py's `cpu_baseline`
leg time THIS code when the directory is present (kind "reference") and the oracle port otherwise (kind "port").
oracle/ref_loader.py imports it (numpy.mat alias for NumPy 2, BLAS threads pinned to 1 as MiSTI.py:23-25).

    python oracle/make_ref.py            # /root/reference -> oracle/_ref/
"""
import hashlib
import json
import os
import zipfile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("MISTI_REFERENCE_DIR", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["MigrationInference.py", "TwoPopulations.py", "OnePopulation.py", "CorrectLambda.py", "migrationIO.py", "psmc.py",
         "MiSTI.py", "setunits.txt", "LICENSE"]


def main():
    if not os.path.isdir(SRC):
        print("make_ref: %s is not present (GPU box): keeping oracle/_ref/ as it is" % SRC)
        return 0 if os.path.isdir(DST) else 1
    os.makedirs(DST, exist_ok=True)
    manifest = {"source": "Genomics-HSE/MiSTI, unmodified, staged by oracle/make_ref.py", "files": {}}
    with zipfile.ZipFile(os.path.join(DST, "misti_reference.zip"), "w", zipfile.ZIP_DEFLATED) as z:
        for name in FILES:
            src = os.path.join(SRC, name)
            if not os.path.exists(src):
                continue
            with open(src, "rb") as f:
                blob = f.read()
            z.writestr(zipfile.ZipInfo(name, date_time=(2020, 1, 1, 0, 0, 0)), blob)  # fixed stamp: reproducible archive
            manifest["files"][name] = hashlib.sha256(blob).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("make_ref: packed %d files into %s/misti_reference.zip" % (len(manifest["files"]), DST))
    return 0


if __name__ == "__main__":
    sys.exit(main())
