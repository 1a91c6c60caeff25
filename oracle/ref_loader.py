"""Loader of the packed, unmodified reference oracle/_ref/misti_reference.zip (see oracle/make_ref.py).

TEST / BENCH INFRASTRUCTURE ONLY: imported by bench.py's CPU legs and by tests; never by the product package.
The reference does `from numpy import mat` (MigrationInference.py:25, TwoPopulations.py:25, OnePopulation.py:24), removed
in NumPy 2: `numpy.mat` is aliased to `numpy.asmatrix` before the import.  BLAS threads are pinned to 1 like
MiSTI.py:23-25 does (a 3x3 expm costs 1.7 ms instead of 3.5 us otherwise)."""
import contextlib
import io
import os
import sys
import warnings

for _k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
    os.environ.setdefault(_k, "1")

REF_ZIP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "misti_reference.zip")


def available():
    return os.path.exists(REF_ZIP)


def load():
    """The reference's MigrationInference class (imported on first call)."""
    if not available():
        raise RuntimeError("oracle/_ref/misti_reference.zip is missing: run `python oracle/make_ref.py` in the build container")
    import numpy
    if not hasattr(numpy, "mat"):
        numpy.mat = numpy.asmatrix
    if REF_ZIP not in sys.path:
        sys.path.insert(0, REF_ZIP)  # zipimport: the modules are read from the archive
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    import MigrationInference
    return MigrationInference.MigrationInference


def quiet(fn, *a, **k):
    """the reference prints from its constructor and its objective: keep that off the bench's output"""
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        return fn(*a, **k)


def make_model(times, lambdas, sfs, splitT, mi=(), pu=(), cpfit=False, smooth=True, unfolded=True, trueEPS=False, sampleDate=0):
    """MigrationInference object of the reference, built the way MiSTI.py builds it (band / pulse values as strings)."""
    MI = load()
    return quiet(MI, list(times), [list(v) for v in lambdas], list(sfs), splitT, [list(map(str, m)) for m in mi],
                 [list(map(str, p)) for p in pu], smooth=smooth, unfolded=unfolded, trueEPS=trueEPS, cpfit=cpfit,
                 sampleDate=sampleDate, mixtureTH=0.0)
