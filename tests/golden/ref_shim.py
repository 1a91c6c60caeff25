"""Import shim for the UNMODIFIED reference at /root/reference (build container only).

Test infrastructure.  The reference does ``from numpy import mat`` (MigrationInference.py:25,
TwoPopulations.py:25, OnePopulation.py:24); ``numpy.mat`` was removed in NumPy 2, so we alias it
to ``numpy.asmatrix`` before importing.  BLAS threads are pinned to 1 like MiSTI.py:23-25.
/root/reference does not exist on the GPU box: nothing outside the fixture generators may
import this module.
"""
import os
for _k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
    os.environ.setdefault(_k, "1")
import sys
import warnings
import numpy

REFERENCE_DIR = os.environ.get("MISTI_REFERENCE_DIR", "/root/reference")


def load():
    """Return the reference modules as a dict (imports them on first call)."""
    if not os.path.isdir(REFERENCE_DIR):
        raise RuntimeError("reference tree %s not present (only exists in the build container)" % REFERENCE_DIR)
    if not hasattr(numpy, "mat"):
        numpy.mat = numpy.asmatrix
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    import migrationIO
    import MigrationInference
    import TwoPopulations
    import OnePopulation
    import CorrectLambda
    return dict(migrationIO=migrationIO,
                MigrationInference=MigrationInference.MigrationInference,
                TwoPopulations=TwoPopulations.TwoPopulations,
                OnePopulation=OnePopulation.OnePopulation,
                CorrectLambda=CorrectLambda.CorrectLambda)


def read_jafs(migrationIO, fn):
    """ReadJAFS without the mutable-default accumulation hazard (migrationIO.py:39)."""
    migrationIO.JAFS.__init__.__defaults__[0].clear()
    return migrationIO.ReadJAFS(fn, True)


def reset_units(migrationIO):
    u = migrationIO.Units
    u.mutRate, u.binsize, u.N0, u.genTime, u.hetloss1, u.hetloss2 = 1.25e-8, 100, 10000, 1, 0.0, 0.0
