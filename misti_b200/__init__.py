"""misti_b200 -- B200-native (sm_100a) evaluation of MiSTI's model hot path.

Host-side mirror of the reference's Python interface for the path (MigrationInference,
TwoPopulations, OnePopulation) over hand-written CUDA kernels reached through a C ABI
(include/misti_b200.h, misti_b200/libmisti_b200.so).  There is no CPU fallback: constructing an
Engine without the built library or without a CUDA device raises.
"""
from ._lib import (FLAG_CORRECT, FLAG_CPFIT, FLAG_SMOOTH, FLAG_UNFOLDED, OK, NEGATIVE_PARAM, CORRECTION_FAILED, NONFINITE,
                   INFINITE_COAL_TIME, MistiLibraryError)
from .engine import Engine, default_engine, llh_constants
from .inference import MigrationInference
from .populations import TwoPopulations, OnePopulation

__all__ = ["Engine", "default_engine", "llh_constants", "MigrationInference", "TwoPopulations", "OnePopulation",
           "MistiLibraryError", "FLAG_CORRECT", "FLAG_CPFIT", "FLAG_SMOOTH", "FLAG_UNFOLDED", "OK", "NEGATIVE_PARAM",
           "CORRECTION_FAILED", "NONFINITE", "INFINITE_COAL_TIME"]
