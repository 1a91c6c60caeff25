import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
for _k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
    os.environ.setdefault(_k, "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _load(name):
    with open(os.path.join(HERE, "golden", name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_datasets():
    return _load("datasets.json")["datasets"]


@pytest.fixture(scope="session")
def golden_cases():
    return _load("evals.json")["cases"]


@pytest.fixture(scope="session")
def golden_fits():
    return _load("fits.json")["fits"]


@pytest.fixture(scope="session")
def golden_solver():
    """per-call nfev / status / solution of the reference's least_squares calls (tests/golden/gen_solver_golden.py)"""
    return {c["name"]: c for c in _load("solver.json")["cases"]}


@pytest.fixture(scope="session")
def golden_tables():
    return _load("tables.json")


@pytest.fixture(scope="session")
def hostsim():
    """Test-only host (g++) build of the device numerics headers; see tests/hostsim/hostsim.cpp."""
    import ctypes
    src = os.path.join(HERE, "hostsim", "hostsim.cpp")
    lib = os.path.join(HERE, "hostsim", "libhostsim.so")
    deps = [src] + [os.path.join(ROOT, "misti_b200", "csrc", n) for n in
                    ("misti_math.cuh", "misti_model.cuh", "misti_jsfs.cuh", "misti_tables.h")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-Wno-unknown-pragmas", "-shared", "-fPIC", "-o", lib, src], check=True)
    return ctypes.CDLL(lib)


@pytest.fixture(scope="session")
def engine():
    import misti_b200
    eng = misti_b200.Engine(0)
    yield eng
    eng.close()
