"""The C-ABI library loads and exports every symbol include/misti_b200.h declares; the ctypes mirror of
its structs has the C layout.  No compute calls (CPU only)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "misti_b200.h")


@pytest.fixture(scope="module")
def lib():
    from misti_b200 import _build, _lib
    _build.build()
    return _lib.load()


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(misti_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from misti_b200 import _lib
    names = _declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), "libmisti_b200.so does not export %s" % n
        assert n in _lib.SIGNATURES, "ctypes binding misses %s" % n
    assert sorted(_lib.SIGNATURES) == names
    assert lib.misti_abi_version() == _lib.ABI_VERSION == 2


def test_struct_layout_matches_c(tmp_path):
    from misti_b200 import _lib
    prog = tmp_path / "layout.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "misti_b200.h"\n'
                    'int main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(misti_model_desc), offsetof(misti_model_desc, band_pop),'
                    'offsetof(misti_model_desc, pulse_pop), offsetof(misti_model_desc, band_val), offsetof(misti_model_desc, pulse_val),'
                    'sizeof(misti_eval_io));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(prog)], check=True)
    out = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    M = _lib.ModelDesc
    assert out == [ctypes.sizeof(M), M.band_pop.offset, M.pulse_pop.offset, M.band_val.offset, M.pulse_val.offset,
                   ctypes.sizeof(_lib.EvalIO)]


def test_constants_agree_with_header():
    from misti_b200 import _lib
    src = open(HEADER).read()
    for name, val in (("MISTI_FLAG_CORRECT", _lib.FLAG_CORRECT), ("MISTI_FLAG_CPFIT", _lib.FLAG_CPFIT),
                      ("MISTI_FLAG_SMOOTH", _lib.FLAG_SMOOTH), ("MISTI_FLAG_UNFOLDED", _lib.FLAG_UNFOLDED),
                      ("MISTI_FLAG_DEVICE_PTRS", _lib.FLAG_DEVICE_PTRS), ("MISTI_STIFF", _lib.STIFF),
                      ("MISTI_MAX_BANDS", _lib.MAX_BANDS), ("MISTI_MAX_PARAMS", _lib.MAX_PARAMS)):
        m = re.search(r"#define %s \(?(-?\d+)u?\)?" % name, src)
        assert m and int(m.group(1)) == val, name


def test_no_cpu_fallback(lib):
    """without a CUDA device the context cannot be created and the Python layer raises (never computes on the host)"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    assert lib.misti_ctx_create(0, None, ctypes.byref(h)) == -3
    import misti_b200
    with pytest.raises(misti_b200.MistiLibraryError):
        misti_b200.Engine(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "misti_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "hostsim" not in text.replace(
                    "tests/hostsim", ""), f
