"""CPU-side checks of the boundary: the library loads and exports every symbol the header declares, and the
product fails loudly — never falls back — when no CUDA device is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    import __graft_entry__ as g
    g.build()
    from biod_b200 import _capi
    assert os.path.exists(_capi.LIB_PATH)
    return _capi.LIB_PATH


def test_every_declared_symbol_is_exported(lib_path):
    hdr = open(os.path.join(ROOT, "include", "biod_b200.h")).read()
    declared = sorted(set(re.findall(r"^(?:const\s+)?[a-z0-9_]+\s*\*?\s*(biodb_[a-z0-9_]+)\(", hdr, re.M)))
    assert len(declared) >= 20
    L = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(L, name), name
    from biod_b200 import _capi
    assert sorted(_capi.EXPORTS) == declared


def test_no_cpu_fallback_without_gpu(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from biod_b200 import BamReader, CudaUnavailable
    from conftest import fixture_bytes
    with pytest.raises(CudaUnavailable):
        BamReader(fixture_bytes("ex1_header.bam"))


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "biod_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".sh")):
                src = open(os.path.join(dp, fn), errors="ignore").read()
                assert "oracle" not in src.lower(), os.path.join(dp, fn)
                assert "zlib.h" not in src and "-lz" not in src, os.path.join(dp, fn)
