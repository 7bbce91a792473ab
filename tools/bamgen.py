"""ctypes wrapper of the synthetic BAM generator (tools/bamgen.cpp).  Input generation only."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_L = None

# seeds per BASELINE.json config index (seed = 0xB10D + config_index)
SEED_BASE = 0xB10D


def lib():
    global _L
    if _L is None:
        so = os.path.join(_HERE, "libbamgen.so")
        src = os.path.join(_HERE, "bamgen.cpp")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
        _L = C.CDLL(so)
        _L.bamgen_bound.restype = C.c_uint64
        _L.bamgen_bound.argtypes = [C.c_uint64, C.c_int]
        _L.bamgen_generate.restype = C.c_uint64
        _L.bamgen_generate.argtypes = [C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_void_p,
                                       C.c_uint64, C.POINTER(C.c_uint64)]
    return _L


def generate(n_reads, n_refs=1, mixed=False, level=-1, seed=SEED_BASE + 2, threads=None, out=None, straddle=False):
    """Returns a numpy uint8 array holding a complete BAM file (a view of `out` when given)."""
    L = lib()
    threads = threads or os.cpu_count() or 1
    cap = int(L.bamgen_bound(n_reads, int(mixed)))
    if out is None:
        out = np.empty(cap, dtype=np.uint8)
    rl = C.c_uint64()
    n = L.bamgen_generate(n_reads, n_refs, int(mixed) | (2 if straddle else 0), level, seed, threads, out.ctypes.data,
                          out.size, C.byref(rl))
    if n == 0:
        raise MemoryError("bamgen: output buffer too small")
    return out[:int(n)]
