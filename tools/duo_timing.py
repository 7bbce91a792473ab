"""Phase cycle counters of the two-warp inflate kernel (a -DBIODB_DUO_TIMING build of the library, BIODB_LIB=...):
one device-resident pileup pass over a synthetic file, then the counters as shares of each warp's total."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biod_b200 import _capi as capi  # noqa: E402
from tools import bamgen  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
L = capi.lib()
data = bamgen.generate(n, 1, False, -1, bamgen.SEED_BASE + 2)
o = capi.Options()
L.biodb_default_options(C.byref(o))
o.resident_input = o.device_output = 1
h = C.c_void_p()
assert L.biodb_open_memory(data.ctypes.data, data.size, C.byref(o), C.byref(h)) == 0
cyc = (C.c_uint64 * 16)()
for it in range(2):
    L.biodb_debug_inflate_cycles(cyc, 1)
    it_ = C.c_void_p()
    assert L.biodb_reads_begin(h, C.byref(it_)) == 0
    rb = capi.RecordBatch()
    nrec = 0
    while L.biodb_reads_next(it_, C.byref(rb)) == 0:
        nrec += rb.n
    s = capi.Stats()
    L.biodb_reads_stats(it_, C.byref(s))
    L.biodb_reads_end(it_)
L.biodb_debug_inflate_cycles(cyc, 0)
v = [int(x) for x in cyc]
dn = ["round1", "wait_resolver", "rounds2+", "hdr_tables", "wait_input", "total"]
rn = ["wait_decoder", "replay", "matches", "walk_flush", "total"]
out = {"records": nrec, "blocks": int(s.n_blocks), "inflate_ms": s.inflate_ms,
       "decoder_cycles_per_block": {k: round(v[i] / max(1, s.n_blocks)) for i, k in enumerate(dn)},
       "resolver_cycles_per_block": {k: round(v[8 + i] / max(1, s.n_blocks)) for i, k in enumerate(rn)}}
print(json.dumps(out))
