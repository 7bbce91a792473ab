"""biod_b200 — B200-native BGZF inflate -> BAM record/CIGAR decode -> pileup.

Everything is computed by libbiod_b200.so (hand-written sm_100a CUDA, C ABI in include/biod_b200.h).
Importing the package does not need a GPU; opening a reader does, and fails loudly without one.
"""
from . import _capi  # noqa: F401
from .bam import (BaiFile, BamFormatException, BamWriter, IndexBuilder, createIndex, BgzfOutputStream, bgzf_compress, BamRead, BamReader, BgzfException, CudaUnavailable, PileupColumn,  # noqa: F401
                  PileupException, ReadException, ZlibException, makePileup, pileupColumns, pileupChunks, PileupChunk, chunk_plan, MaqSnpCaller, DiploidCall5)
