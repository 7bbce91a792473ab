"""ctypes binding of libbiod_b200.so (include/biod_b200.h).  Fails loudly when the CUDA library is
missing: there is no CPU fallback anywhere in this package."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BIODB_LIB") or os.path.join(_HERE, "libbiod_b200.so")   # BIODB_LIB: kernel-variant experiments

OK, EOF = 0, 1
ERR_BGZF, ERR_ZLIB, ERR_FORMAT, ERR_TRUNCATED, ERR_IO, ERR_CUDA, ERR_CIGAR, ERR_UNSORTED, ERR_ARG, ERR_NOMEM = \
    -1, -2, -3, -4, -5, -6, -7, -8, -9, -10

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)
u64p = C.POINTER(C.c_uint64)


class Error(C.Structure):
    _fields_ = [("status", C.c_int32), ("zlib_errnum", C.c_int32), ("file_offset", C.c_uint64),
                ("message", C.c_char * 256)]


class Options(C.Structure):
    _fields_ = [("device", C.c_int32), ("blocks_per_batch", C.c_int32), ("verify_crc", C.c_int32),
                ("want_offsets", C.c_int32), ("pin_input", C.c_int32), ("resident_input", C.c_int32),
                ("device_output", C.c_int32), ("reserved", C.c_int32 * 1)]


class RecordBatch(C.Structure):
    _fields_ = [("n", C.c_uint64), ("first_index", C.c_uint64), ("data", u8p), ("data_len", C.c_uint64),
                ("rec_off", u64p), ("block_size", i32p), ("ref_id", i32p), ("pos", i32p), ("end_pos", i32p),
                ("bin_mq_nl", u32p), ("flag_nc", u32p), ("l_seq", i32p), ("cigar_off", u64p), ("cigar", u32p),
                ("start_voffset", u64p), ("end_voffset", u64p)]


class Stats(C.Structure):
    _fields_ = [("total_ms", C.c_double), ("inflate_ms", C.c_double), ("scan_ms", C.c_double), ("pileup_ms", C.c_double),
                ("inflate_launches", C.c_uint64), ("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64), ("compressed_bytes", C.c_uint64), ("uncompressed_bytes", C.c_uint64),
                ("n_blocks", C.c_uint64), ("n_records", C.c_uint64)]


class PileupParams(C.Structure):
    _fields_ = [("single_ref", C.c_int32), ("skip_zero_coverage", C.c_int32), ("use_md_tag", C.c_int32),
                ("want_query_offset", C.c_int32), ("start_from", C.c_uint64), ("end_at", C.c_uint64),
                ("counts_only", C.c_int32), ("compact_reads", C.c_int32), ("maq_mode", C.c_int32), ("reserved", C.c_int32 * 1)]


class MaqParams(C.Structure):
    _fields_ = [("depcorr", C.c_float), ("eta", C.c_float), ("minimum_call_quality", C.c_float),
                ("minimum_base_quality", C.c_int32)]


class ShardInfo(C.Structure):
    _fields_ = [("first_voffset", C.c_uint64), ("end_voffset", C.c_uint64), ("halo_voffset", C.c_uint64),
                ("lo_ref", C.c_int32), ("hi_ref", C.c_int32), ("lo_pos", C.c_int64), ("hi_pos", C.c_int64),
                ("n_halo_records", C.c_uint64), ("n_own_records", C.c_uint64)]


class ColumnBatch(C.Structure):
    _fields_ = [("n_columns", C.c_uint64), ("n_entries", C.c_uint64), ("ref_id", C.c_int32),
                ("last_of_pileup", C.c_int32), ("position", u64p), ("col_off", u64p), ("n_starting_here", u32p),
                ("read_idx", u32p), ("base", u8p), ("qual", u8p), ("query_offset", u32p), ("counts", u32p),
                ("last_read", u32p), ("live_mask", u64p), ("n_stragglers", C.c_uint64), ("strag_col", u32p), ("strag_idx", u32p),
                ("n_runs", C.c_uint64), ("run_pos", u64p), ("run_first_col", u32p),
                ("base4", u8p), ("n_special", C.c_uint64), ("special_entry", u32p), ("special_base", u8p),
                ("reference_base", u8p),
                ("n_calls", C.c_uint64), ("call_col", u32p), ("call_pos", u64p), ("call_gt", u8p), ("call_ref", u8p),
                ("call_qual", C.POINTER(C.c_float)),
                ("maq_gt0", u8p), ("maq_gt1", u8p), ("maq_s0", C.POINTER(C.c_float)), ("maq_s1", C.POINTER(C.c_float)),
                ("maq_n_valid", C.POINTER(C.c_uint16))]


_lib = None


def lib():
    """Load libbiod_b200.so.  Raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with biod_b200/csrc/build.sh — biod_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.biodb_version.restype = C.c_char_p
    L.biodb_default_options.argtypes = [C.POINTER(Options)]
    L.biodb_open.restype = C.c_int
    L.biodb_open.argtypes = [C.c_char_p, C.POINTER(Options), C.POINTER(vp)]
    L.biodb_open_memory.restype = C.c_int
    L.biodb_open_memory.argtypes = [vp, C.c_size_t, C.POINTER(Options), C.POINTER(vp)]
    L.biodb_close.argtypes = [vp]
    L.biodb_last_error.restype = C.POINTER(Error)
    L.biodb_last_error.argtypes = [vp]
    L.biodb_open_error.restype = C.POINTER(Error)
    L.biodb_header_text.restype = C.c_int
    L.biodb_header_text.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t)]
    L.biodb_n_refs.restype = C.c_int32
    L.biodb_n_refs.argtypes = [vp]
    L.biodb_ref_info.restype = C.c_int
    L.biodb_ref_info.argtypes = [vp, C.c_int32, C.POINTER(C.c_char_p), i32p, i32p]
    L.biodb_reads_start_voffset.restype = C.c_uint64
    L.biodb_reads_start_voffset.argtypes = [vp]
    L.biodb_file_size.restype = C.c_uint64
    L.biodb_file_size.argtypes = [vp]
    L.biodb_reads_begin.restype = C.c_int
    L.biodb_reads_begin.argtypes = [vp, C.POINTER(vp)]
    L.biodb_reads_next.restype = C.c_int
    L.biodb_reads_next.argtypes = [vp, C.POINTER(RecordBatch)]
    L.biodb_reads_end.argtypes = [vp]
    L.biodb_reads_progress.restype = C.c_float
    L.biodb_reads_progress.argtypes = [vp]
    L.biodb_pileup_begin.restype = C.c_int
    L.biodb_pileup_begin.argtypes = [vp, C.POINTER(PileupParams), C.POINTER(vp)]
    L.biodb_pileup_begin_shard.restype = C.c_int
    L.biodb_pileup_begin_shard.argtypes = [vp, C.POINTER(PileupParams), C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(vp)]
    L.biodb_pileup_shard_info.argtypes = [vp, C.POINTER(ShardInfo)]
    L.biodb_pileup_begin_shard_at.restype = C.c_int
    L.biodb_pileup_begin_shard_at.argtypes = [vp, C.POINTER(PileupParams), C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(vp)]
    L.biodb_pileup_begin_shard_span.restype = C.c_int
    L.biodb_pileup_begin_shard_span.argtypes = [vp, C.POINTER(PileupParams), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(vp)]
    L.biodb_pileup_begin_shard_span_at.restype = C.c_int
    L.biodb_pileup_begin_shard_span_at.argtypes = [vp, C.POINTER(PileupParams), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(vp)]
    L.biodb_pileup_shard_reach.argtypes = [vp, u64p]
    L.biodb_shard_cuts.restype = C.c_int
    L.biodb_shard_cuts.argtypes = [vp, C.c_uint32, u64p, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
    L.biodb_pileup_begin_range.restype = C.c_int
    L.biodb_pileup_begin_range.argtypes = [vp, C.POINTER(PileupParams), C.c_uint64, C.c_uint64, C.c_int32, C.c_int64, C.c_int32,
                                           C.c_int64, C.POINTER(vp)]
    L.biodb_pileup_maq_params.restype = C.c_int
    L.biodb_pileup_maq_params.argtypes = [vp, C.POINTER(MaqParams)]
    L.biodb_pileup_next.restype = C.c_int
    L.biodb_pileup_next.argtypes = [vp, C.POINTER(ColumnBatch)]
    L.biodb_pileup_end.argtypes = [vp]
    L.biodb_pileup_ref_id.restype = C.c_int32
    L.biodb_pileup_ref_id.argtypes = [vp]
    L.biodb_pileup_totals.argtypes = [vp, u64p, u64p, u64p]
    L.biodb_reads_stats.argtypes = [vp, C.POINTER(Stats)]
    L.biodb_pileup_stats.argtypes = [vp, C.POINTER(Stats)]
    L.biodb_dev_inflate.restype = C.c_int
    L.biodb_dev_inflate.argtypes = [vp, vp, vp, vp, vp, C.c_uint32, vp, vp, vp, vp]
    L.biodb_input_is_pinned.restype = C.c_int32
    L.biodb_input_is_pinned.argtypes = [vp]
    L.biodb_debug_inflate_counters.restype = C.c_int
    L.biodb_debug_inflate_counters.argtypes = [u64p, C.c_int32]
    L.biodb_debug_md_chain.restype = C.c_int64
    L.biodb_debug_md_chain.argtypes = [vp, vp, vp, vp, C.c_uint64, C.c_int32, C.c_uint64, vp, C.c_uint64]
    L.biodb_index_open.restype = C.c_int
    L.biodb_index_open.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.biodb_index_close.argtypes = [vp]
    L.biodb_index_n_refs.restype = C.c_int32
    L.biodb_index_n_refs.argtypes = [vp]
    L.biodb_reads_begin_regions.restype = C.c_int
    L.biodb_reads_begin_regions.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp, vp, C.POINTER(vp)]
    L.biodb_index_regions_chunks.restype = C.c_int64
    L.biodb_index_regions_chunks.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp, vp, C.c_uint64]
    L.biodb_index_chunks.restype = C.c_int64
    L.biodb_index_chunks.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, C.c_uint64]
    L.biodb_index_last_linear_offset.restype = C.c_int32
    L.biodb_index_last_linear_offset.argtypes = [vp, C.c_int32, u64p]
    L.biodb_index_builder_begin.restype = C.c_int
    L.biodb_index_builder_begin.argtypes = [C.c_int32, C.c_int32, C.POINTER(vp)]
    L.biodb_index_builder_put.restype = C.c_int
    L.biodb_index_builder_put.argtypes = [vp, C.c_uint64, vp, vp, vp, vp, vp, vp, vp]
    L.biodb_index_builder_finish.restype = C.c_int
    L.biodb_index_builder_finish.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.biodb_index_builder_error.restype = C.c_char_p
    L.biodb_index_builder_error.argtypes = [vp]
    L.biodb_index_builder_end.argtypes = [vp]
    L.biodb_reads_begin_region.restype = C.c_int
    L.biodb_reads_begin_region.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(vp)]
    L.biodb_pileup_begin_region.restype = C.c_int
    L.biodb_pileup_begin_region.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(PileupParams), C.POINTER(vp)]
    L.biodb_reads_begin_between.restype = C.c_int
    L.biodb_reads_begin_between.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(vp)]
    L.biodb_bgzf_compress_bound.restype = C.c_size_t
    L.biodb_bgzf_compress_bound.argtypes = [C.c_size_t]
    L.biodb_bgzf_compress.restype = C.c_int
    L.biodb_bgzf_compress.argtypes = [C.c_int32, vp, C.c_size_t, C.c_int32, C.c_int32, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.biodb_writer_begin.restype = C.c_int
    L.biodb_writer_begin.argtypes = [C.c_int32, C.c_int32, C.POINTER(vp)]
    L.biodb_writer_header.restype = C.c_int
    L.biodb_writer_header.argtypes = [vp, C.c_char_p, C.c_size_t, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_int32)]
    L.biodb_writer_records.restype = C.c_int
    L.biodb_writer_records.argtypes = [vp, vp, C.c_size_t]
    L.biodb_writer_flush.restype = C.c_int
    L.biodb_writer_flush.argtypes = [vp]
    L.biodb_writer_drain.restype = C.c_int
    L.biodb_writer_drain.argtypes = [vp, C.c_uint32, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.biodb_writer_finish.restype = C.c_int
    L.biodb_writer_finish.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.biodb_writer_layout.restype = C.c_int
    L.biodb_writer_layout.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.biodb_writer_index.restype = C.c_int
    L.biodb_writer_index.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.biodb_writer_debug_set_output.restype = C.c_int
    L.biodb_writer_debug_set_output.argtypes = [vp, vp, C.c_size_t]
    L.biodb_writer_error.restype = C.c_char_p
    L.biodb_writer_error.argtypes = [vp]
    L.biodb_writer_end.argtypes = [vp]
    L.biodb_debug_deflate_block.restype = C.c_int64
    L.biodb_debug_deflate_block.argtypes = [vp, C.c_uint32, vp, C.c_uint32, C.c_int32]
    L.biodb_debug_deflate_stats.restype = C.c_int
    L.biodb_debug_deflate_stats.argtypes = [C.c_int32, vp]
    L.biodb_debug_md_dna.restype = C.c_int64
    L.biodb_debug_md_dna.argtypes = [vp, C.c_int64, vp, C.c_uint64]
    L.biodb_dev_scan_workspace_bytes.restype = C.c_size_t
    L.biodb_dev_scan_workspace_bytes.argtypes = [C.c_uint32]
    L.biodb_dev_scan_records.restype = C.c_int
    L.biodb_dev_scan_records.argtypes = [vp, C.c_uint64, vp, C.c_uint32, C.c_int32, vp, vp, vp, C.c_size_t, vp]
    _lib = L
    return L


EXPORTS = [
    "biodb_version", "biodb_default_options", "biodb_open", "biodb_open_memory", "biodb_close", "biodb_last_error",
    "biodb_open_error", "biodb_header_text", "biodb_n_refs", "biodb_ref_info", "biodb_reads_start_voffset",
    "biodb_file_size", "biodb_input_is_pinned", "biodb_reads_begin", "biodb_reads_next", "biodb_reads_end", "biodb_reads_progress",
    "biodb_pileup_begin", "biodb_pileup_next", "biodb_pileup_end", "biodb_pileup_ref_id", "biodb_pileup_totals",
    "biodb_reads_stats", "biodb_pileup_stats", "biodb_pileup_begin_shard", "biodb_pileup_begin_shard_at", "biodb_pileup_begin_shard_span", "biodb_pileup_begin_shard_span_at", "biodb_pileup_shard_info",
    "biodb_pileup_shard_reach", "biodb_shard_cuts", "biodb_pileup_begin_range", "biodb_pileup_maq_params",
    "biodb_dev_inflate", "biodb_dev_scan_records", "biodb_dev_scan_workspace_bytes", "biodb_debug_inflate_counters", "biodb_debug_md_chain",
    "biodb_debug_md_dna", "biodb_index_open", "biodb_index_close", "biodb_index_n_refs", "biodb_index_chunks", "biodb_index_regions_chunks", "biodb_reads_begin_regions", "biodb_index_last_linear_offset", "biodb_index_builder_begin", "biodb_index_builder_put",
    "biodb_index_builder_finish", "biodb_index_builder_error", "biodb_index_builder_end",
    "biodb_reads_begin_region", "biodb_reads_begin_between", "biodb_pileup_begin_region", "biodb_bgzf_compress_bound", "biodb_bgzf_compress",
    "biodb_debug_deflate_block", "biodb_debug_deflate_stats", "biodb_writer_begin", "biodb_writer_header", "biodb_writer_records", "biodb_writer_flush",
    "biodb_writer_drain", "biodb_writer_finish", "biodb_writer_layout", "biodb_writer_index", "biodb_writer_debug_set_output", "biodb_writer_error", "biodb_writer_end",
]
