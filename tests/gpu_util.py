"""Shared helpers of the GPU parity tests: run the CUDA path through the C ABI and gather whole-file tables."""
import numpy as np

from biod_b200 import BamReader

REC_FIELDS = ["block_size", "ref_id", "pos", "end_pos", "bin_mq_nl", "flag_nc", "l_seq"]


def gpu_records(data, blocks_per_batch=0, want_offsets=True):
    """Decode every record on the GPU.  Returns (reader, dict of whole-file arrays, list of raw record bytes)."""
    rd = BamReader(data, blocks_per_batch=blocks_per_batch, want_offsets=want_offsets)
    cols = {f: [] for f in REC_FIELDS + ["start_voffset", "end_voffset", "n_cigar_rec"]}
    cigar, raws = [], []
    err = None
    try:
        for b in rd.read_batches(copy=True):
            for f in REC_FIELDS:
                cols[f].append(getattr(b, f))
            cols["start_voffset"].append(b.start_voffset)
            cols["end_voffset"].append(b.end_voffset)
            cols["n_cigar_rec"].append(np.diff(b.cigar_off).astype(np.uint64))
            cigar.append(b.cigar)
            for i in range(b.n):
                o = int(b.rec_off[i]) + 4
                raws.append(b.data[o:o + int(b.block_size[i])].tobytes())
    except Exception as e:  # noqa: BLE001 - the tests look at the class
        err = e
    out = {k: (np.concatenate(v) if v else np.zeros(0)) for k, v in cols.items()}
    out["cigar"] = np.concatenate(cigar) if cigar else np.zeros(0, dtype=np.uint32)
    return rd, out, raws, err


def gpu_pileup(data, single_ref, blocks_per_batch=0, **kw):
    rd = BamReader(data, blocks_per_batch=blocks_per_batch)
    pos, ref, cov, nstart, ridx, base, qual, qoff, refb = [], [], [], [], [], [], [], [], []
    for b in rd.column_batches(single_ref, want_query_offset=True, copy=True, **kw):
        if b.reference_base is not None:
            refb.append(b.reference_base)
        else:
            assert not kw.get("use_md_tag"), "use_md_tag batches must carry reference_base"
        pos.append(b.position)
        ref.append(np.full(b.n_columns, b.ref_id, dtype=np.int32))
        cov.append(np.diff(b.col_off).astype(np.uint64))
        nstart.append(b.n_starting_here)
        ridx.append(b.read_idx)
        base.append(b.base)
        qual.append(b.qual)
        qoff.append(b.query_offset)
    cat = lambda v, dt: np.concatenate(v) if v else np.zeros(0, dtype=dt)  # noqa: E731
    cov = cat(cov, np.uint64)
    return dict(col_pos=cat(pos, np.uint64), col_ref=cat(ref, np.int32), cov=cov,
                col_off=np.concatenate([[0], np.cumsum(cov)]).astype(np.uint64), n_start=cat(nstart, np.uint32),
                read_idx=cat(ridx, np.uint32), base=cat(base, np.uint8), qual=cat(qual, np.uint8),
                qoff=cat(qoff, np.uint32), ref_base=cat(refb, np.uint8))


def assert_pileup_equal(g, o):
    """g: gpu_pileup dict, o: oracle.Pileup"""
    assert len(g["col_pos"]) == o.n_columns, (len(g["col_pos"]), o.n_columns)
    assert np.array_equal(g["col_pos"], o.col_pos)
    assert np.array_equal(g["col_ref"], o.col_ref)
    assert np.array_equal(g["col_off"], o.col_off)
    assert np.array_equal(g["n_start"], o.n_start)
    assert np.array_equal(g["read_idx"], o.read_idx)
    assert np.array_equal(g["base"], o.base)
    assert np.array_equal(g["qual"], o.qual)
    assert np.array_equal(g["qoff"], o.qoff)


def gpu_pileup_sharded(data, n_shards, halo_blocks=8, blocks_per_batch=0, exact=True, spans=None, **kw):
    """Run every shard of a sharded pileup (one after the other on this GPU) and stitch the column tables the way
    biod_b200.stitch does across ranks: concatenate in shard order, rebase read_idx by the record counts.
    exact=True: the shards run the way several GPUs run them — all from the halo_blocks guess, then exact_halos says which
    halos were too short and those shards run again from the exact offset (res["redone"]).  exact=False: the guess only.
    spans=[2, 1, 4]: the shards are run as passes over shards 0-1, 2 and 3-6 (biodb_pileup_begin_shard_span)."""
    from biod_b200.stitch import exact_halos_of_spans
    rd = BamReader(data, blocks_per_batch=blocks_per_batch)

    spans = spans or [1] * n_shards
    assert sum(spans) == n_shards
    firsts = [sum(spans[:k]) for k in range(len(spans))]

    def run(k, halo_voffset=None):
        pos, ref, cov, nstart, ridx, base, qual, qoff, refb = [], [], [], [], [], [], [], [], []
        info = {}
        for b in rd.column_batches(False, want_query_offset=True, copy=True, shard=(firsts[k], n_shards, spans[k]), halo_blocks=halo_blocks,
                                   halo_voffset=halo_voffset, shard_info=info, **kw):
            pos.append(b.position)
            ref.append(np.full(b.n_columns, b.ref_id, dtype=np.int32))
            cov.append(np.diff(b.col_off).astype(np.uint64))
            nstart.append(b.n_starting_here)
            ridx.append(b.read_idx)
            base.append(b.base)
            qual.append(b.qual)
            qoff.append(b.query_offset)
            if b.reference_base is not None:
                refb.append(b.reference_base)
        return (pos, ref, cov, nstart, ridx, base, qual, qoff, refb), info

    parts, infos = [], []
    for k in range(len(spans)):
        part, info = run(k)
        parts.append(part)
        infos.append(info)
    # a span's halo must reach back to the first earlier record that reaches its FIRST shard's columns
    need, redo = exact_halos_of_spans([i["reach"] for i in infos], [i["halo_voffset"] for i in infos], firsts)
    if exact:
        for t in redo:
            parts[t], infos[t] = run(t, need[t])
            assert infos[t]["halo_voffset"] == need[t]
    rec_base = np.concatenate([[0], np.cumsum([i["n_own_records"] for i in infos])]).astype(np.uint64)
    cat = lambda v, dt: np.concatenate(v) if v else np.zeros(0, dtype=dt)  # noqa: E731
    out = {k: [] for k in ("col_pos", "col_ref", "cov", "n_start", "read_idx", "base", "qual", "qoff", "ref_base")}
    for s, (pos, ref, cov, nstart, ridx, base, qual, qoff, refb) in enumerate(parts):
        out["col_pos"] += pos
        out["col_ref"] += ref
        out["cov"] += cov
        out["n_start"] += nstart
        shift = (int(rec_base[s]) - int(infos[s]["n_halo_records"])) % 2**32
        out["read_idx"] += [((r.astype(np.uint64) + np.uint64(shift)) & np.uint64(0xFFFFFFFF)).astype(np.uint32) for r in ridx]
        out["base"] += base
        out["qual"] += qual
        out["qoff"] += qoff
        out["ref_base"] += refb
    dts = dict(col_pos=np.uint64, col_ref=np.int32, cov=np.uint64, n_start=np.uint32, read_idx=np.uint32, base=np.uint8,
               qual=np.uint8, qoff=np.uint32, ref_base=np.uint8)
    res = {k: cat(v, dts[k]) for k, v in out.items()}
    res["col_off"] = np.concatenate([[0], np.cumsum(res["cov"])]).astype(np.uint64)
    res["shards"] = infos
    res["redone"] = redo
    res["halo_ok"] = not redo or exact
    return res
