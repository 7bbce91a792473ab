"""world_size-2 gloo test of the only cross-rank step of the path: the column-table stitch."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from biod_b200.stitch import stitch_counts
    from conftest import fixture_bytes
    from oracle import oracle as orc
    # each rank owns one reference of ex1_header.bam (shards never mix references, splitter.d:83-85)
    b = orc.Bam(fixture_bytes("ex1_header.bam")).decode()
    p = b.pileup_columns()
    sel = p.col_ref == rank
    cov = np.diff(p.col_off)[sel]
    n_rec = int((b.ref_id == rank).sum())
    r = stitch_counts(int(sel.sum()), int(cov.sum()), n_rec)
    out.put((rank, r))
    dist.destroy_process_group()


def test_stitch_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # pins of test/unittests.d:334-367: 1470 chr1 columns then 1567 chr2 columns
    assert res[0]["per_rank"] == res[1]["per_rank"]
    assert [x[0] for x in res[0]["per_rank"]] == [1470, 1567]
    assert res[0]["col_base"] == 0 and res[1]["col_base"] == 1470
    assert res[1]["ent_base"] == res[0]["per_rank"][0][1]
    assert res[0]["totals"][0] == 3037
    assert res[0]["totals"][2] <= 3270


def test_stitch_single_process():
    sys.path.insert(0, ROOT)
    from biod_b200.stitch import stitch_counts
    r = stitch_counts(10, 300, 7)
    assert r["world"] == 1 and r["col_base"] == 0 and r["totals"] == (10, 300, 7)
