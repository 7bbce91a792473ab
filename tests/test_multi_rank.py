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


def _reach_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from biod_b200.stitch import NONE, exact_halos, gather_reach
    # rank 0's read at voffset 0x5000 reaches into shard 1, whose halo was guessed at 0x9000: too late
    reach = [NONE, 0x5000] if rank == 0 else [NONE, NONE]
    rows, used = gather_reach(reach, 0 if rank == 0 else 0x9000)
    out.put((rank, exact_halos(rows, used)))
    dist.destroy_process_group()


def test_exact_halo_exchange_two_ranks():
    """The second collective of the path (include/biod_b200.h, biodb_pileup_shard_reach): every rank learns where the
    halo of every shard must start, and which shards have to be run again."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30500 + os.getpid() % 1000
    procs = [ctx.Process(target=_reach_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == res[1]
    need, redo = res[0]
    assert need[1] == 0x5000 and redo == [1]


def test_exact_halos_single_process():
    sys.path.insert(0, ROOT)
    from biod_b200.stitch import NONE, exact_halos, gather_reach
    rows, used = gather_reach([NONE, NONE, NONE], 0)
    assert rows == [[NONE, NONE, NONE]] and used == [0]
    # three shards: shard 0 reaches shards 1 and 2, shard 1 reaches shard 2 from further on; shard 2's guess was too short
    need, redo = exact_halos([[NONE, 100, 120], [NONE, NONE, 300], [NONE, NONE, NONE]], [0, 90, 250])
    assert need == [NONE, 100, 120] and redo == [2]


def test_exact_halos_of_spans():
    """Workers that run spans of shards (unequal shares on common cuts): the halo a worker needs is what earlier workers
    report for the FIRST shard of its span."""
    sys.path.insert(0, ROOT)
    from biod_b200.stitch import NONE, exact_halos_of_spans
    # 6 shards; three workers run the shards [0, 1], [2] and [3, 4, 5].  Worker 0's own records reach into shards 2
    # (from offset 700) and 3 (from 900); worker 1's into shard 3 (from 1500) and 4.
    rows = [[NONE, NONE, 700, 900, NONE, NONE],
            [NONE, NONE, NONE, 1500, 1600, NONE],
            [NONE] * 6]
    need, redo = exact_halos_of_spans(rows, [0, 650, 1400], [0, 2, 3])
    assert need == [NONE, 700, 900]              # worker 2 needs the earlier of 900 and 1500
    assert redo == [2]                           # its halo started at 1400, behind 900; worker 1 started early enough
    need, redo = exact_halos_of_spans(rows, [0, 650, 880], [0, 2, 3])
    assert redo == []
