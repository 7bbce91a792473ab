"""Column-table stitch across shards: every rank learns every shard's column / entry counts and
derives the global offsets of its own shard — an all-gather of three integers per rank (NCCL on GPUs, gloo in
the CPU tests).  Shards are otherwise independent: no data-path collective exists on this hot path."""
import torch
import torch.distributed as dist


def stitch_counts(n_columns, n_entries, n_records, device="cpu"):
    """Returns dict(rank, world, col_base, ent_base, rec_base, totals=(cols, entries, records), per_rank=[...])."""
    mine = torch.tensor([n_columns, n_entries, n_records], dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        world, rank = dist.get_world_size(), dist.get_rank()
        allc = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allc, mine)
        table = torch.stack(allc).cpu()
    else:
        world, rank = 1, 0
        table = mine.cpu().unsqueeze(0)
    base = torch.cumsum(table, 0) - table          # exclusive scan over ranks
    tot = table.sum(0)
    return dict(rank=rank, world=world, col_base=int(base[rank, 0]), ent_base=int(base[rank, 1]),
                rec_base=int(base[rank, 2]), totals=tuple(int(x) for x in tot),
                per_rank=[tuple(int(x) for x in row) for row in table])


def halo_sufficient(shards):
    """Exactness check of a sharded pileup (see biodb_shard_info): `shards` = list of shard-info dicts in shard order.
    True iff no read outside a shard's halo reaches into its column range."""
    for s in range(1, len(shards)):
        lo_ref, lo_pos = shards[s]["lo_ref"], shards[s]["lo_pos"]
        for q in range(s):
            if shards[q]["hi_ref"] != lo_ref:
                continue
            if shards[s]["halo_coffset"] <= shards[q]["first_coffset"]:
                continue                      # the halo re-reads all of shard q
            m = shards[q]["max_end_outside_tail"] if q == s - 1 else shards[q]["max_end_all"]
            if m > lo_pos:
                return False
    return True
