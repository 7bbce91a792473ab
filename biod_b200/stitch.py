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


NONE = 2**64 - 1


def exact_halos(reach_rows, halo_used):
    """The exact halo check of a sharded pileup (include/biod_b200.h, biodb_pileup_shard_reach).
    reach_rows[j][t] = virtual offset of shard j's first own record that reaches into shard t's columns (NONE if none);
    halo_used[t] = where shard t's halo started.  Returns (need, redo): need[t] = the exact start of shard t's halo
    (NONE when no earlier read reaches it), redo = the shards whose halo started behind that and must be run again with
    biodb_pileup_begin_shard_at(need[t])."""
    n = len(halo_used)
    need = [min([reach_rows[j][t] for j in range(t)], default=NONE) for t in range(n)]
    redo = [t for t in range(n) if need[t] < halo_used[t]]
    return need, redo


def exact_halos_of_spans(reach_rows, halo_used, firsts):
    """exact_halos for workers that each run a SPAN of consecutive shards (biodb_pileup_begin_shard_span): worker j runs
    the shards from firsts[j] on, its reach row is indexed by shard, and its halo must reach back to the first earlier
    record that reaches the columns of its FIRST shard.  Returns (need, redo) per worker."""
    return exact_halos([[row[a] for a in firsts] for row in reach_rows], halo_used)


def gather_reach(reach_row, halo_voffset, device="cpu"):
    """All-gather of every rank's reach row and halo start (one row of world + 1 integers per rank); returns
    (reach_rows, halo_used) for exact_halos.  Offsets travel as int64 (NONE = -1)."""
    enc = lambda v: -1 if v >= NONE else int(v)  # noqa: E731
    dec = lambda v: NONE if v < 0 else int(v)  # noqa: E731
    mine = torch.tensor([enc(v) for v in reach_row] + [enc(halo_voffset)], dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        allv = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
        dist.all_gather(allv, mine)
        table = torch.stack(allv).cpu().tolist()
    else:
        table = [mine.cpu().tolist()]
    return [[dec(v) for v in row[:-1]] for row in table], [dec(row[-1]) for row in table]
