"""Host <-> device copy bandwidth of this node, per GPU and with all ranks copying AT THE SAME TIME (the ceiling the
end-to-end numbers of bench.py live under).  One process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_bw.py
or `python tools/pcie_bw.py` for one GPU.  Prints one JSON line (rank 0): GB/s per rank and summed, device->host alone,
host->device alone, and both directions at once — pinned host buffers of 1 GiB, CUDA events, barrier before every leg."""
import json
import os

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", str(rank)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def leg(pairs, reps=6):
    """pairs: [(dst, src, stream)] copied together, reps times; GB/s of each pair on this rank."""
    for dst, src, st in pairs:
        with torch.cuda.stream(st):
            dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in pairs]
    for (dst, src, st), (a, b) in zip(pairs, ev):
        with torch.cuda.stream(st):
            a.record()
            for _ in range(reps):
                dst.copy_(src, non_blocking=True)
            b.record()
    torch.cuda.synchronize()
    return [reps * n / (a.elapsed_time(b) * 1e-3) / 1e9 for a, b in ev]


res = {"d2h": leg([(h, d, s1)])[0], "h2d": leg([(d, h, s1)])[0]}
both = leg([(h, d, s1), (d2, h2, s2)])
res["bidir_d2h"], res["bidir_h2d"] = both
t = torch.tensor([res["d2h"], res["h2d"], res["bidir_d2h"], res["bidir_h2d"]], dtype=torch.float64, device="cuda")
if world > 1:
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    table = torch.stack(allt).cpu()
else:
    table = t.cpu().unsqueeze(0)
if rank == 0:
    names = ["d2h", "h2d", "bidir_d2h", "bidir_h2d"]
    print(json.dumps({"n_gpus": world, "unit": "GB/s", "all ranks copying at the same time": True,
                      "per_rank": {k: [round(float(x), 1) for x in table[:, i]] for i, k in enumerate(names)},
                      "sum": {k: round(float(table[:, i].sum()), 1) for i, k in enumerate(names)}}))
if world > 1:
    dist.destroy_process_group()
