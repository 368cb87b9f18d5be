"""Multi-GPU sharding of an extension batch: jobs are independent, so a batch is cut into
contiguous ranges balanced by DP cells (qlen*tlen), one range per rank/GPU; there is no exchange step.  The only
communication is the gather of the 24-byte results to the caller (host side, for SAM emission).

`extend_sharded` runs inside a torch.distributed process group (NCCL on a GPU box, gloo on CPU in the tests): every
rank computes its own range with `compute(cfg, jobs, qpool, tpool) -> RES_DT[]` (on a GPU box:
KswB200(local_rank).extend_batch) and the results are all-gathered in caller order.
"""
from __future__ import annotations

import numpy as np

from .ksw import RES_DT


def shard_ranges(qlen: np.ndarray, tlen: np.ndarray, world: int):
    """Contiguous [begin, end) per rank with (nearly) equal sums of qlen*tlen; ranks may get empty ranges."""
    n = int(len(qlen))
    cost = qlen.astype(np.int64) * np.maximum(tlen.astype(np.int64), 1)
    csum = np.concatenate([[0], np.cumsum(cost)])
    total = int(csum[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r // world
        cuts.append(int(np.searchsorted(csum, target, side="left")))
    cuts.append(n)
    cuts = [min(max(c, 0), n) for c in cuts]
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def extend_sharded(compute, cfg, jobs: np.ndarray, qpool: np.ndarray, tpool: np.ndarray, group=None) -> np.ndarray:
    """Every rank passes the same batch; returns the full result array on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    ranges = shard_ranges(jobs["qlen"], jobs["tlen"], world)
    b, e = ranges[rank]
    mine = compute(cfg, jobs[b:e], qpool, tpool) if e > b else np.zeros(0, dtype=RES_DT)
    assert mine.dtype == RES_DT and mine.shape[0] == e - b
    # all_gather of equally sized int32 buffers (padded to the largest shard)
    cap = max(x[1] - x[0] for x in ranges)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    buf = torch.zeros((cap, 6), dtype=torch.int32, device=dev)
    if e > b:
        buf[: e - b] = torch.from_numpy(np.ascontiguousarray(mine).view(np.int32).reshape(-1, 6)).to(dev)
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    res = np.zeros(jobs.shape[0], dtype=RES_DT)
    for r, (rb, re_) in enumerate(ranges):
        if re_ > rb:
            res[rb:re_] = outs[r][: re_ - rb].cpu().numpy().reshape(-1).view(RES_DT)
    return res
