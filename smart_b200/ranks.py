"""One-process-per-GPU plumbing for bench.py: each rank owns one part of the global text (the same
partition rule the in-process multi-device partitioner uses, sb_shard_range) and the per-part counts are
summed.  The data path has NO collective: text shards trivially with an m-1 halo;
torch.distributed is used only for the barrier, the max-over-ranks timing and the 8-byte count sum."""
from __future__ import annotations

import os
from dataclasses import dataclass


@dataclass
class RankInfo:
    rank: int
    world: int
    local_rank: int


def rank_info() -> RankInfo:
    return RankInfo(int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
                    int(os.environ.get("LOCAL_RANK", "0")))


def init_process_group(backend: str):
    """Join the torchrun rendezvous (MASTER_ADDR/MASTER_PORT from the environment). No-op for world 1."""
    info = rank_info()
    if info.world == 1:
        return None
    import torch.distributed as dist

    if not dist.is_initialized():
        dist.init_process_group(backend=backend, rank=info.rank, world_size=info.world)
    return dist


def reduce_step(dist, device, count: int, elapsed_ms: float, kernel_ms: float):
    """(sum of counts, max elapsed, max kernel time) over all ranks; identity when dist is None."""
    if dist is None:
        return count, elapsed_ms, kernel_ms
    import torch

    c = torch.tensor([count], dtype=torch.int64, device=device)
    t = torch.tensor([elapsed_ms, kernel_ms], dtype=torch.float64, device=device)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(c.item()), float(t[0].item()), float(t[1].item())


def barrier(dist, device=None):
    if dist is None:
        return
    import torch

    if device is not None and str(device).startswith("cuda"):
        dist.barrier(device_ids=[torch.device(device).index])
    else:
        dist.barrier()


def cpu_group(dist):
    """A gloo group next to the NCCL one: its barrier blocks on the HOST.  Used around the legs where only rank 0
    works (in-process multi-device run, CPU baseline): an NCCL barrier would keep a kernel spinning on every other
    GPU meanwhile."""
    if dist is None:
        return None
    return dist.new_group(backend="gloo")


def host_barrier(group):
    if group is None:
        return
    import torch.distributed as dist

    dist.barrier(group=group)
