#!/usr/bin/env python
"""Aggregate host->device copy bandwidth of this box with one process per GPU copying at the same time
(plain torch pinned tensor -> cuda tensor, nothing of ours): the ceiling for bench.py's `e2e` at N GPUs.
Launch: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_probe.py"""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << 30
    src = torch.empty(n, dtype=torch.uint8).pin_memory()
    src.fill_(rank + 1)
    dst = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    res = {}
    for mode in ("alone", "together"):
        if world > 1:
            dist.barrier()
        if mode == "alone" and world > 1:
            # ranks take turns
            mine = None
            for r in range(world):
                if r == rank:
                    t0 = time.perf_counter()
                    for _ in range(8):
                        dst.copy_(src, non_blocking=True)
                    torch.cuda.synchronize()
                    mine = 8 * n / (time.perf_counter() - t0) / 1e9
                dist.barrier()
            gbps = mine
        else:
            t0 = time.perf_counter()
            for _ in range(8):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            gbps = 8 * n / (time.perf_counter() - t0) / 1e9
        t = torch.tensor([gbps], device="cuda")
        if world > 1:
            lst = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(lst, t)
            res[mode] = [round(float(x), 1) for x in lst]
        else:
            res[mode] = [round(gbps, 1)]
    if rank == 0:
        res["sum_together"] = round(sum(res["together"]), 1)
        try:
            res["cpus"] = os.cpu_count()
            res["numa_nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
        except OSError:
            pass
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
