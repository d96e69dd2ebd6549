#!/usr/bin/env python
"""BASELINE.json configs[3]: a 64 GiB sigma=4 text sharded with an m-1 halo across 1/2/4/8 B200 (in-process
partitioner, one stream per device), per-shard counts summed on the host.  Strong scaling: same text, more GPUs.
Usage (8-GPU box): python tools/config4.py [--gib 64] > gpurun_out/config4.jsonl"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smart_b200 as sb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=64.0)
    ap.add_argument("--sigma", type=int, default=4)
    a = ap.parse_args()
    n = int(a.gib * (1 << 30))
    seed = 64
    ndev = sb.device_count()
    ms = (8, 64, 1024)
    pats = {m: sb.textgen_host(n // 2 - m // 2, m, a.sigma, seed) for m in ms}  # straddles a shard seam for even G
    ref = {}
    for G in (1, 2, 4, 8):
        if G > ndev:
            break
        t0 = time.perf_counter()
        text = sb.Text.generate(n, a.sigma, seed, devices=list(range(G)))
        gen_s = time.perf_counter() - t0
        for algo in ("bf", "hor"):
            for m in ms:
                sb.search(algo, text, pats[m])
                ts, ks = [], []
                for _ in range(5):
                    t1 = time.perf_counter()
                    r = sb.search(algo, text, pats[m])
                    ts.append(time.perf_counter() - t1)
                    ks.append(sb.last_timing().kernel_ms)
                key = (algo, m)
                ref.setdefault(key, r.count)
                rec = {"gpus": G, "algo": algo, "m": m, "count": r.count, "same_as_1gpu": r.count == ref[key],
                       "wall_ms": float(np.median(ts)) * 1e3, "kernel_ms_max_over_devices": float(np.median(ks)),
                       "text_GBps": n / float(np.median(ts)) / 1e9, "kind": sb.last_timing().filter_kind,
                       "text_generation_s": round(gen_s, 2)}
                print(json.dumps(rec), flush=True)
        text.free()


if __name__ == "__main__":
    main()
