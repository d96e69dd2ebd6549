#!/usr/bin/env python
"""Kernel-only GB/s (CUDA events on the launching stream) per (sigma, m, filter kernel) on a resident text.
Usage: python tools/sweep.py [--gib 2] [--sigmas 256,4,2] [--out gpurun_out/sweep.jsonl]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smart_b200 as sb  # noqa: E402

FILTERS = {0: 0, 1: sb.FLAG_FILTER_SHIFT, 2: sb.FLAG_FILTER_B2, 4: sb.FLAG_FILTER_B4, 8: sb.FLAG_FILTER_B8, 16: sb.FLAG_FILTER_B16,
           32: sb.FLAG_FILTER_B32, 408: sb.FLAG_FILTER_B4L8, 416: sb.FLAG_FILTER_B4L16, 816: sb.FLAG_FILTER_B8L16}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=2.0)
    ap.add_argument("--sigmas", default="256,4,2")
    ap.add_argument("--ms", default="1,2,3,4,5,8,9,12,14,15,16,24,31,32,63,64,256,1024,4096")
    ap.add_argument("--algo", default="bf")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="gpurun_out/sweep.jsonl")
    ap.add_argument("--default-only", action="store_true")
    ap.add_argument("--append", action="store_true", help="append to --out instead of replacing it")
    ap.add_argument("--flags", type=int, default=0, help="SB_FLAG_* ORed into every search (8192 = NO_PACK: byte kernels)")
    a = ap.parse_args()
    n = int(a.gib * (1 << 30))
    out = open(a.out, "a" if a.append else "w")
    for sigma in [int(s) for s in a.sigmas.split(",")]:
        seed = 1000 + sigma
        with sb.Text.generate(n, sigma, seed) as t:
            for m in [int(v) for v in a.ms.split(",")]:
                x = sb.textgen_host(n // 3 + 17, m, sigma, seed)
                if a.default_only:
                    fl = [None]
                else:
                    fl = [None] + ([0] if m <= 4 else [1, 2]) + [b for b in (4, 8, 16, 32) if m >= 2 * b - 1]
                    fl += [k for k, need in ((408, 11), (416, 19), (816, 23)) if m >= need]
                for f in fl:
                    flags = a.flags | (0 if f is None else FILTERS[f])
                    ts, cnt = [], None
                    for _ in range(a.reps):
                        r = sb.search(a.algo, t, x, flags=flags)
                        tm = sb.last_timing()
                        ts.append(tm.kernel_ms)
                        cnt = r.count
                    rec = {"algo": a.algo, "sigma": sigma, "m": m, "filter": tm.filter_kind, "default": f is None, "count": cnt, "verified": tm.verified,
                           "kernel_ms_median": float(np.median(ts)), "GBps": n / (float(np.median(ts)) * 1e-3) / 1e9,
                           "GBps_best": n / (min(ts) * 1e-3) / 1e9}
                    out.write(json.dumps(rec) + "\n")
                    out.flush()
                    print(rec, flush=True)


if __name__ == "__main__":
    main()
