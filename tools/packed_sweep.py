#!/usr/bin/env python
"""Text rate of the 2-bit packed scan against the byte kernels on a sigma <= 4 text (kernel-only, CUDA events).
Usage: python tools/packed_sweep.py [--gib 4] [--sigma 4] [--out gpurun_out/packed_sweep.jsonl]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smart_b200 as sb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=4.0)
    ap.add_argument("--sigma", type=int, default=4)  # 2 -> 1-bit, 3..4 -> 2-bit, 5..16 -> 4-bit copy
    ap.add_argument("--ms", default="15,16,20,23,24,30,31,40,47,64,79,100,128,143,160,200,256,400,512,1024")
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--out", default="gpurun_out/packed_sweep.jsonl")
    a = ap.parse_args()
    n = int(a.gib * (1 << 30))
    out = open(a.out, "w")
    with sb.Text.generate(n, a.sigma, 1000 + a.sigma) as t:
        assert t.packed_bits in (1, 2, 4)
        for m in [int(v) for v in a.ms.split(",")]:
            x = sb.textgen_host(n // 3 + 17, m, a.sigma, 1000 + a.sigma)
            for algo in ("bf", "hor"):
                row = {"sigma": a.sigma, "m": m, "algo": algo}
                for label, flags in (("packed", 0), ("bytes", sb.FLAG_NO_PACK)):
                    ts = []
                    for _ in range(a.reps):
                        r = sb.search(algo, t, x, flags=flags)
                        tm = sb.last_timing()
                        ts.append(tm.kernel_ms)
                    row[label] = {"kind": tm.filter_kind, "count": r.count, "verified": tm.verified,
                                  "text_GBps": n / (float(np.median(ts)) * 1e-3) / 1e9}
                assert row["packed"]["count"] == row["bytes"]["count"]
                out.write(json.dumps(row) + "\n")
                out.flush()
                print(m, algo, row["packed"]["kind"], round(row["packed"]["text_GBps"]), "|", row["bytes"]["kind"],
                      round(row["bytes"]["text_GBps"]), "| count", row["packed"]["count"], flush=True)


if __name__ == "__main__":
    main()
