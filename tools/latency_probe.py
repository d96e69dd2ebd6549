#!/usr/bin/env python
"""Where the host-side microseconds of one sb_search64 go (resident text): table build, upload enqueue,
launch-to-count wall time against the kernel's own CUDA-event time.
Usage: python tools/latency_probe.py [--gib 4] [--reps 20]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smart_b200 as sb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=4.0)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    n = int(a.gib * (1 << 30))
    with sb.Text.generate(n, 256, 1256) as t:
        for algo in ("bf", "kmp", "hor", "so"):
            for m in (4, 16, 64, 256):
                x = sb.textgen_host(n // 3 + 17, m, 256, 1256)
                rows = []
                for _ in range(a.reps):
                    t0 = time.perf_counter()
                    sb.search(algo, t, x)
                    t1 = time.perf_counter()
                    tm = sb.last_timing()
                    rows.append(((t1 - t0) * 1e3, tm.prep_host_ms, tm.upload_ms, tm.wall_ms, tm.kernel_ms))
                r = np.median(np.array(rows), axis=0)
                print(f"{algo:4s} m={m:4d} call {r[0]*1e3:7.1f} us = prep {r[1]*1e3:6.1f} + upload-enqueue {r[2]*1e3:6.1f} + "
                      f"scan wall {r[3]*1e3:7.1f} (kernel {r[4]*1e3:7.1f}, gap {(r[3]-r[4])*1e3:6.1f}) "
                      f"+ binding {(r[0]-r[1]-r[2]-r[3])*1e3:6.1f}", flush=True)


if __name__ == "__main__":
    main()
