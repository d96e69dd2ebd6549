#!/usr/bin/env python
"""Multi-pattern scan: effective text GB/s (npat x n bytes / time) vs one scan per pattern."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smart_b200 as sb  # noqa: E402


def main():
    n = 4 << 30
    out = []
    for sigma in (256, 4):
        with sb.Text.generate(n, sigma, 2024) as t:
            for m, npat in ((8, 500), (16, 500), (32, 500), (64, 500), (256, 500), (16, 50), (16, 2000)):
                offs = np.random.default_rng(m).integers(0, n - m, npat)
                pats = np.stack([sb.textgen_host(int(o), m, sigma, 2024) for o in offs])
                sb.search_multi("bf", t, pats)
                t0 = time.perf_counter()
                counts, pre, run = sb.search_multi("bf", t, pats)
                wall = time.perf_counter() - t0
                tm = sb.last_timing()
                t1 = time.perf_counter()
                single = [sb.search("bf", t, p).count for p in pats[:20]]
                per_single = (time.perf_counter() - t1) / 20
                assert list(counts[:20]) == single
                rec = {"sigma": sigma, "m": m, "npat": npat, "kind": tm.filter_kind, "multi_ms": wall * 1e3,
                       "kernel_ms": tm.kernel_ms, "single_ms_per_pattern": per_single * 1e3,
                       "effective_TBps": npat * n / wall / 1e12, "speedup_vs_single": per_single * npat / wall,
                       "verified": tm.verified, "min_count": int(counts.min())}
                print(json.dumps(rec), flush=True)
                out.append(rec)
    json.dump(out, open("gpurun_out/multi_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
