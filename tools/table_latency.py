#!/usr/bin/env python
"""Wall time of ONE search on a small resident text, default mode against SB_FLAG_TABLE_SCAN: what the parity tests
and the driver's per-pattern loop pay per call (table build + upload + one-thread-per-4-KiB kernel + read-back)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smart_b200 as sb  # noqa: E402


def main():
    rng = np.random.default_rng(5)
    for tname, y in (("rand256_64k", rng.integers(0, 256, 65536, dtype=np.uint8)),
                     ("rand4_256k", rng.integers(97, 101, 262144, dtype=np.uint8)),
                     ("rand2_64k", rng.integers(48, 50, 65536, dtype=np.uint8))):
        with sb.Text.upload(y) as t:
            for algo in ("kmp", "so", "hor", "bm", "aut", "br", "ebom", "qf34", "wfr", "twfrq4", "zt", "tsw", "kr"):
                for m in (8, 64, 1024):
                    x = y[1000:1000 + m].copy()
                    rec = {"text": tname, "algo": algo, "m": m}
                    for label, flags in (("default", 0), ("table", sb.FLAG_TABLE_SCAN)):
                        sb.search(algo, t, x, flags=flags)
                        t0 = time.perf_counter()
                        for _ in range(5):
                            r = sb.search(algo, t, x, flags=flags)
                        rec[label + "_ms"] = round((time.perf_counter() - t0) / 5 * 1e3, 3)
                        tm = sb.last_timing()
                        rec[label + "_kernel_ms"] = round(tm.kernel_ms, 3)
                        rec[label + "_pre_ms"] = round(tm.pre_ms, 3) if hasattr(tm, "pre_ms") else None
                        rec[label + "_count"] = r.count
                    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
