#!/usr/bin/env python
"""Worst-case inputs: periodic texts where EVERY position is a candidate and most are matches."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smart_b200 as sb  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64 << 20
    for name, unit in (("a^n", b"a"), ("(ab)^n", b"ab"), ("(aab)^n", b"aab")):
        y = np.frombuffer((unit * (n // len(unit) + 1))[:n], dtype=np.uint8)
        with sb.Text.upload(y) as t:
            for m in (2, 8, 16, 64, 256, 1024, 4096):
                for pname, x in (("periodic", y[:m].copy()), ("near", np.concatenate([y[:m - 1], np.frombuffer(b"c", np.uint8)]))):
                    for algo in ("bf", "hor"):
                        t0 = time.perf_counter()
                        r = sb.search(algo, t, x)
                        dt = time.perf_counter() - t0
                        tm = sb.last_timing()
                        expect = (n - m) // len(unit) + 1 if pname == "periodic" else 0
                        print(json.dumps({"text": name, "m": m, "pattern": pname, "algo": algo, "count": r.count,
                                          "expected": expect, "ok": r.count == expect, "ms": round(dt * 1e3, 2),
                                          "GBps": round(n / dt / 1e9, 1), "kind": tm.filter_kind,
                                          "verified": tm.verified}), flush=True)


if __name__ == "__main__":
    main()
