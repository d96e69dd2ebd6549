#!/usr/bin/env python
"""Worst-case inputs: periodic texts where EVERY position is a candidate and most are matches (periodic patterns),
or every position passes the block fingerprints and fails late (near-periodic patterns).  Each text carries one
foreign byte 'c' at its end, so the near patterns stay inside the text's alphabet (otherwise they are answered
without a scan).  Rates are text bytes per second of the scan kernel(s) (CUDA events) and of the whole call.
Usage: python tools/adversarial.py [n_bytes] [flags]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smart_b200 as sb  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256 << 20
    flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    for name, unit in (("a^n", b"a"), ("(ab)^n", b"ab"), ("(aab)^n", b"aab")):
        y = np.frombuffer((unit * (n // len(unit) + 1))[:n], dtype=np.uint8).copy()
        y[-1] = ord("c")
        with sb.Text.upload(y) as t:
            for m in (2, 8, 16, 24, 40, 64, 256, 1024, 4096):
                per = y[:m].copy()
                near = per.copy()
                near[-1] = ord("c")                      # foreign last byte
                swap = per.copy()
                swap[-1] = ord("b") if per[-1] == ord("a") else ord("a")  # in-alphabet last byte (a^n: 'b' is foreign)
                mid = per.copy()
                mid[m // 2] = ord("c")
                for pname, x in (("periodic", per), ("near", near), ("swap", swap), ("mid", mid)):
                    for algo in ("bf", "hor"):
                        sb.search(algo, t, x, flags=flags)
                        t0 = time.perf_counter()
                        r = sb.search(algo, t, x, flags=flags)
                        dt = time.perf_counter() - t0
                        tm = sb.last_timing()
                        nocc = (n - 1 - m) // len(unit) + 1
                        expect = nocc if pname == "periodic" else None
                        print(json.dumps({"text": name, "m": m, "pattern": pname, "algo": algo, "count": r.count,
                                          "expected": expect, "ok": expect is None or r.count == expect,
                                          "kernel_ms": round(tm.kernel_ms, 3), "wall_ms": round(dt * 1e3, 3),
                                          "GBps_kernel": round(n / max(tm.kernel_ms, 1e-6) / 1e6, 1),
                                          "GBps_wall": round(n / dt / 1e9, 1), "kind": tm.filter_kind,
                                          "verified": tm.verified}), flush=True)


if __name__ == "__main__":
    main()
