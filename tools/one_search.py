#!/usr/bin/env python
"""ONE search on a synthetic text, for ncu captures of a single kernel.
Usage: python tools/one_search.py --text a|ab|aab|rand --sigma 4 --mib 256 --m 64 --pattern periodic|near|swap|mid|cut --algo bf [--flags N] [--reps 3]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smart_b200 as sb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--text", default="rand")
    ap.add_argument("--sigma", type=int, default=256)
    ap.add_argument("--mib", type=int, default=256)
    ap.add_argument("--m", type=int, default=64)
    ap.add_argument("--pattern", default="cut")
    ap.add_argument("--algo", default="bf")
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    n = a.mib << 20
    if a.text == "rand":
        t = sb.Text.generate(n, a.sigma, 7)
        x = sb.textgen_host(n // 3, a.m, a.sigma, 7)
    else:
        unit = a.text.encode()
        y = np.frombuffer((unit * (n // len(unit) + 1))[:n], dtype=np.uint8).copy()
        y[-1] = ord("c")
        t = sb.Text.upload(y)
        x = y[:a.m].copy()
        if a.pattern == "near":
            x[-1] = ord("c")
        elif a.pattern == "swap":
            x[-1] = ord("b") if x[-1] == ord("a") else ord("a")
        elif a.pattern == "mid":
            x[a.m // 2] = ord("c")
    for _ in range(a.reps):
        r = sb.search(a.algo, t, x, flags=a.flags)
        tm = sb.last_timing()
        print(r.count, tm.filter_kind, round(tm.kernel_ms, 4), "ms", round(n / tm.kernel_ms / 1e6, 1), "GB/s", tm.verified)
    t.free()


if __name__ == "__main__":
    main()
