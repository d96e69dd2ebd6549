#!/usr/bin/env python
"""Filter choice on a skewed (English-letter-frequency, i.i.d.) text: default vs forced filters."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smart_b200 as sb  # noqa: E402

FREQ = {" ": 18.3, "e": 10.2, "t": 7.5, "a": 6.5, "o": 6.2, "n": 5.7, "i": 5.7, "s": 5.3, "r": 5.0, "h": 5.0,
        "l": 3.3, "d": 3.3, "u": 2.3, "c": 2.2, "m": 2.0, "f": 2.0, "w": 1.7, "g": 1.6, "p": 1.5, "y": 1.4,
        "b": 1.3, "v": 0.8, "k": 0.6, ",": 1.0, ".": 0.9, "\n": 1.2, "x": 0.15, "j": 0.1, "q": 0.08, "z": 0.07}


def main():
    n = 256 << 20
    rng = np.random.default_rng(1)
    sym = np.frombuffer("".join(FREQ).encode(), dtype=np.uint8)
    p = np.array(list(FREQ.values()))
    y = sym[rng.choice(len(sym), size=n, p=p / p.sum())]
    # a text-like pattern source: also try the most frequent symbols only
    with sb.Text.upload(y) as t:
        print("entropy estimate bits/byte:", round(t.entropy_bits, 2))
        for m in (5, 8, 12, 16, 32, 64, 256, 1024):
            pats = {"cut": y[12345678:12345678 + m].copy(), "common": np.frombuffer((" the " * 300)[:m].encode(), np.uint8).copy()}
            for pname, x in pats.items():
                row = {"m": m, "pattern": pname}
                for label, fl in (("default", 0), ("win8", sb.FLAG_FILTER_SHIFT), ("B4", sb.FLAG_FILTER_B4),
                                  ("B8", sb.FLAG_FILTER_B8), ("B16", sb.FLAG_FILTER_B16)):
                    for algo in ("bf", "hor"):
                        ts = []
                        for _ in range(3):
                            r = sb.search(algo, t, x, flags=fl)
                            tm = sb.last_timing()
                            ts.append(tm.kernel_ms)
                        row[f"{label}/{algo}"] = f"f{tm.filter_kind}:{n / (min(ts) * 1e-3) / 1e9:.0f}:v{tm.verified}"
                print(json.dumps(row))


if __name__ == "__main__":
    main()
