#!/usr/bin/env python
"""A broad, SMALL pass over every kernel family for `compute-sanitizer --tool memcheck` (one tool per gpurun call):
texts of 70 KiB - 1 MiB, every filter kind, the slow paths (dense candidates, periodic runs), positions, the packed
copies, the table-driven scans, the host entries.  Prints OK when every count equals the oracle's."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib  # noqa: E402
import smart_b200 as sb  # noqa: E402


def main():
    ora = oracle_lib.Oracle()
    rng = np.random.default_rng(1)
    n = 300_001
    per = np.resize(np.frombuffer(b"ab", np.uint8), n).copy()
    per[-1] = ord("c")
    texts = {"s256": sb.textgen_host(0, n, 256, 1), "s4": sb.textgen_host(0, n, 4, 2), "s2": sb.textgen_host(0, 70_001, 2, 3),
             "s16": sb.textgen_host(0, n, 16, 4), "ab": per, "a": np.full(80_000, 97, np.uint8)}
    checked = 0
    for name, y in texts.items():
        with sb.Text.upload(y) as t, sb.Text.upload(y, devices=[0, 0, 0]) as t3:
            for m in (1, 2, 3, 4, 5, 8, 9, 12, 15, 16, 19, 23, 24, 31, 32, 47, 63, 64, 100, 200, 300, 1024, 4200):
                if m >= y.size:
                    continue
                for k in (0, y.size - m, y.size // 2):
                    x = y[k:k + m].copy()
                    want = ora.bf_positions(x, y)
                    for flags in (0, sb.FLAG_NO_PACK, sb.FLAG_NO_PACK | sb.FLAG_NO_SKIP, sb.FLAG_NO_PERIOD):
                        for a in ("bf", "hor"):
                            assert sb.search(a, t, x, flags=flags).count == want.size, (name, m, k, a, flags)
                            checked += 1
                    assert sb.search("kmp", t3, x).count == want.size
                    r = sb.search("bm", t3, x, flags=sb.FLAG_POSITIONS, positions_cap=want.size + 1)
                    assert np.array_equal(r.positions, want), (name, m, k)
                    if m <= 300 or name in ("s256", "s4"):
                        for a in ("kmp", "so", "hor", "bm", "qs", "raita", "tunedbm", "kr", "hash3", "qf34", "aut", "bom", "ebom"):
                            if sb.applicable(a, m):
                                assert sb.search(a, t, x, flags=sb.FLAG_TABLE_SCAN).count == want.size, (name, m, a)
                    x2 = x.copy()
                    x2[-1] ^= 1  # near miss
                    assert sb.search("bf", t, x2).count == ora.bf_count(x2, y)
                    assert sb.search("hor", t, x2, flags=sb.FLAG_NO_PACK).count == ora.bf_count(x2, y)
            tickets = [sb.submit("bf", t, y[i:i + 9].copy()) for i in range(0, 300, 7)]
            got = [sb.collect(tk).count for tk in tickets]
            assert got == [ora.bf_count(y[i:i + 9], y) for i in range(0, 300, 7)]
            pats = np.stack([y[o:o + 16] for o in rng.integers(0, y.size - 16, 50)])
            counts, _, _ = sb.search_multi("bf", t, pats)
            assert counts.tolist() == [ora.bf_count(p, y) for p in pats]
        jobs = [("bf", y[10:18].copy()), ("kmp", y[100:140].copy()), ("hor", y[5:305].copy())]
        assert [r.count for r in sb.search_host_batch(jobs, y)] == [ora.bf_count(x, y) for _, x in jobs]
    with sb.Text.generate_rand(2, 100000) as t:
        assert np.array_equal(t.download(0, 100000), sb.textgen_rand_host(2, 0, 100000))
    print("OK", checked)


if __name__ == "__main__":
    main()
