"""The reference harness's own fixtures (src/test.c:252-382): the 16 deterministic (pattern, text, count)
cases of its 20; cases 12-15 are time-seeded random there and are replaced by seeded equivalents
(pattern = prefix of a random sigma-128 text, count checked against brute force)."""
import numpy as np

TEST_C_FIXTURES = [
    # (pattern, text, expected count)            test.c case
    (b"a", b"a" * 10, 10),                       # 1
    (b"aa", b"a" * 10, 9),                       # 2
    (b"a" * 10, b"a" * 10, 1),                   # 3
    (b"b", b"a" * 10, 0),                        # 4
    (b"ab", b"ab" * 5, 5),                       # 5
    (b"a", b"ab" * 5, 5),                        # 6
    (b"aba", b"ab" * 5, 4),                      # 7
    (b"abc", b"ab" * 5, 0),                      # 8
    (b"ba", b"ab" * 5, 4),                       # 9
    (b"babbbbb", b"ab" * 5, 0),                  # 10
    (b"bcdefg", b"bcdefghilm", 1),               # 11
    (b"a" * 40, b"a" * 64, 25),                  # 16
    (b"ab" * 20, b"ab" * 32, 13),                # 17
    (b"ab" * 19 + b"ac", b"ab" * 32, 0),         # 18
    (b"babbbbb", b"abababbbbb", 1),              # 19
    (b"bababb", b"abababbbbb", 1),               # 20
]


def seeded_random_fixtures():
    """test.c cases 12-15 (:324-353) with a fixed seed: text of n random bytes < 128, pattern = its prefix."""
    rng = np.random.default_rng(12345)
    out = []
    for n, m in ((10, 4), (10, 4), (64, 40), (64, 40)):
        y = rng.integers(0, 128, n, dtype=np.uint8)
        out.append((y[:m].tobytes(), y.tobytes()))
    return out
