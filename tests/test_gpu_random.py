"""Randomised parity (hypothesis): arbitrary small texts and patterns over tiny alphabets — the regime where
borders, periods, overlaps and text-end effects are dense — through every algorithm and both entry points."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import smart_b200 as sb
from golden_texts import ALGOS, expected

pytestmark = pytest.mark.gpu




@st.composite
def text_and_pattern(draw):
    sigma = draw(st.sampled_from([1, 2, 3, 4, 16, 256]))
    n = draw(st.integers(1, 3000))
    seed = draw(st.integers(0, 2**31 - 1))
    rng = np.random.default_rng(seed)
    if draw(st.booleans()):   # periodic text
        period = rng.integers(0, sigma, draw(st.integers(1, 9)), dtype=np.uint8)
        y = np.resize(period, n)
    else:
        y = rng.integers(0, sigma, n, dtype=np.uint8)
    m = draw(st.integers(1, min(n, 300)))
    if draw(st.booleans()):   # pattern cut from the text (>= 1 occurrence)
        k = draw(st.integers(0, n - m))
        x = y[k:k + m].copy()
        if draw(st.booleans()) and m > 1:
            x[draw(st.integers(0, m - 1))] ^= 1
    else:
        x = rng.integers(0, sigma, m, dtype=np.uint8)
    return y, x


@settings(max_examples=400, deadline=None, suppress_health_check=list(HealthCheck))
@given(text_and_pattern())
def test_random_small_cases(oracle, case):
    y, x = case
    want = oracle.bf_positions(x, y)
    with sb.Text.upload(y) as t, sb.Text.upload(y, devices=[0, 0]) as t2:
        for a in ALGOS:
            assert sb.search(a, t, x).count == expected(a, x.size, want.size, y.size), (a, x.size, y.size)
        r = sb.search("kmp", t2, x, flags=sb.FLAG_POSITIONS, positions_cap=max(1, y.size))
        assert np.array_equal(r.positions, want)
        for flags in (sb.FLAG_DENSE, sb.FLAG_FILTER_SHIFT if x.size >= 5 else 0, sb.FLAG_NO_SKIP, sb.FLAG_NO_PERIOD):
            assert sb.search("hor", t, x, flags=flags).count == want.size
        # the algorithms' own table-driven scans (kmpNext, hbc, bmGs, masks, shift tables, QF bits, the automaton)
        for a in ("kmp", "mp", "hor", "bm", "qs", "raita", "tunedbm", "so", "sa", "kr", "hash3", "hash5", "hash8", "qf23",
                  "qf34", "qf62", "aut", "bom", "ebom", "bndm", "br", "tvsbs", "smith", "skip", "fjs", "sbndm2",
                  "ssabs", "fs", "graspm", "tbm", "ac", "zt", "skip3", "skip6", "wfr", "wfrq2", "wfrq4",
                  "bndmq2", "bndmq4", "fndm", "sbndm", "ts", "kmpskip", "bom2", "sebom", "iom", "qlqs", "tsw", "tvsbs-w2",
                  "wfr3", "twfr", "twfr5", "twfrq2", "bxs", "bxs2", "bxs4", "sbndm-bmh", "lbndm", "bsdm", "bsdm4", "bsdm7", "aoso2", "aoso6", "svm0", "wc", "fs-w1", "fbom", "sfbom", "wom", "pbmh"):
            assert sb.search(a, t2, x, flags=sb.FLAG_TABLE_SCAN).count == expected(a, x.size, want.size, y.size), (a, x.size)
        # asynchronous tickets give the same answers as the blocking calls
        tickets = [sb.submit(a, t, x) for a in ("bf", "bm", "so", "qf23", "wfr", "sbndm2")]
        assert [sb.collect(tk).count for tk in tickets] == [expected(a, x.size, want.size, y.size)
                                                             for a in ("bf", "bm", "so", "qf23", "wfr", "sbndm2")]
    assert sb.search_host("bm", x, y).count == want.size


@settings(max_examples=25, deadline=None, suppress_health_check=list(HealthCheck))
@given(st.integers(0, 2**31 - 1), st.sampled_from([2, 4, 256]), st.sampled_from([7, 16, 40, 130, 260]))
def test_random_pattern_sets(oracle, seed, sigma, m):
    rng = np.random.default_rng(seed)
    y = rng.integers(0, sigma, 60000, dtype=np.uint8)
    npat = int(rng.integers(1, 80))
    pats = np.stack([y[o:o + m] for o in rng.integers(0, y.size - m, npat)])
    if npat > 2:
        pats[1] = rng.integers(0, sigma, m, dtype=np.uint8)
    with sb.Text.upload(y) as t:
        counts, _, _ = sb.search_multi("bf", t, pats)
        assert counts.tolist() == [oracle.bf_count(p, y) for p in pats]


@st.composite
def packed_text_and_pattern(draw):
    """texts of at least 64 KiB over 2..16 byte values (they also get a packed copy), every pattern shape"""
    nsym = draw(st.sampled_from([2, 3, 4, 5, 9, 16]))
    seed = draw(st.integers(0, 2**31 - 1))
    rng = np.random.default_rng(seed)
    alphabet = rng.choice(256, nsym, replace=False).astype(np.uint8)
    n = draw(st.integers(65536, 300000))
    kind = draw(st.sampled_from(["uniform", "skewed", "periodic", "blocks"]))
    if kind == "uniform":
        idx = rng.integers(0, nsym, n)
    elif kind == "skewed":
        p = rng.dirichlet(np.full(nsym, 0.4))
        idx = rng.choice(nsym, n, p=p)
    elif kind == "periodic":
        idx = np.resize(rng.integers(0, nsym, draw(st.integers(1, 700))), n)
    else:  # long runs with noise
        idx = np.repeat(rng.integers(0, nsym, n // 50 + 1), 50)[:n]
        noise = rng.integers(0, n, n // 200)
        idx[noise] = rng.integers(0, nsym, noise.size)
    y = alphabet[idx]
    m = draw(st.one_of(st.integers(1, 40), st.integers(7, 1300)))  # the short ones: scan_pswar on the 1-/2-bit copy
    k = draw(st.sampled_from([0, n - m, int(rng.integers(0, n - m))]))
    x = y[k:k + m].copy()
    mut = draw(st.sampled_from(["none", "last", "first", "middle", "foreign"]))
    if mut == "last":
        x[-1] = alphabet[(int(np.where(alphabet == x[-1])[0][0]) + 1) % nsym]
    elif mut == "first":
        x[0] = alphabet[(int(np.where(alphabet == x[0])[0][0]) + 1) % nsym]
    elif mut == "middle":
        x[m // 2] = alphabet[(int(np.where(alphabet == x[m // 2])[0][0]) + 1) % nsym]
    elif mut == "foreign":
        x[m // 3] = next(b for b in range(256) if b not in set(alphabet.tolist()))
    return y, x


@settings(max_examples=200, deadline=None, suppress_health_check=list(HealthCheck))
@given(packed_text_and_pattern())
def test_random_packed_texts(oracle, case):
    """The packed-copy kernels (and whatever the planner picks instead) against the oracle: counts for a prefix-
    and a suffix-anchored family, sharded and unsharded, and the sorted positions."""
    y, x = case
    want = oracle.bf_positions(x, y)
    with sb.Text.upload(y) as t, sb.Text.upload(y, devices=[0, 0, 0]) as t3:
        assert t.packed_bits in (1, 2, 4)
        for a in ("bf", "hor", "kmp", "wfr", "sbndm"):
            assert sb.search(a, t, x).count == expected(a, x.size, want.size, y.size), (a, x.size, y.size, sb.last_timing().filter_kind)
        assert sb.search("qs", t3, x).count == want.size
        assert sb.search("bf", t, x, flags=sb.FLAG_NO_PACK).count == want.size
        r = sb.search("so", t3, x, flags=sb.FLAG_POSITIONS, positions_cap=want.size + 1)
        assert np.array_equal(r.positions, want)
        if want.size > 1:  # a short buffer gets the smallest positions, the count stays complete
            r = sb.search("bf", t, x, flags=sb.FLAG_POSITIONS, positions_cap=want.size // 2)
            assert r.count == want.size and np.array_equal(r.positions, want[: want.size // 2])
        assert [r.count for r in sb.search_host_batch([("bf", x), ("hor", x)], y)] == [want.size, want.size]
