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
