"""The oracle (oracle/smart_oracle.c) pinned against the reference's own fixtures, the survey's known-answer
vectors, the committed golden outputs of the unmodified reference, and (when oracle/_ref is present) the
reference itself on fresh random inputs."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_lib
from fixtures import TEST_C_FIXTURES, seeded_random_fixtures
from golden_texts import ALGOS, GOLDEN, TEXT_NAMES, make_patterns, make_text, expected

REF_DATA = "/root/reference/data"




@pytest.mark.parametrize("x,y,count", TEST_C_FIXTURES)
def test_testc_fixtures(oracle, x, y, count):
    assert oracle.bf_count(x, y, fast=False) == count
    assert oracle.bf_count(x, y, fast=True) == count
    assert len(oracle.bf_positions(x, y)) == count
    for a in ALGOS:
        assert oracle.search(a, x, y) == expected(a, len(x), count, len(y)), a


def test_testc_random_cases(oracle):
    for x, y in seeded_random_fixtures():
        c = oracle.bf_count(x, y, fast=False)
        assert c >= 1  # test.c:330,340: the pattern is a prefix of the text
        for a in ALGOS:
            assert oracle.search(a, x, y) == expected(a, len(x), c, len(y)), a


def test_positions_sorted_and_exact(oracle):
    y = b"abracadabra abracadabra"
    pos = oracle.bf_positions(b"abra", y)
    assert pos.tolist() == [0, 7, 12, 19]
    assert oracle.bf_positions(b"a", b"aaaa", cap=2).tolist() == [0, 1]


@pytest.fixture(scope="module")
def golden():
    return json.load(open(os.path.join(GOLDEN, "counts.json")))


@pytest.mark.parametrize("tname", TEXT_NAMES)
def test_oracle_vs_golden_reference_counts(oracle, golden, tname):
    """Every restated scan against the recorded output of the unmodified reference algorithm."""
    assert tuple(golden["algos"]) == ALGOS
    y = make_text(tname, oracle)
    cases = [c for c in golden["cases"] if c["text"] == tname]
    assert cases
    for idx, c in enumerate(cases):
        x = make_patterns(tname, y, c["m"])[c["pi"]]
        assert hashlib.sha1(bytes(x)).hexdigest()[:16] == c["sha1"], "pattern recipe drifted"
        ref = dict(zip(ALGOS, c["ref"]))
        assert oracle.bf_count(x, y) == ref["bf"]
        # the slow restated scans on a third of the cases, the cheap ones everywhere
        algos = ALGOS if idx % 3 == 0 else ("kmp", "hor", "so", "hash3", "qf23")
        for ai, a in enumerate(algos):
            # the weak-factor tables saturate on the two-symbol text and every window is read in full, O(n*m) as in
            # the reference: its 37 variants take turns there for the long patterns
            if tname == "rand2" and c["m"] >= 512 and "wfr" in a and (ai + idx) % 4:
                continue
            assert oracle.search(a, x, y) == ref[a], (tname, c["m"], c["pi"], a)


def test_golden_reference_equals_bf_or_minus1(golden):
    for c in golden["cases"]:
        ref = dict(zip(ALGOS, c["ref"]))
        for a in ALGOS:
            assert ref[a] == expected(a, c["m"], ref["bf"]), (c["text"], c["m"], a)


def test_textgen_stream_md5(oracle):
    """md5 of the files the reference textgen writes (glibc rand stream)."""
    assert hashlib.md5(oracle.textgen(0, 5_000_000).tobytes()).hexdigest() == "a8e4cecf43689f3964b4fdfe191aa01b"
    assert hashlib.md5(oracle.textgen(1, 5_000_000).tobytes()).hexdigest() == "c371ad58eda30fa18f35a77b1e775ece"


@pytest.mark.skipif(not os.path.exists(REF_DATA), reason="reference corpus only exists in the build container")
def test_appendix_b_known_answers(oracle):
    bible = np.fromfile(f"{REF_DATA}/englishTexts/bible.txt", dtype=np.uint8)
    for p, full, first in [(b"e", 396042, 101542), (b"th", 148979, 39535), (b"the", 93459, 26408),
                           (b" the ", 59835, 17398), (b"LORD", 6369, 2321), (b"In the beginning", 4, 1),
                           (b"and the LORD said unto Moses", 0, 0), (b"zzzz", 0, 0)]:
        assert oracle.bf_count(p, bible) == full
        assert oracle.bf_count(p, bible[: 1 << 20]) == first
        assert oracle.search("kmp", p, bible) == full
    hs = np.fromfile(f"{REF_DATA}/protein/hs.txt", dtype=np.uint8)
    for p, c in [(b"LL", 33659), (b"MAS", 527), (b"GPPG", 1022), (b"EEEEEEEE", 213)]:
        assert oracle.bf_count(p, hs) == c
    r2 = oracle.textgen(0, 5_000_000)
    assert oracle.bf_count(bytes([0, 1]), r2) == 1249983
    assert oracle.bf_count(bytes([1] * 8), r2) == 19544
    assert oracle.bf_count(bytes([0, 1] * 16), r2) == 0
    assert oracle.bf_count(r2[1000:1064], r2) == 1
    assert oracle.bf_count(r2[5000:6024], r2) == 1
    r4 = oracle.textgen(1, 5_000_000)
    assert oracle.bf_count(bytes([0, 1, 2, 3]), r4) == 19300
    assert oracle.bf_count(bytes([3] * 6), r4) == 1261
    assert oracle.bf_count(r4[77:93], r4) == 1


@pytest.mark.ref
@pytest.mark.skipif(not oracle_lib.ref_available(), reason="oracle/_ref not built")
def test_oracle_vs_live_reference(oracle):
    """Fresh random inputs through the compiled reference and the restatement."""
    rng = np.random.default_rng(99)
    refs = {a: oracle_lib.RefAlgo(a) for a in ALGOS}
    for sigma in (2, 4, 20, 256):
        y = rng.integers(0, sigma, 20000, dtype=np.uint8)
        for m in (1, 2, 3, 5, 8, 9, 16, 31, 33, 64, 200, 1000):
            k = int(rng.integers(0, y.size - m))
            x = y[k : k + m].copy()
            for a in ALGOS:
                assert oracle.search(a, x, y) == refs[a].search(x, y), (sigma, m, a)


def test_rand_parity_generator_host_mirror_equals_glibc_stream(oracle):
    """sb_textgen_rand_host jumps into glibc's rand() stream (x^P mod x^31 - x^28 - 1 over Z_2^32) instead of stepping
    from the start; it must reproduce SMART's rand<sigma>.txt files (src/textgen.c:34-54) at any offset.  The device
    generator (sb_text_generate_rand) runs the same jump + recurrence per 4096-byte block."""
    import hashlib

    import smart_b200 as sb

    md5 = {2: "a8e4cecf43689f3964b4fdfe191aa01b", 4: "c371ad58eda30fa18f35a77b1e775ece"}
    for which, sigma in enumerate((2, 4, 8, 16, 32, 64, 128, 250)):
        if which not in (0, 1, 3, 7):
            continue  # (each oracle call replays the stream from its start)
        ref = oracle.textgen(which, 5000000)
        whole = sb.textgen_rand_host(sigma, 0, 5000000)
        assert np.array_equal(whole, ref), sigma
        if sigma in md5:
            assert hashlib.md5(whole.tobytes()).hexdigest() == md5[sigma]
        for off, ln in ((1, 100), (4095, 10), (4096, 5000), (1234567, 4099), (4999990, 10)):
            assert np.array_equal(sb.textgen_rand_host(sigma, off, ln), ref[off:off + ln]), (sigma, off)
    with pytest.raises(sb.SmartB200Error):
        sb.textgen_rand_host(3, 0, 10)
