"""BASELINE.json configs[0] and [1] at their real sizes on the GPU box: the bundled corpus buffers (a git-ignored
copy under data_local/, see tests/demos/corpus_demo.sh; it travels with the snapshot) loaded the way SMART's driver
loads them (smart.c:95-138: the files of index.txt concatenated up to TSIZE), 500 text-extracted patterns per length
(smart.c:148-158, seeded), bf / kmp / hor / so: every count against the oracle, a subset against the compiled
reference algorithms themselves.  Skipped where data_local/ is absent."""
import os
import re

import numpy as np
import pytest

import oracle_lib
import smart_b200 as sb

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "data_local")
ALGOS4 = ("bf", "kmp", "hor", "so")


def load_buffer(name, tsize):
    """getText (smart.c:95-138): the files named #file# in index.txt, in order, until tsize bytes are read"""
    d = os.path.join(DATA, name)
    parts, total = [], 0
    for f in re.findall(r"#([^#\n]+)#", open(os.path.join(d, "index.txt"), errors="replace").read()):
        if total >= tsize:
            break
        b = np.fromfile(os.path.join(d, f.strip()), dtype=np.uint8)[: tsize - total]
        parts.append(b)
        total += b.size
    return np.concatenate(parts)


def check_pattern_sets(y, lengths, npat, seed, oracle, ref_every=25):
    """npat patterns per length through the asynchronous entry (up to 256 scans in flight), four algorithms"""
    rng = np.random.default_rng(seed)
    refs = {a: oracle_lib.RefAlgo(a) for a in ALGOS4} if oracle_lib.ref_available() else {}
    ybuf = next(iter(refs.values())).make_text(y) if refs else None
    searched = 0
    with sb.Text.upload(y) as t:
        for m in lengths:
            pats = [y[k:k + m].copy() for k in rng.integers(0, y.size - m, npat)]
            want = [oracle.bf_count(x, y) for x in pats]
            assert min(want) >= 1  # cut from the text (smart.c:330's own in-band check)
            for a in ALGOS4:
                for lo in range(0, npat, 250):
                    tickets = [sb.submit(a, t, x) for x in pats[lo:lo + 250]]
                    got = [sb.collect(tk).count for tk in tickets]
                    assert got == want[lo:lo + 250], (a, m, lo)
                    searched += len(got)
            for i in range(0, npat, ref_every):  # the unmodified reference algorithms agree with the oracle
                for a, r in refs.items():
                    assert r.search_padded(pats[i], ybuf, y.size) == want[i], (a, m, i)
    return searched


@pytest.mark.skipif(not os.path.exists(os.path.join(DATA, "englishTexts", "index.txt")), reason="data_local/englishTexts absent")
@pytest.mark.parametrize("tsize_mb", [1, 7])
def test_config0_english_texts_500_patterns(oracle, tsize_mb):
    """configs[0]: englishTexts, default TSIZE (1 MiB = the start of bible.txt) and -tsize 7 (all 6 520 792 bytes),
    -short lengths m = 2, 4, ..., 32 (sets.h:24), 500 patterns per length"""
    y = load_buffer("englishTexts", tsize_mb << 20)
    assert y.size == (1 << 20 if tsize_mb == 1 else 6520792)
    n = check_pattern_sets(y, range(2, 33, 2), 500, 2026 + tsize_mb, oracle)
    assert n == 16 * 500 * 4


@pytest.mark.skipif(not os.path.exists(os.path.join(DATA, "protein", "index.txt")), reason="data_local/protein absent")
def test_config0_protein_500_patterns(oracle):
    y = load_buffer("protein", 7 << 20)
    assert y.size == 7154401  # hs + mj + sc + hi, below the -tsize 7 bound
    check_pattern_sets(y, (2, 4, 8, 16, 32), 500, 77, oracle, ref_every=50)


def test_config1_genome_standin_m_2_to_1024(oracle):
    """configs[1]: data/genome (sigma = 4), m = 2 ... 1024.  ecoli.txt is absent from the reference tree
    (.MISSING_LARGE_BLOBS), so the text is a stand-in of the same size over {a,c,g,t} from the seeded generator,
    labelled as such."""
    codes = sb.textgen_host(0, 4638690, 4, 1234)
    y = np.frombuffer(b"acgt", dtype=np.uint8)[codes]
    check_pattern_sets(y, (2, 4, 8, 16, 32, 64, 128, 256, 512, 1024), 500, 5, oracle, ref_every=50)
