"""The product's preprocessing stage (smart_b200/csrc/prep.c via sb_prep_tables) must build tables
bit-identical to the reference's: checked against tables dumped from the reference's own exported pre*()
functions (tests/golden/tables.json), against SURVEY.md Appendix B.2, and against the oracle restatement on
random patterns.  Host-only: no GPU needed."""
import hashlib
import json
import os

import numpy as np
import pytest

import smart_b200 as sb
from golden_texts import GOLDEN


def iter_list(e, nxt):
    while e >= 0:
        yield e
        e = nxt[e]


def arr(ct, k):
    return np.ctypeslib.as_array(ct)[:k].tolist()


@pytest.fixture(scope="module")
def table_cases():
    return json.load(open(os.path.join(GOLDEN, "tables.json")))["tables"]


def test_tables_vs_reference_dump(table_cases, oracle):
    for ent in table_cases:
        x = bytes.fromhex(ent["pattern"])
        m = len(x)
        t = sb.prep_tables("kmp", x)
        assert arr(t.kmpNext, m + 1) == ent["kmpNext"]
        assert arr(t.mpNext, m + 1) == ent["mpNext"]
        assert arr(sb.prep_tables("mp", x).mpNext, m + 1) == ent["mpNext"]
        assert arr(sb.prep_tables("hor", x).hbc, 256) == ent["hbc"]
        t = sb.prep_tables("bm", x)
        assert arr(t.hbc, 256) == ent["bmBc"]
        assert arr(t.suff, m) == ent["suff"]
        assert arr(t.bmGs, m) == ent["bmGs"]
        assert arr(sb.prep_tables("qs", x).qsBc, 256) == ent["qsBc"]
        t = sb.prep_tables("so", x)
        assert arr(t.soS, 256) == ent["soS"] and t.soLim == ent["soLim"]
        t = sb.prep_tables("sa", x)
        assert arr(t.saS, 256) == ent["saS"] and t.saF == 1 << (min(m, 32) - 1)
        if m >= 2:
            assert arr(sb.prep_tables("raita", x).hbc, 256) == ent["hbc"]
        assert arr(sb.prep_tables("tunedbm", x).hbc, 256) == ent["hbc"]
        assert arr(sb.prep_tables("fs", x).fsGs, m + 1) == ent["fsGs"]  # Pre_GS dumped from fs.c
        # second batch: the reference keeps private copies of these builders (checked equal in make_golden.py)
        assert arr(sb.prep_tables("fs-w1", x).fsGs, m + 1) == ent["fsGs"]
        assert arr(sb.prep_tables("qlqs", x).qsBc, 256) == ent["qsBc"]
        t = sb.prep_tables("zt", x)
        assert arr(t.suff, m) == ent["suff"] and arr(t.bmGs, m) == ent["bmGs"]
        zt = np.ctypeslib.as_array(t.ztBc).astype(np.int32)
        assert hashlib.sha1(zt.tobytes()).hexdigest() == ent["ztBc_sha1"]
        assert hashlib.sha1(oracle.table("ztBc", x).astype(np.int32).tobytes()).hexdigest() == ent["ztBc_sha1"]
        if m >= 4:
            assert arr(sb.prep_tables("lwfr4", x).kmpNext, m + 1) == ent["kmpNext"]
        tsw = np.ctypeslib.as_array(sb.prep_tables("tsw", x).brBc).astype(np.int32)
        assert hashlib.sha1(tsw.tobytes()).hexdigest() == ent["brBc_sha1"]
        for a, digest in ent["wfrF_sha1"].items():  # the 37 weak-factor variants share two table builders
            F = bytes(sb.prep_tables(a, x).wfrF)
            assert hashlib.sha1(F).hexdigest()[:16] == digest, (a, m)
            if a == "wfr" or a.startswith("wfrq"):  # and the oracle's restatement of the same tables
                Q = int(a[4]) if a.startswith("wfrq") else 0
                assert hashlib.sha1(oracle.table("wfr", x, Q=Q).tobytes()).hexdigest()[:16] == digest, (a, m)
        br = np.ctypeslib.as_array(sb.prep_tables("br", x).brBc).astype(np.int32)
        assert hashlib.sha1(br.tobytes()).hexdigest() == ent["brBc_sha1"]  # preBrBc dumped from br.c


def test_appendix_b2_known_answers():
    x = b"GCAGAGAG"
    A, Cc, G, T = ord("A"), ord("C"), ord("G"), ord("T")
    t = sb.prep_tables("kmp", x)
    assert arr(t.kmpNext, 9) == [-1, 0, 0, -1, 1, -1, 1, -1, 1]
    assert arr(t.mpNext, 9) == [-1, 0, 0, 0, 1, 0, 1, 0, 1]
    h = sb.prep_tables("hor", x).hbc
    assert (h[A], h[Cc], h[G], h[T], h[0], h[255]) == (1, 6, 2, 8, 8, 8)
    t = sb.prep_tables("bm", x)
    assert arr(t.suff, 8) == [1, 0, 0, 2, 0, 4, 0, 8]
    assert arr(t.bmGs, 8) == [7, 7, 7, 2, 7, 4, 7, 1]
    q = sb.prep_tables("qs", x).qsBc
    assert (q[A], q[Cc], q[G], q[T], q[1]) == (2, 7, 1, 9, 9)
    t = sb.prep_tables("so", x)
    assert t.soLim == 0xFFFFFF80
    assert (t.soS[A], t.soS[Cc], t.soS[G], t.soS[T], t.soS[7]) == (0xFFFFFFAB, 0xFFFFFFFD, 0xFFFFFF56, 0xFFFFFFFF,
                                                                    0xFFFFFFFF)
    t = sb.prep_tables("sa", x)
    assert (t.saF, t.saS[A], t.saS[Cc], t.saS[G], t.saS[T]) == (0x80, 0x54, 0x02, 0xA9, 0)
    t = sb.prep_tables("hash3", x)
    sh = arr(t.hashShift, 256)
    assert (sh[211], sh[213], sh[227], sh[229], t.hashSh1) == (1, 4, 5, 0, 2)
    assert all(v == 6 for i, v in enumerate(sh) if i not in (211, 213, 227, 229))
    t = sb.prep_tables("qf23", x)
    B = arr(t.qfB, 64)
    assert {i: v for i, v in enumerate(B) if v} == {11: 2, 15: 2, 31: 1, 57: 1}


def test_tables_vs_oracle_random(oracle):
    rng = np.random.default_rng(5)
    shapes = [(1, 256), (2, 2), (3, 4), (5, 4), (8, 2), (8, 256), (9, 3), (33, 2), (64, 256), (257, 4), (4200, 2)]
    pats = [rng.integers(0, s, m, dtype=np.uint8) for m, s in shapes]
    pats += [np.frombuffer(b"ab" * 50, dtype=np.uint8), np.frombuffer(b"a" * 77, dtype=np.uint8),
             np.frombuffer(b"aab" * 33 + b"c", dtype=np.uint8)]
    for x in pats:
        m = x.size
        t = sb.prep_tables("kmp", x)
        assert arr(t.kmpNext, m + 1) == oracle.table("kmpNext", x).tolist()
        assert arr(t.mpNext, m + 1) == oracle.table("mpNext", x).tolist()
        t = sb.prep_tables("bm", x)
        assert arr(t.bmGs, m) == oracle.table("bmGs", x).tolist()
        assert arr(t.suff, m) == oracle.table("suff", x).tolist()
        assert arr(t.hbc, 256) == oracle.table("hbc", x).tolist()
        assert arr(sb.prep_tables("qs", x).qsBc, 256) == oracle.table("qsBc", x).tolist()
        S, lim = oracle.table("so", x)
        t = sb.prep_tables("so", x)
        assert arr(t.soS, 256) == S.tolist() and t.soLim == lim
        S, F = oracle.table("sa", x)
        t = sb.prep_tables("sa", x)
        assert arr(t.saS, 256) == S.tolist() and t.saF == F
        t = sb.prep_tables("ac", x)
        assert arr(t.kmpNext, m + 1) == oracle.table("kmpNext", x).tolist() and t.acEll == oracle.table("acEll", x)
        t = sb.prep_tables("smith", x)
        assert arr(t.hbc, 256) == oracle.table("hbc", x).tolist() and arr(t.qsBc, 256) == oracle.table("qsBc", x).tolist()
        for name in ("tbm", "ag"):
            t = sb.prep_tables(name, x)
            assert arr(t.bmGs, m) == oracle.table("bmGs", x).tolist() and arr(t.suff, m) == oracle.table("suff", x).tolist()
        h, nx = oracle.table("buckets", x)
        for name in ("skip", "kmpskip"):
            t = sb.prep_tables(name, x)
            assert arr(t.bucketHead, 256) == h.tolist() and arr(t.bucketNext, m) == nx.tolist()
        t = sb.prep_tables("fs", x)
        assert arr(t.hbc, 256) == oracle.table("lastocc", x).tolist() and arr(t.fsGs, m + 1) == oracle.table("fsGs", x).tolist()
        t = sb.prep_tables("graspm", x)
        h, nx = oracle.table("graspm", x)
        assert arr(t.hbc, 256) == oracle.table("lastocc", x).tolist() and arr(t.bucketHead, 256) == h.tolist()
        e_list = lambda head, nxt, c: [e for e in iter_list(head[c], nxt)]
        for c in set(x.tolist()) | {0}:
            assert e_list(arr(t.bucketHead, 256), arr(t.bucketNext, m + 1), c) == e_list(h.tolist(), nx.tolist(), c)
        assert arr(sb.prep_tables("bndm", x).bndmB, 256) == oracle.table("bndmB", x).tolist()
        for name in ("br", "tvsbs"):
            assert arr(sb.prep_tables(name, x).brBc, 65536) == oracle.table("brBc", x).tolist()
        t = sb.prep_tables("fjs", x)
        assert arr(t.qsBc, 256) == oracle.table("qsBc", x).tolist() and arr(t.kmpNext, m + 1) == oracle.table("kmpNext", x).tolist()
        assert arr(sb.prep_tables("ssabs", x).qsBc, 256) == oracle.table("qsBc", x).tolist()
        d, hx = oracle.table("kr", x)
        t = sb.prep_tables("kr", x)
        assert (t.krD, t.krHx) == (d, hx)
        for q in (3, 5, 8):
            t = sb.prep_tables(f"hash{q}", x)
            if m < q:
                assert t is None
                continue
            sh, sh1 = oracle.table("hashq", x, q=q)
            assert arr(t.hashShift, 256) == sh.tolist() and t.hashSh1 == sh1 and t.hashQ == q
        for name in ("qf23", "qf24", "qf26", "qf28", "qf33", "qf34", "qf36", "qf42", "qf43", "qf44", "qf62", "qf63"):
            Q, S_ = int(name[2]), int(name[3])
            t = sb.prep_tables(name, x)
            if m <= Q:
                assert t is None
                continue
            assert arr(t.qfB, 1 << (Q * S_)) == oracle.table("qf", x, Q=Q, S=S_).tolist(), name


def test_not_applicable_domains():
    """-1 domains of the reference."""
    for name, min_m in [("raita", 2), ("hash3", 3), ("hash5", 5), ("hash8", 8), ("qf23", 3), ("qf33", 4),
                        ("qf44", 5), ("qf62", 7), ("bf", 1), ("kmp", 1), ("so", 1)]:
        for m in range(1, 10):
            assert sb.applicable(name, m) == (m >= min_m), (name, m)
    # every registered algorithm against the rule table of the tests (golden_texts.min_m)
    from golden_texts import ALGOS, EXTRA_ALGOS, min_m as rule

    assert sb.ALGORITHMS == tuple(ALGOS) + EXTRA_ALGOS
    for name in ALGOS + EXTRA_ALGOS:
        for m in (1, 2, 3, 4, 5, 6, 7, 8, 9, 31, 32, 33):
            assert sb.applicable(name, m) == (m >= rule(name)), (name, m)


def test_prep_tables_struct_can_be_reused_across_families():
    """the scalars that lie between the big arrays (qfQ, qfS, wfrQ) and everything before them are reset
    on every call; a struct that last held a QF or WFR pattern does not leak that into the next family."""
    x = np.frombuffer(b"abracadabra", dtype=np.uint8)
    t = sb.Tables()
    L = sb.lib()
    import ctypes as C

    def prep(name):
        assert L.sb_prep_tables(sb.algo_id(name), x.ctypes.data, x.size, C.byref(t)) == 0

    prep("qf23")
    assert (t.qfQ, t.qfS) == (2, 3)
    prep("wfrq4")
    assert t.wfrQ == 4 and (t.qfQ, t.qfS) == (0, 0)
    prep("hash3")
    assert t.hashQ == 3 and t.wfrQ == 0 and (t.qfQ, t.qfS) == (0, 0)
    prep("kmp")
    assert t.hashQ == 0 and t.wfrQ == 0 and t.kmpNext[0] == -1 and t.m == x.size
    fresh = sb.prep_tables("kmp", x)
    assert list(t.kmpNext[: x.size + 1]) == list(fresh.kmpNext[: x.size + 1]) and list(t.hbc) == list(fresh.hbc)


def test_automaton_table_equals_the_reference_builder(oracle):
    """sb_prep_automaton (from the border table) == the oracle's restatement of preSMA == the reference's own preSMA
    (include/AUTOMATON.h:185-198, exported by oracle/_ref/aut.so) on random and periodic patterns."""
    import ctypes as C

    import oracle_lib

    rng = np.random.default_rng(9)
    ref = oracle_lib.RefAlgo("aut") if oracle_lib.ref_available("aut") else None
    pats = [np.frombuffer(b"abracadabra", np.uint8), np.frombuffer(b"aaaaaaaa", np.uint8), np.frombuffer(b"a", np.uint8),
            np.frombuffer(b"abababab" * 20, np.uint8)]
    pats += [rng.integers(0, s_, m, dtype=np.uint8) for s_ in (2, 4, 256) for m in (1, 2, 7, 33, 300, 4200)]
    for x in pats:
        x = np.ascontiguousarray(x)
        ours = sb.prep_automaton(x).astype(np.int32)
        want = oracle.table("aut", x)
        assert np.array_equal(ours, want), x.size
        if ref is not None and x.size <= 300:
            t = np.full((x.size + 1) * 256, -1, dtype=np.int32)  # aut.c:33 memset(-1) before preSMA
            ref.call_pre("preSMA", x.ctypes.data_as(C.c_void_p), C.c_int(x.size), t.ctypes.data_as(C.c_void_p))
            assert np.array_equal(t.reshape(x.size + 1, 256), want), x.size


def test_bom_factor_oracle_equals_the_reference_builder(oracle):
    """sb_prep_bom (dense table) == the oracle's list-based restatement == the reference's own oracle() queried through
    its getTransition() (bom.c:28-70, both exported by oracle/_ref/bom.so), transitions and terminal flags."""
    import ctypes as C

    import oracle_lib

    rng = np.random.default_rng(19)
    ref = oracle_lib.RefAlgo("bom") if oracle_lib.ref_available("bom") else None
    pats = [np.frombuffer(b"abracadabra", np.uint8), np.frombuffer(b"aaaaaaaa", np.uint8), np.frombuffer(b"a", np.uint8),
            np.frombuffer(b"abababab" * 9, np.uint8), np.frombuffer(b"gcagagag", np.uint8)]
    pats += [rng.integers(0, s_, m, dtype=np.uint8) for s_ in (2, 4, 256) for m in (1, 2, 7, 33, 200, 4200)]
    for x in pats:
        x = np.ascontiguousarray(x)
        m = x.size
        trans, term = sb.prep_bom(x)
        otrans, oterm = oracle.table("bom", x)
        assert np.array_equal(trans, otrans) and np.array_equal(term, oterm), m
        if ref is not None and m <= 200:
            L = (C.c_void_p * 4202)()
            T = (C.c_char * 4202)()
            ref.L.oracle.restype = None
            ref.L.oracle(x.ctypes.data_as(C.c_void_p), C.c_int(m), T, L)
            ref.L.getTransition.restype = C.c_int
            for p in range(m + 1):
                assert (T[p] != b"\x00") == bool(term[p]), (m, p)
                for c in set(x.tolist()) | {0, 255}:
                    got = ref.L.getTransition(x.ctypes.data_as(C.c_void_p), C.c_int(p), L, C.c_ubyte(c))
                    assert got == trans[p, c], (m, p, c)


def test_round2_bucket_and_run_builders_follow_their_definitions():
    """sb_prep_skipq (skip3.c:33-42, skip5.c:36-45) and sb_prep_bsdm (bsdm.c:33-55, bsdm4.c:38-66) feed the table-driven
    kernels only; the GPU tests pin them through counts, this pins them on the CPU by their definitions: every pattern
    q-gram filed exactly once under its hash, newest first; the first longest run of pairwise distinct q-grams."""
    import ctypes as C

    L = sb.lib()
    rng = np.random.default_rng(2026)

    def skip_hash(p, Q):
        if Q == 3:
            return (int(p[0]) << 4) + (int(p[1]) << 2) + int(p[2])
        return sum(int(p[k]) << (Q - 1 - k) for k in range(Q))

    def bsdm_hash(p, Q):
        step = 2 if Q == 4 else 1
        return sum(int(p[t]) << (step * (Q - 1 - t)) for t in range(Q)) & 0xFFFF

    for sigma in (2, 4, 20, 256):
        for m in (8, 9, 31, 64, 257):
            x = rng.integers(0, sigma, m, dtype=np.uint8)
            xp = x.ctypes.data_as(C.c_void_p)
            for Q in (3, 5, 6, 7, 8):
                head = np.zeros(65536, dtype=np.int32)
                nxt = np.zeros(m, dtype=np.int32)
                L.sb_prep_skipq(xp, C.c_int(m), C.c_int(Q), head.ctypes.data_as(C.c_void_p), nxt.ctypes.data_as(C.c_void_p))
                want = {}
                for i in range(m - Q + 1):
                    want.setdefault(skip_hash(x[i:i + Q], Q), []).insert(0, i)
                assert set(np.nonzero(head >= 0)[0].tolist()) == set(want)
                for h, offs in want.items():
                    got, e = [], int(head[h])
                    while e >= 0:
                        got.append(e)
                        e = int(nxt[e])
                    assert got == offs, (sigma, m, Q, h)
            for Q in (1, 4, 5, 7, 8):
                pos = np.zeros(65536, dtype=np.int32)
                st, ln = C.c_int(-1), C.c_int(-1)
                L.sb_prep_bsdm(xp, C.c_int(m), C.c_int(Q), pos.ctypes.data_as(C.c_void_p), C.byref(st), C.byref(ln))
                g = [bsdm_hash(x[i:i + Q], Q) for i in range(m - Q + 1)]
                best_st, best_len, j = 0, 0, 0
                for i in range(len(g)):  # the longest window of pairwise distinct grams ending at i starts at j
                    while g[i] in g[j:i]:
                        j += 1
                    if i - j + 1 > best_len:
                        best_st, best_len = j, i - j + 1
                assert (st.value, ln.value) == (best_st, best_len), (sigma, m, Q)
                want_pos = np.full(65536, -1, dtype=np.int32)
                for i in range(best_len):
                    want_pos[g[best_st + i]] = i
                assert np.array_equal(pos, want_pos), (sigma, m, Q)
