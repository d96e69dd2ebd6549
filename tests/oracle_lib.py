"""Loader for the parity checker: oracle/libsmart_oracle.so (our CPU restatement) and, when present,
oracle/_ref/<algo>.so (the unmodified reference sources compiled by oracle/Makefile).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libsmart_oracle.so")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
XSIZE = 4200


def build_oracle(ref: bool = True):
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-C", ORACLE_DIR] + targets, check=True, capture_output=True)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def as_u8(b) -> np.ndarray:
    if isinstance(b, np.ndarray):
        return np.ascontiguousarray(b, dtype=np.uint8)
    return np.frombuffer(bytes(b), dtype=np.uint8).copy()


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build_oracle(ref=False)
        L = C.CDLL(ORACLE_SO)
        L.ora_bf_count.restype = C.c_ulonglong
        L.ora_bf_count.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.ora_bf_count_fast.restype = C.c_ulonglong
        L.ora_bf_count_fast.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.ora_bf_positions.restype = C.c_ulonglong
        L.ora_bf_positions.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.ora_search.restype = C.c_longlong
        L.ora_search.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        L.ora_textgen.restype = None
        L.ora_textgen.argtypes = [C.c_int, C.c_void_p, C.c_size_t]
        self.L = L

    # ---- result specification
    def bf_count(self, x, y, fast=True) -> int:
        x, y = as_u8(x), as_u8(y)
        f = self.L.ora_bf_count_fast if fast else self.L.ora_bf_count
        return int(f(_ptr(x), x.size, _ptr(y), y.size))

    def bf_positions(self, x, y, cap=None) -> np.ndarray:
        x, y = as_u8(x), as_u8(y)
        cap = cap if cap is not None else max(1, y.size)
        out = np.zeros(cap, dtype=np.uint64)
        c = int(self.L.ora_bf_positions(_ptr(x), x.size, _ptr(y), y.size, _ptr(out), cap))
        return out[: min(c, cap)]

    def search(self, algo: str, x, y) -> int:
        x, y = as_u8(x), as_u8(y)
        r = int(self.L.ora_search(algo.encode(), _ptr(x), x.size, _ptr(y), y.size))
        if r == -2:
            raise ValueError(f"oracle: unknown algorithm {algo}")
        return r

    def textgen(self, which: int, length: int) -> np.ndarray:
        out = np.zeros(length, dtype=np.uint8)
        self.L.ora_textgen(which, _ptr(out), length)
        return out

    # ---- tables
    def table(self, name: str, x, **kw):
        x = as_u8(x)
        m = x.size
        L = self.L
        ia = lambda k: np.zeros(k, dtype=np.int32)
        if name == "mpNext":
            t = ia(m + 1); L.ora_pre_mp(_ptr(x), m, _ptr(t)); return t
        if name == "kmpNext":
            t = ia(m + 1); L.ora_pre_kmp(_ptr(x), m, _ptr(t)); return t
        if name == "hbc":
            t = ia(256); L.ora_pre_hbc(_ptr(x), m, _ptr(t)); return t
        if name == "qsBc":
            t = ia(256); L.ora_pre_qsbc(_ptr(x), m, _ptr(t)); return t
        if name == "suff":
            t = ia(m); L.ora_suffixes(_ptr(x), m, _ptr(t)); return t
        if name == "bmGs":
            s, t = ia(m), ia(m); L.ora_pre_bmgs(_ptr(x), m, _ptr(s), _ptr(t)); return t
        if name == "so":
            t = np.zeros(256, dtype=np.uint32); L.ora_pre_so.restype = C.c_uint
            lim = L.ora_pre_so(_ptr(x), m, _ptr(t)); return t, int(lim)
        if name == "sa":
            t = np.zeros(256, dtype=np.uint32); L.ora_pre_sa.restype = C.c_uint
            f = L.ora_pre_sa(_ptr(x), m, _ptr(t)); return t, int(f)
        if name == "hashq":
            xz = np.concatenate([x, np.zeros(1, np.uint8)])
            t = ia(256); sh1 = C.c_int(0)
            L.ora_pre_hashq(_ptr(xz), m, int(kw["q"]), _ptr(t), C.byref(sh1)); return t, sh1.value
        if name == "kr":
            d, hx = C.c_int(0), C.c_int(0)
            L.ora_pre_kr(_ptr(x), m, C.byref(d), C.byref(hx)); return d.value, hx.value
        if name == "buckets":
            h, nx = ia(256), ia(m); L.ora_pre_buckets(_ptr(x), m, _ptr(h), _ptr(nx)); return h, nx
        if name == "lastocc":
            t = ia(256); L.ora_pre_lastocc(_ptr(x), m, _ptr(t)); return t
        if name == "fsGs":
            t = ia(m + 1); L.ora_pre_fsgs(_ptr(x), m, _ptr(t)); return t
        if name == "graspm":
            h, nx = ia(256), ia(m + 1); L.ora_pre_graspm(_ptr(x), m, _ptr(h), _ptr(nx)); return h, nx
        if name == "bndmB":
            t = np.zeros(256, dtype=np.uint32); L.ora_pre_bndm(_ptr(x), m, _ptr(t)); return t
        if name == "brBc":
            t = ia(65536); L.ora_pre_brbc(_ptr(x), m, _ptr(t)); return t
        if name == "acEll":
            return int(L.ora_pre_ac_ell(_ptr(x), m))
        if name == "ztBc":
            t = ia(65536); L.ora_pre_ztbc(_ptr(x), m, _ptr(t)); return t
        if name == "wfr":
            t = np.zeros(65536, dtype=np.uint8); L.ora_pre_wfr(_ptr(x), m, int(kw.get("Q", 0)), _ptr(t)); return t
        if name == "aut":
            t = ia((m + 1) * 256); L.ora_pre_aut(_ptr(x), m, _ptr(t)); return t.reshape(m + 1, 256)
        if name == "bom":
            tr = np.zeros((m + 1) * 256, dtype=np.int16); te = np.zeros(m + 1, dtype=np.uint8)
            L.ora_pre_bom(_ptr(x), m, _ptr(tr), _ptr(te)); return tr.reshape(m + 1, 256), te
        if name == "qf":
            Q, S = int(kw["Q"]), int(kw["S"])
            t = ia(1 << (Q * S)); L.ora_pre_qf(_ptr(x), m, Q, S, _ptr(t)); return t
        raise KeyError(name)


def textgen_np(off: int, length: int, sigma: int, seed: int) -> np.ndarray:
    """numpy restatement of the product's counter-based text generator (include/smart_b200.h sb_textgen_host):
    bytes [off, off+length) of stream (seed, sigma).  Lets the reference arm of bench.py and the CPU baseline build
    the workload text without loading libsmartb200.so.  Word q (8 bytes, little endian) is
    mix64(seed + (q+1)*0x9E3779B97F4A7C15) (splitmix64 finaliser); sigma = 256 keeps the bytes, a power of two masks
    each byte with sigma-1, any other sigma maps byte b to (b*sigma) >> 8."""
    if length <= 0:
        return np.zeros(0, dtype=np.uint8)
    q0, q1 = off >> 3, (off + length + 7) >> 3
    with np.errstate(over="ignore"):
        q = np.arange(q0 + 1, q1 + 1, dtype=np.uint64)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + q * np.uint64(0x9E3779B97F4A7C15)
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    b = z.view(np.uint8) if z.dtype.byteorder in ("=", "<", "|") else z.byteswap().view(np.uint8)
    if sigma != 256:
        if sigma & (sigma - 1) == 0:
            b = b & np.uint8(sigma - 1)
        else:
            b = ((b.astype(np.uint16) * np.uint16(sigma)) >> np.uint16(8)).astype(np.uint8)
    lo = off - (q0 << 3)
    return np.ascontiguousarray(b[lo: lo + length])


def ref_available(algo: str = "bf") -> bool:
    return os.path.exists(os.path.join(REF_DIR, f"{algo}.so"))


class RefAlgo:
    """The reference's own search() (src/algos/<algo>.c) loaded from oracle/_ref/<algo>.so.

    The reference writes sentinels past y[n] (up to 2m bytes) and times itself through the globals of
    main.h:34-37, so the caller's text is copied into a buffer with 3*XSIZE+64 bytes of zeroed slack and the
    globals are pointed at private storage first.  int n / int count as in the reference.
    """

    def __init__(self, algo: str):
        path = os.path.join(REF_DIR, f"{algo}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.algo = algo
        self.L = C.CDLL(path)
        self.L.search.restype = C.c_int
        self.L.search.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        self._run = C.c_double(0)
        self._pre = C.c_double(0)
        self._timer = (C.c_double * 2)()
        C.c_void_p.in_dll(self.L, "run_time").value = C.addressof(self._run)
        C.c_void_p.in_dll(self.L, "pre_time").value = C.addressof(self._pre)
        C.c_void_p.in_dll(self.L, "_timer").value = C.addressof(self._timer)

    SLACK = 3 * XSIZE + 64

    def make_text(self, y) -> np.ndarray:
        y = as_u8(y)
        buf = np.zeros(y.size + self.SLACK, dtype=np.uint8)
        buf[: y.size] = y
        return buf

    def search_padded(self, x, ybuf: np.ndarray, n: int) -> int:
        """ybuf must come from make_text (slack after n is re-zeroed here)."""
        x = as_u8(x)
        xz = np.zeros(x.size + 16, dtype=np.uint8)  # p[m] == 0 as the driver guarantees (smart.c:313)
        xz[: x.size] = x
        ybuf[n : n + self.SLACK] = 0
        return int(self.L.search(_ptr(xz), x.size, _ptr(ybuf), n))

    def search(self, x, y) -> int:
        y = as_u8(y)
        return self.search_padded(x, self.make_text(y), y.size)

    @property
    def run_ms(self) -> float:
        return self._run.value

    @property
    def pre_ms(self) -> float:
        return self._pre.value

    def call_pre(self, fname: str, *args):
        fn = getattr(self.L, fname)
        return fn(*args)
