"""ctypes binding of include/smart_b200.h.

Argument meaning and error behaviour mirror the reference's plugin interface
(``int search(unsigned char *x, int m, unsigned char *y, int n)``, src/algos/include/main.h:39):
``search`` returns the occurrence count, ``NOT_APPLICABLE`` (-1) where the reference algorithm
returns -1, and raises :class:`SmartB200Error` for anything the C layer reports as an error.
The library is loaded from the tree (never from site-packages) and loading fails loudly if it
has not been built.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SMART_B200_LIB selects an alternative in-tree build (tuning variants made by tools/build_variants.sh)
_LIB_PATH = os.environ.get("SMART_B200_LIB") or os.path.join(_HERE, "lib", "libsmartb200.so")

XSIZE = 4200
NOT_APPLICABLE = -1
FLAG_POSITIONS = 0x1
FLAG_NO_SKIP = 0x2
FLAG_DENSE = 0x4
FLAG_NO_PERIOD = 0x8
FLAG_FILTER_SHIFT = 0x10
FLAG_FILTER_B4 = 0x20
FLAG_FILTER_B8 = 0x40
FLAG_FILTER_B16 = 0x80
FLAG_FILTER_B32 = 0x100
FLAG_FILTER_B4L8 = 0x200
FLAG_FILTER_B4L16 = 0x400
FLAG_FILTER_B8L16 = 0x800
FLAG_FILTER_B2 = 0x1000
FLAG_NO_PACK = 0x2000
FLAG_TABLE_SCAN = 0x4000

ALGORITHMS = (
    "bf", "mp", "kmp", "bm", "hor", "qs", "raita", "tunedbm", "so", "sa", "kr",
    "hash3", "hash5", "hash8",
    "qf23", "qf24", "qf26", "qf28", "qf33", "qf34", "qf36", "qf42", "qf43", "qf44", "qf62", "qf63",
    "smith", "tbm", "ag", "ac",
    "bndm", "br", "tvsbs", "fjs", "ssabs", "ts",
    "skip", "kmpskip", "fs", "graspm",
    # second batch: registered with -1 domain and scan direction, prep.c
    "aoso2", "aoso4", "aoso6", "faoso2", "faoso4", "svm0", "blim", "ufndmq2", "ufndmq4", "ufndmq6", "ufndmq8", "aut",
    "smoa", "tw", "rcolussi", "bndmq2", "bndmq4", "bndmq6", "lbndm", "sbndm", "sbndm-bmh", "fsbndm-w2", "fndm", "tndm",
    "tndma", "bxs", "bxs2", "bxs4", "bxs6", "bxs8", "bww", "dbww", "dbww2", "bom", "bom2", "ebom",
    "fbom", "sebom", "sfbom", "ildm1", "ildm2", "fdm", "bsdm", "bsdm4", "bsdm5", "bsdm7", "bsdm8", "ww",
    "wfr", "wfr2", "wfr3", "wfr4", "wfr5", "wfr6", "wfr7", "wfr8", "wfrq2", "wfrq3", "wfrq4", "wfrq5",
    "wfrq6", "wfrq7", "wfrq8", "twfr", "twfr2", "twfr3", "twfr4", "twfr5", "twfr6", "twfr7", "twfr8", "twfrq2",
    "twfrq3", "twfrq4", "twfrq5", "twfrq6", "twfrq7", "twfrq8", "lwfr2", "lwfr3", "lwfr4", "lwfr5", "lwfr6", "lwfr7",
    "lwfr8", "skip3", "skip5", "skip6", "skip7", "skip8", "bfs", "ffs", "iom", "wom", "pbmh", "qlqs",
    "tsw", "wc", "zt", "tvsbs-w2", "fs-w1", "ssef",
    # the reference's default selection (src/algorithms.h:133); exact count also where the reference's own is wrong
    "sbndm2",
)


class SmartB200Error(RuntimeError):
    pass


def library_path() -> str:
    return _LIB_PATH


def build_library(verbose: bool = False) -> str:
    """Compile the CUDA/C sources in-tree for sm_100a (make in smart_b200/csrc)."""
    out = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if verbose or out.returncode:
        print(out.stdout[-4000:], out.stderr[-4000:])
    if out.returncode:
        raise SmartB200Error("building libsmartb200.so failed")
    return _LIB_PATH


class Timing(C.Structure):
    _fields_ = [
        ("prep_host_ms", C.c_double),
        ("upload_ms", C.c_double),
        ("kernel_ms", C.c_double),
        ("wall_ms", C.c_double),
        ("kernels_launched", C.c_int),
        ("filter_kind", C.c_int),
        ("verified", C.c_ulonglong),
    ]


class Tables(C.Structure):
    _fields_ = [
        ("m", C.c_int),
        ("mpNext", C.c_int * (XSIZE + 1)),
        ("kmpNext", C.c_int * (XSIZE + 1)),
        ("hbc", C.c_int * 256),
        ("qsBc", C.c_int * 256),
        ("suff", C.c_int * XSIZE),
        ("bmGs", C.c_int * XSIZE),
        ("soS", C.c_uint * 256),
        ("soLim", C.c_uint),
        ("saS", C.c_uint * 256),
        ("saF", C.c_uint),
        ("hashShift", C.c_int * 256),
        ("hashSh1", C.c_int),
        ("hashQ", C.c_int),
        ("krD", C.c_int),
        ("krHx", C.c_int),
        ("acEll", C.c_int),
        ("bucketHead", C.c_int * 256),
        ("bucketNext", C.c_int * (XSIZE + 1)),
        ("fsGs", C.c_int * (XSIZE + 1)),
        ("bndmB", C.c_uint * 256),
        ("brBc", C.c_int * 65536),
        ("qfQ", C.c_int),
        ("qfS", C.c_int),
        ("qfB", C.c_int * (1 << 18)),
        ("wfrQ", C.c_int),
        ("wfrF", C.c_ubyte * 65536),
        ("ztBc", C.c_int * 65536),
    ]


class HostJob(C.Structure):
    _fields_ = [
        ("algo", C.c_int),
        ("x", C.c_void_p),
        ("m", C.c_int),
        ("flags", C.c_uint),
        ("starts", C.c_ulonglong),
        ("count", C.c_longlong),
        ("pre_ms", C.c_double),
        ("run_ms", C.c_double),
    ]


_lib = None

# every symbol include/smart_b200.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("sb_algo_count", C.c_int, []),
    ("sb_algo_name", C.c_char_p, [C.c_int]),
    ("sb_algo_id", C.c_int, [C.c_char_p]),
    ("sb_algo_applicable", C.c_int, [C.c_int, C.c_int]),
    ("sb_device_count", C.c_int, []),
    ("sb_set_device", C.c_int, [C.c_int]),
    ("sb_get_device", C.c_int, []),
    ("sb_last_error", C.c_char_p, []),
    ("sb_version", C.c_char_p, []),
    ("sb_text_upload", C.c_void_p, [C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.c_int]),
    ("sb_text_generate", C.c_void_p, [C.c_size_t, C.c_int, C.c_uint64, C.POINTER(C.c_int), C.c_int]),
    ("sb_text_generate_rand", C.c_void_p, [C.c_int, C.c_size_t, C.POINTER(C.c_int), C.c_int]),
    ("sb_textgen_rand_host", C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t]),
    ("sb_text_generate_part", C.c_void_p, [C.c_size_t, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int]),
    ("sb_text_upload_part", C.c_void_p, [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int]),
    ("sb_shard_range", C.c_int, [C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                 C.POINTER(C.c_size_t)]),
    ("sb_text_length", C.c_size_t, [C.c_void_p]),
    ("sb_text_shards", C.c_int, [C.c_void_p]),
    ("sb_text_entropy_bits", C.c_double, [C.c_void_p]),
    ("sb_text_packed_bits", C.c_int, [C.c_void_p]),
    ("sb_text_download", C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]),
    ("sb_text_free", None, [C.c_void_p]),
    ("sb_textgen_host", None, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_uint64]),
    ("sb_host_alloc", C.c_void_p, [C.c_size_t]),
    ("sb_host_free", None, [C.c_void_p]),
    ("sb_search64", C.c_longlong, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_uint, C.c_void_p, C.c_size_t,
                                   C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    ("sb_search64_submit", C.c_longlong, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_uint]),
    ("sb_search_collect", C.c_longlong, [C.c_longlong, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    ("sb_search_multi64", C.c_longlong, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_void_p,
                                         C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    ("sb_last_timing", None, [C.POINTER(Timing)]),
    ("sb_stream_timer_begin", C.c_int, [C.c_int]),
    ("sb_stream_timer_end_ms", C.c_double, [C.c_int]),
    ("sb_search_host", C.c_longlong, [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_uint,
                                      C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    ("sb_search_host_batch", C.c_longlong, [C.POINTER(HostJob), C.c_int, C.c_void_p, C.c_size_t]),
    ("sb_plugin_search", C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                   C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    ("sb_prep_tables", C.c_int, [C.c_int, C.c_void_p, C.c_int, C.POINTER(Tables)]),
    ("sb_prep_automaton", C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    ("sb_prep_bom", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
]


def lib() -> C.CDLL:
    """Load the in-tree shared library (fails loudly when it is missing)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise SmartB200Error(
                f"{_LIB_PATH} is missing: build it with `make -C smart_b200/csrc` "
                "(or __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(_LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _err() -> str:
    return (lib().sb_last_error() or b"").decode(errors="replace")


def _as_bytes_ptr(buf):
    """(pointer, length, keepalive) for bytes / bytearray / numpy uint8 arrays"""
    if isinstance(buf, np.ndarray):
        if buf.dtype != np.uint8 or not buf.flags["C_CONTIGUOUS"]:
            buf = np.ascontiguousarray(buf, dtype=np.uint8)
        return buf.ctypes.data, buf.size, buf
    if isinstance(buf, (bytes, bytearray, memoryview)):
        arr = np.frombuffer(buf, dtype=np.uint8)
        return arr.ctypes.data, arr.size, arr
    raise TypeError(f"unsupported buffer type {type(buf)!r}")


def device_count() -> int:
    return lib().sb_device_count()


def set_device(device: int) -> None:
    """CUDA device of search_host / plugin_search (one process per GPU: its local rank)."""
    if lib().sb_set_device(device) != 0:
        raise SmartB200Error(lib().sb_last_error().decode())


def get_device() -> int:
    return lib().sb_get_device()


def algo_id(name: str) -> int:
    i = lib().sb_algo_id(name.encode())
    if i < 0:
        raise SmartB200Error(f"unknown algorithm {name!r}")
    return i


def applicable(name: str, m: int) -> bool:
    return lib().sb_algo_applicable(algo_id(name), m) == 0


def shard_range(n_total: int, nparts: int, part: int):
    """(base, own, len) of part `part`: owns starts [base, base+own), needs bytes [base, base+len)."""
    b, o, l = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
    if lib().sb_shard_range(n_total, nparts, part, C.byref(b), C.byref(o), C.byref(l)) != 0:
        raise SmartB200Error("sb_shard_range: bad arguments")
    return b.value, o.value, l.value


def _devices(devices):
    if devices is None:
        return None, 1
    arr = (C.c_int * len(devices))(*devices)
    return arr, len(devices)


class Text:
    """Device-resident text (replaces SMART's shm text segment, smart.c:553-568)."""

    def __init__(self, handle):
        if not handle:
            raise SmartB200Error(_err())
        self._h = C.c_void_p(handle)

    @classmethod
    def upload(cls, data, devices=None) -> "Text":
        ptr, n, keep = _as_bytes_ptr(data)
        dev, nd = _devices(devices)
        return cls(lib().sb_text_upload(ptr, n, dev, nd))

    @classmethod
    def generate(cls, n: int, sigma: int, seed: int, devices=None) -> "Text":
        dev, nd = _devices(devices)
        return cls(lib().sb_text_generate(n, sigma, seed, dev, nd))

    @classmethod
    def generate_rand(cls, sigma: int, n: int = 5000000, devices=None) -> "Text":
        """SMART's own rand<sigma> buffer (src/textgen.c: glibc rand() % sigma), generated on the device"""
        dev, nd = _devices(devices)
        return cls(lib().sb_text_generate_rand(sigma, n, dev, nd))

    @classmethod
    def generate_part(cls, n_total: int, sigma: int, seed: int, nparts: int, part: int, device: int = 0) -> "Text":
        return cls(lib().sb_text_generate_part(n_total, sigma, seed, nparts, part, device))

    @classmethod
    def upload_part(cls, data, nparts: int, part: int, device: int = 0) -> "Text":
        ptr, n, keep = _as_bytes_ptr(data)
        return cls(lib().sb_text_upload_part(ptr, n, nparts, part, device))

    def __len__(self) -> int:
        return lib().sb_text_length(self._h)

    @property
    def shards(self) -> int:
        return lib().sb_text_shards(self._h)

    @property
    def entropy_bits(self) -> float:
        return lib().sb_text_entropy_bits(self._h)

    @property
    def packed_bits(self) -> int:
        """bits per character of the packed copy of a text over <= 4 distinct byte values (2, or 1 for two), else 0"""
        return lib().sb_text_packed_bits(self._h)

    def download(self, off: int, length: int) -> np.ndarray:
        out = np.empty(length, dtype=np.uint8)
        if lib().sb_text_download(self._h, off, length, out.ctypes.data) != 0:
            raise SmartB200Error(_err())
        return out

    def free(self):
        if self._h:
            lib().sb_text_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.free()


@dataclass
class SearchResult:
    count: int
    pre_ms: float
    run_ms: float
    positions: np.ndarray | None = None


def last_timing() -> Timing:
    t = Timing()
    lib().sb_last_timing(C.byref(t))
    return t


def stream_timer_begin(device: int = 0):
    if lib().sb_stream_timer_begin(device) != 0:
        raise SmartB200Error(_err())


def stream_timer_end_ms(device: int = 0) -> float:
    ms = lib().sb_stream_timer_end_ms(device)
    if ms < 0:
        raise SmartB200Error(_err())
    return ms


def search(algo: str, text: Text, pattern, flags: int = 0, positions_cap: int = 0) -> SearchResult:
    """sb_search64: count (and optionally sorted positions) of `pattern` in a resident text."""
    ptr, m, keep = _as_bytes_ptr(pattern)
    pre, run = C.c_double(0), C.c_double(0)
    pos = None
    pos_ptr = None
    if flags & FLAG_POSITIONS:
        pos = np.empty(max(positions_cap, 1), dtype=np.uint64)
        pos_ptr = pos.ctypes.data
    r = lib().sb_search64(algo_id(algo), text._h, ptr, m, flags, pos_ptr, positions_cap, C.byref(pre), C.byref(run))
    if r < -1:
        raise SmartB200Error(f"sb_search64 failed ({r}): {_err()}")
    if pos is not None:
        pos = pos[: max(0, min(int(r), positions_cap))]
    return SearchResult(int(r), pre.value, run.value, pos)


def submit(algo: str, text: Text, pattern, flags: int = 0) -> int:
    """sb_search64_submit: enqueue a count-only search, return its ticket (the pattern is copied at submit time)."""
    ptr, m, keep = _as_bytes_ptr(pattern)
    tk = lib().sb_search64_submit(algo_id(algo), text._h, ptr, m, flags)
    if tk < 0:
        raise SmartB200Error(f"sb_search64_submit failed ({tk}): {_err()}")
    return int(tk)


def collect(ticket: int) -> SearchResult:
    """sb_search_collect: wait for one ticket; run_ms is the CUDA-event time of its scan kernels."""
    pre, run = C.c_double(0), C.c_double(0)
    r = lib().sb_search_collect(ticket, C.byref(pre), C.byref(run))
    if r < -1:
        raise SmartB200Error(f"sb_search_collect failed ({r}): {_err()}")
    return SearchResult(int(r), pre.value, run.value)


def search_multi(algo: str, text: Text, patterns, flags: int = 0):
    """sb_search_multi64: `patterns` is a 2-D uint8 array (npat x m). Returns (counts array, pre_ms, run_ms)."""
    pats = np.ascontiguousarray(patterns, dtype=np.uint8)
    if pats.ndim != 2:
        raise ValueError("patterns must be a 2-D array (npat x m)")
    npat, m = pats.shape
    counts = np.zeros(npat, dtype=np.uint64)
    pre, run = C.c_double(0), C.c_double(0)
    r = lib().sb_search_multi64(algo_id(algo), text._h, pats.ctypes.data, npat, m, flags, counts.ctypes.data,
                                C.byref(pre), C.byref(run))
    if r < -1:
        raise SmartB200Error(f"sb_search_multi64 failed ({r}): {_err()}")
    if r == -1:
        return None, pre.value, run.value
    return counts, pre.value, run.value


def search_host(algo: str, pattern, text, flags: int = 0) -> SearchResult:
    """sb_search_host: text in HOST memory, streamed through device 0 (the e2e path)."""
    pptr, m, k1 = _as_bytes_ptr(pattern)
    tptr, n, k2 = _as_bytes_ptr(text)
    pre, run = C.c_double(0), C.c_double(0)
    r = lib().sb_search_host(algo_id(algo), pptr, m, tptr, n, flags, C.byref(pre), C.byref(run))
    if r < -1:
        raise SmartB200Error(f"sb_search_host failed ({r}): {_err()}")
    return SearchResult(int(r), pre.value, run.value)


def search_host_batch(jobs, text, starts: int = 0):
    """sb_search_host_batch: `jobs` is a list of (algo, pattern) or (algo, pattern, flags); the host text crosses PCIe
    once for all of them.  starts > 0: only start positions below it are counted (the buffer is one part of a larger
    text plus its halo).  Returns a list of SearchResult (count -1 where the algorithm is not applicable)."""
    tptr, n, keep_t = _as_bytes_ptr(text)
    arr = (HostJob * len(jobs))()
    keep = []
    for i, job in enumerate(jobs):
        algo, pattern = job[0], job[1]
        pptr, m, k = _as_bytes_ptr(pattern)
        keep.append(k)
        arr[i].algo, arr[i].x, arr[i].m, arr[i].flags = algo_id(algo), pptr, m, (job[2] if len(job) > 2 else 0)
        arr[i].starts = starts
    r = lib().sb_search_host_batch(arr, len(jobs), tptr, n)
    if r < -1:
        raise SmartB200Error(f"sb_search_host_batch failed ({r}): {_err()}")
    return [SearchResult(int(j.count), j.pre_ms, j.run_ms) for j in arr]


def plugin_search(algo: str, pattern, text) -> int:
    """The legacy plugin body: reference signature, int count, -1 when not applicable."""
    pptr, m, k1 = _as_bytes_ptr(pattern)
    tptr, n, k2 = _as_bytes_ptr(text)
    pre, run = C.c_double(0), C.c_double(0)
    return lib().sb_plugin_search(algo_id(algo), pptr, m, tptr, n, C.byref(pre), C.byref(run))


def prep_tables(algo: str, pattern) -> Tables | None:
    """Build the preprocessing tables of `algo` for `pattern` (None when not applicable)."""
    ptr, m, keep = _as_bytes_ptr(pattern)
    t = Tables()
    r = lib().sb_prep_tables(algo_id(algo), ptr, m, C.byref(t))
    if r == -1:
        return None
    if r != 0:
        raise SmartB200Error("sb_prep_tables: bad argument")
    return t


def prep_automaton(pattern) -> np.ndarray:
    """sb_prep_automaton: the transition table of aut.c as an (m+1) x 256 uint16 array"""
    ptr, m, keep = _as_bytes_ptr(pattern)
    out = np.zeros((m + 1, 256), dtype=np.uint16)
    if lib().sb_prep_automaton(ptr, m, out.ctypes.data) != 0:
        raise SmartB200Error("sb_prep_automaton: bad argument")
    return out


def prep_bom(pattern):
    """sb_prep_bom: (trans, term) of bom.c's factor oracle: (m+1) x 256 int16 (-1 = undefined) and m+1 flags"""
    ptr, m, keep = _as_bytes_ptr(pattern)
    trans = np.zeros((m + 1, 256), dtype=np.int16)
    term = np.zeros(m + 1, dtype=np.uint8)
    if lib().sb_prep_bom(ptr, m, trans.ctypes.data, term.ctypes.data) != 0:
        raise SmartB200Error("sb_prep_bom: bad argument")
    return trans, term


def textgen_rand_host(sigma: int, off: int, length: int) -> np.ndarray:
    """host mirror of Text.generate_rand: bytes [off, off+length) of SMART's rand<sigma>.txt"""
    out = np.empty(length, dtype=np.uint8)
    if lib().sb_textgen_rand_host(out.ctypes.data, sigma, off, length) != 0:
        raise SmartB200Error("sb_textgen_rand_host: bad arguments")
    return out


def textgen_host(off: int, length: int, sigma: int, seed: int) -> np.ndarray:
    out = np.empty(length, dtype=np.uint8)
    lib().sb_textgen_host(out.ctypes.data, off, length, sigma, seed)
    return out


class pinned_buffer:
    """Pinned host memory (cudaHostAlloc) exposed as a numpy uint8 array."""

    def __init__(self, nbytes: int):
        self.ptr = lib().sb_host_alloc(nbytes)
        if not self.ptr:
            raise SmartB200Error(_err())
        self.array = np.ctypeslib.as_array((C.c_ubyte * nbytes).from_address(self.ptr))

    def free(self):
        if self.ptr:
            self.array = None
            lib().sb_host_free(self.ptr)
            self.ptr = None
