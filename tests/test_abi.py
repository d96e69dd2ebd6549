"""The C-ABI library loads and exports every symbol include/smart_b200.h declares; host-only behaviour."""
import ctypes
import os
import re

import numpy as np
import pytest

import smart_b200 as sb
from smart_b200 import binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "smart_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = ctypes.CDLL(sb.library_path())
    names = header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/smart_b200.h but not exported"
    assert sorted(n for n, _, _ in binding.SYMBOLS) == names, "python binding out of sync with the header"


def test_registry_matches_reference_names():
    L = sb.lib()
    names = [L.sb_algo_name(i).decode() for i in range(L.sb_algo_count())]
    assert tuple(names) == sb.ALGORITHMS
    assert L.sb_algo_id(b"KMP") == names.index("kmp")  # smart.c:142 lower-cases registry names
    assert L.sb_algo_id(b"nope") == -1
    assert L.sb_algo_name(999) is None


def test_textgen_host_is_counter_based():
    a = sb.textgen_host(0, 4096, 4, 42)
    b = sb.textgen_host(1000, 100, 4, 42)
    assert np.array_equal(a[1000:1100], b)
    assert a.max() <= 3 and set(np.unique(a)) == {0, 1, 2, 3}
    assert sb.textgen_host(0, 4096, 2, 42).max() == 1
    assert sb.textgen_host(0, 65536, 250, 42).max() == 249
    full = sb.textgen_host(0, 1 << 16, 256, 7)
    assert len(np.unique(full)) == 256
    assert not np.array_equal(full, sb.textgen_host(0, 1 << 16, 256, 8))


def test_shard_ranges_partition_the_text():
    for n in (0, 1, 4095, 4096, 10_000_000, (1 << 32) + 12345):
        for parts in (1, 2, 3, 8):
            nxt = 0
            for g in range(parts):
                base, own, ln = sb.shard_range(n, parts, g)
                assert base == nxt and base % 4096 == 0 or own == 0
                assert ln == min(n - base, own + 4199)
                nxt = base + own
            assert nxt == n


def test_no_gpu_fails_loudly():
    """Without a CUDA device the product must raise, never fall back to a CPU path."""
    if sb.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(sb.SmartB200Error):
        sb.Text.upload(b"abcabc")
    with pytest.raises(sb.SmartB200Error):
        sb.search_host("bf", b"abc", b"abcabc")
    assert sb.plugin_search("bf", b"abc", b"abcabc") == -1


def test_product_does_not_link_the_oracle():
    import subprocess

    out = subprocess.run(["ldd", sb.library_path()], capture_output=True, text=True).stdout
    assert "oracle" not in out
    nm = subprocess.run(["nm", "-D", sb.library_path()], capture_output=True, text=True).stdout
    assert "ora_" not in nm
    for root, _, files in os.walk(os.path.join(ROOT, "smart_b200")):
        for f in files:
            if f.endswith((".c", ".cu", ".cuh", ".h", ".py")):
                txt = open(os.path.join(root, f), errors="replace").read()
                assert "smart_oracle" not in txt and "oracle_lib" not in txt, f
