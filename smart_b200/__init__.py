"""smart_b200 — B200-native exact-string-matching scan engine behind SMART's plugin surface.

The product is the C-ABI shared library ``smart_b200/lib/libsmartb200.so`` (CUDA kernels for
sm_100a + C host preprocessing, see ``include/smart_b200.h``) and the C host programs under
``smart_b200/host``.  This Python package is a thin ctypes binding used by the tests and by
``bench.py``; it adds no compute of its own and has no CPU fallback.
"""
from .binding import (  # noqa: F401
    ALGORITHMS,
    FLAG_FILTER_B4,
    FLAG_FILTER_B8,
    FLAG_FILTER_B16,
    FLAG_FILTER_B32,
    FLAG_FILTER_B4L8,
    FLAG_FILTER_B4L16,
    FLAG_FILTER_B8L16,
    FLAG_FILTER_B2,
    FLAG_NO_PACK,
    FLAG_FILTER_SHIFT,
    FLAG_NO_SKIP,
    FLAG_DENSE,
    FLAG_NO_PERIOD,
    FLAG_POSITIONS,
    NOT_APPLICABLE,
    SmartB200Error,
    Tables,
    Text,
    Timing,
    algo_id,
    applicable,
    build_library,
    device_count,
    get_device,
    set_device,
    last_timing,
    lib,
    library_path,
    pinned_buffer,
    prep_tables,
    search,
    search_host,
    search_multi,
    shard_range,
    stream_timer_begin,
    stream_timer_end_ms,
    plugin_search,
    textgen_host,
)
