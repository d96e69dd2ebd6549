"""The N>1 path of bench.py (one process per GPU) exercised on the CPU with world_size-2 gloo: the
partition rule (sb_shard_range), the per-part counting and the reduction plumbing in smart_b200/ranks.py.
The scanner here is the oracle — this checks the host logic, not the kernels."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, m, out_q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    import smart_b200 as sb
    from smart_b200 import ranks

    dist = ranks.init_process_group("gloo")
    info = ranks.rank_info()
    assert (info.rank, info.world) == (rank, world)
    ora = oracle_lib.Oracle()
    base, own, ln = sb.shard_range(n, world, rank)
    # this rank regenerates only its own part (+ halo) of the global text
    part = sb.textgen_host(base, ln, 4, 77)
    x = sb.textgen_host(n // 2 - m // 2, m, 4, 77)  # straddles the boundary between the two parts
    hi = max(0, min(own, n - m + 1 - base))          # owned start positions
    local = ora.bf_count(x, part[: min(ln, hi + m - 1)]) if hi > 0 else 0
    ranks.barrier(dist, "cpu")
    ranks.host_barrier(ranks.cpu_group(dist))  # the host-side barrier bench.py uses around rank-0-only legs
    total, max_ms, max_k = ranks.reduce_step(dist, "cpu", local, 10.0 * (rank + 1), 1.0 + rank)
    out_q.put((rank, local, total, max_ms, max_k))
    dist.destroy_process_group()


@pytest.mark.parametrize("m", [2, 16, 1024])
def test_two_ranks_sum_to_whole(oracle, m):
    import torch.multiprocessing as mp

    import smart_b200 as sb

    n, world = 3_000_000, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, m, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole = sb.textgen_host(0, n, 4, 77)
    x = sb.textgen_host(n // 2 - m // 2, m, 4, 77)
    want = oracle.bf_count(x, whole)
    assert want >= 1
    assert all(r[2] == want for r in res)            # every rank sees the summed count
    assert sum(r[1] for r in res) == want            # no occurrence lost or double counted at the seam
    assert all(r[3] == 20.0 and r[4] == 2.0 for r in res)  # max over ranks


def test_single_rank_is_identity():
    from smart_b200 import ranks

    assert ranks.reduce_step(None, "cpu", 5, 1.5, 0.5) == (5, 1.5, 0.5)
    ranks.barrier(None)
    ranks.host_barrier(ranks.cpu_group(None))


def test_numpy_text_generator_equals_the_library():
    """bench.py's CPU legs build the workload text with oracle_lib.textgen_np so that the reference arm never
    loads libsmartb200.so; it must be the same stream as sb_textgen_host / the device generator."""
    import oracle_lib
    import smart_b200 as sb

    for sigma in (256, 4, 2, 16, 3, 250, 1):
        for off, ln in ((0, 100), (5, 77), (123457, 4096), (8, 8), (7, 1), (1 << 33, 65)):
            assert np.array_equal(oracle_lib.textgen_np(off, ln, sigma, 20261018), sb.textgen_host(off, ln, sigma, 20261018))
