#!/usr/bin/env python
"""bench.py — text GB/s scanned (HBM-resident) on B200, the metric of BASELINE.json.

Headline workload (BASELINE.json configs[2], the "1xB200 roofline run"): a synthetic uniform sigma=256 text of
4 GiB PER GPU, resident in HBM, generated on the device by the counter-based generator (textgen-style raw
byte values); one STEP = one pass of the search hot path for every (algorithm, m) in
{bf, kmp, hor, so} x {4, 16, 64, 256} = 16 searches = 64 GiB of text scanned per GPU per step, submitted
asynchronously (sb_search64_submit) and collected at the end of the step.  Patterns are cut out of the text at
seeded offsets (so every search must find >= 1 occurrence, smart.c:148-158,330).  With N GPUs (one process per
GPU under torchrun) the global text is N x 4 GiB, sharded with an m-1 halo; per-part counts are summed; there is
no collective on the data path ("scaling": "weak").

Second leg on the same JSON line, `config4` (BASELINE.json configs[3], the north-star multi-GPU shape): ONE
64 GiB sigma=4 text, 64/N GiB per rank, m in {8, 64, 1024}, STRONG scaling (same text at every N), byte kernels
(SB_FLAG_NO_PACK|SB_FLAG_NO_SKIP: the read-once roofline) and default kernels reported separately, time = max
over ranks, counts summed over ranks; plus `config4.inprocess`: the same text through the single-box partitioner
(sb_text_generate over devices 0..N-1 in ONE process, rank 0) with its counts checked against the per-rank sums.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
  python bench.py --impl reference [...]                          the reference's own CPU search() (oracle/_ref)

Prints ONE JSON line (rank 0).  `value` is timed with the text already in HBM; `e2e` runs the same searches
through the host-buffer entry the legacy plugins use (sb_search_host), host->device copy of the text inside
the timed region; `roofline` is the dominant scan kernel against the measured HBM bandwidth;
`cpu_baseline` is the reference's search() of the same algorithms on this box's host cores, with the
reference's fastest CPU code (EPSM, SSE4.2) beside it as `cpu_baseline.best`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GIB = 1 << 30
METRIC = "text GB/s scanned (HBM-resident)"
UNIT = "GB/s"
SEED = 20261018
FILTER_NAMES = {0: "scan_swar_kernel", 1: "scan_window8_kernel", 4: "scan_block_kernel<4>",
                8: "scan_block_kernel<8>", 16: "scan_block_kernel<16>", 32: "scan_block_kernel<32>",
                108: "scan_skip_kernel<8>", 116: "scan_skip_kernel<16>", 132: "scan_skip_kernel<32>",
                2: "scan_window12_kernel", 3: "scan_window16_kernel", 408: "scan_block_kernel<4,8>",
                416: "scan_block_kernel<4,16>", 816: "scan_block_kernel<8,16>",
                **{20 + L: f"scan_block_kernel<2,{L}>" for L in range(3, 9)},
                **{300 + K: f"scan_packed_kernel<{K}>" for K in (1, 2, 4, 8)},
                311: "scan_packed_half_kernel<1>", 312: "scan_packed_half_kernel<2>",
                **{500 + k: f"scan_pswar_kernel<1,{k}>" for k in range(1, 33)},
                **{540 + k: f"scan_pswar_kernel<2,{k}>" for k in range(1, 17)}}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gib-per-gpu", type=float, default=4.0)
    ap.add_argument("--sigma", type=int, default=256)
    ap.add_argument("--ms", default="4,16,64,256")
    ap.add_argument("--algos", default="bf,kmp,hor,so")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-sample-mib", type=int, default=128)
    ap.add_argument("--flags", type=int, default=0, help="SB_FLAG_* tuning overrides passed to every search")
    ap.add_argument("--no-config4", action="store_true", help="skip the 64 GiB sigma=4 strong-scaling leg")
    ap.add_argument("--c4-gib", type=float, default=64.0, help="total text size of the config4 leg")
    ap.add_argument("--c4-reps", type=int, default=10)
    return ap.parse_args()


def patterns_for(n_total, sigma, ms):
    """Same patterns on every rank: cut from the GLOBAL text via the host mirror of the generator."""
    import oracle_lib

    rng = np.random.default_rng(SEED)
    pats = {}
    for m in ms:
        off = int(rng.integers(0, n_total - m))
        pats[m] = oracle_lib.textgen_np(off, m, sigma, SEED)
    return pats


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic(kernel_name):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel_name)
        except Exception:
            return None
    return None


# --------------------------------------------------------------------------------------------- CPU legs

_CPU = {}


def _cpu_worker(job):
    algo, ci, m = job
    import oracle_lib

    st = _CPU
    buf, n = st["chunks"][ci]
    if st["kind"] == "port":  # oracle/_ref absent (fresh checkout on a box without the reference): the C restatement
        ora = st["refs"].get("_oracle")
        if ora is None:
            ora = st["refs"]["_oracle"] = oracle_lib.Oracle()
        return ora.search(algo, st["pats"][m], buf[:n])
    ref = st["refs"].get(algo)
    if ref is None:
        ref = st["refs"][algo] = oracle_lib.RefAlgo(algo)
    return ref.search_padded(st["pats"][m], buf, n)


def cpu_reference_setup(sample_bytes, sigma, ms, cores, n_total):
    """Text sample = first `sample_bytes` of the workload text, cut into one chunk per core with an
    (max m)-1 overlap so the chunks can be searched independently.  Built with the numpy
    mirror of the generator: nothing on the CPU legs loads libsmartb200.so."""
    import oracle_lib

    y = oracle_lib.textgen_np(0, sample_bytes, sigma, SEED)
    pats = {m: oracle_lib.textgen_np(1000 + 7 * m, m, sigma, SEED) for m in ms}  # cut from inside the sample
    mmax = max(ms)
    per = (sample_bytes + cores - 1) // cores
    chunks = []
    for c in range(cores):
        lo = c * per
        hi = min(sample_bytes, lo + per + mmax - 1)
        if lo >= sample_bytes:
            break
        buf = np.zeros(hi - lo + oracle_lib.RefAlgo.SLACK, dtype=np.uint8)  # calloc: 16-byte aligned (epsm.c reads __m128i)
        buf[: hi - lo] = y[lo:hi]
        chunks.append((buf, hi - lo))
    _CPU.update(refs={}, chunks=chunks, pats=pats, kind="reference" if oracle_lib.ref_available() else "port")
    return len(chunks)


def cpu_reference_step(pool, algos, ms, nchunks, must_find=True):
    """One step of the workload on the CPU sample; returns seconds."""
    t0 = time.perf_counter()
    for a in algos:
        for m in ms:
            res = pool.map(_cpu_worker, [(a, ci, m) for ci in range(nchunks)])
            if must_find and sum(res) == 0:
                raise RuntimeError(f"reference {a} found no occurrence of a text-extracted pattern (m={m})")
    return time.perf_counter() - t0


def run_cpu_reference(args, algos, ms, steps, warmup, n_total, with_best=True):
    import multiprocessing as mp

    import oracle_lib

    cores = len(os.sched_getaffinity(0))
    sample = args.cpu_sample_mib << 20
    nchunks = cpu_reference_setup(sample, args.sigma, ms, cores, n_total)
    ctx = mp.get_context("fork")
    best = None
    with ctx.Pool(processes=nchunks) as pool:
        for _ in range(warmup):
            cpu_reference_step(pool, algos, ms, nchunks)
        t = 0.0
        for _ in range(steps):
            t += cpu_reference_step(pool, algos, ms, nchunks)
        # the reference's fastest CPU code on the same sample and lengths: EPSM,
        # SSE4.2 (epsm.c:341-419, search16 :225-338).  A speed baseline only: for m >= 16 it misses an occurrence
        # that ends on the last text byte (epsm.c:330), so its count is not checked.
        if with_best and _CPU["kind"] == "reference" and oracle_lib.ref_available("epsm"):
            try:
                cpu_reference_step(pool, ["epsm"], ms, nchunks, must_find=False)
                reps = 10  # EPSM runs at tens of GB/s: one pass of the sample is only a few milliseconds
                tb = sum(cpu_reference_step(pool, ["epsm"], ms, nchunks, must_find=False) for _ in range(reps)) / reps
                best = {"algo": "epsm", "value": sample * len(ms) / tb / 1e9, "unit": UNIT, "cores": nchunks,
                        "m": list(ms), "seconds": tb, "reps": reps,
                        "note": "reference src/algos/epsm.c (SSE4.2), same sample, one search per m; speed only "
                                "(drops an occurrence ending on the last text byte for m >= 16, epsm.c:330)"}
            except Exception as e:  # reported, never required
                best = {"algo": "epsm", "value": None, "error": repr(e)}
    bytes_per_step = sample * len(algos) * len(ms)
    gbs = bytes_per_step * steps / t / 1e9
    return {"value": gbs, "unit": UNIT, "cores": nchunks, "kind": _CPU["kind"], "best": best,
            "sample": f"first {args.cpu_sample_mib} MiB of the workload text, {len(algos)}x{len(ms)} searches per step "
                      f"({','.join(algos)} x m={','.join(map(str, ms))}), one chunk per core with m-1 overlap, "
                      + ("reference search() compiled -O3 -msse4 (oracle/_ref)" if _CPU["kind"] == "reference"
                         else "oracle/smart_oracle.c restatement (oracle/_ref not built)") + f", {steps} step(s)",
            "seconds": t, "ms_per_step": 1e3 * t / steps}


# --------------------------------------------------------------------------------------------- main legs

def main_reference(args, algos, ms):
    info_rank = int(os.environ.get("RANK", "0"))
    if info_rank != 0:
        return 0
    n_total = int(args.gib_per_gpu * GIB) * args.gpus
    cb = run_cpu_reference(args, algos, ms, args.steps, max(1, min(args.warmup, 2)), n_total)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, algos, ms, n_total),
        "cpu_baseline": {"value": cb["value"], "unit": UNIT, "cores": cb["cores"], "kind": cb["kind"],
                         "sample": cb["sample"], "best": cb["best"]},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print_json(line)
    return 0


def workload_config(args, algos, ms, n_total):
    return {"workload": f"BASELINE configs[2]: synthetic uniform rand{args.sigma} text, "
                        f"{args.gib_per_gpu:g} GiB per GPU resident in HBM, m in {{{','.join(map(str, ms))}}}",
            "algorithms": algos, "m": ms, "sigma": args.sigma, "text_bytes_total": n_total,
            "searches_per_step": len(algos) * len(ms), "patterns": "cut from the text at seeded offsets",
            "parallelism": f"text sharded over {args.gpus} GPU(s), m-1 halo, counts summed, no collective",
            "l2": "inputs (4 GiB per GPU) larger than L2 (126 MB); no flush needed"}


def main_ours(args, algos, ms):
    import torch

    import smart_b200 as sb
    from smart_b200 import ranks

    info = ranks.rank_info()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the scan path has no CPU fallback")
    torch.cuda.set_device(info.local_rank)
    device = f"cuda:{info.local_rank}"
    dist = ranks.init_process_group("nccl")
    cpu_group = ranks.cpu_group(dist)  # host-side barrier for the legs where only rank 0 works
    world = info.world
    if world != args.gpus and info.rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)

    n_part = int(args.gib_per_gpu * GIB)
    n_total = n_part * world
    text = sb.Text.generate_part(n_total, args.sigma, SEED, world, info.rank, info.local_rank)
    base, own, ln = sb.shard_range(n_total, world, info.rank)
    pats = patterns_for(n_total, args.sigma, ms)

    def step(extra_flags=0):
        """all searches of the step are submitted before the first result is waited for"""
        tickets = [sb.submit(a, text, pats[m], flags=args.flags | extra_flags) for a in algos for m in ms]
        cnt, kms, launches = 0, {}, 0
        for tk in tickets:
            r = sb.collect(tk)
            if r.count < 0:
                raise RuntimeError("algorithm not applicable to a bench pattern length")
            tm = sb.last_timing()
            cnt += r.count
            launches += tm.kernels_launched
            k = kms.setdefault(tm.filter_kind, [0.0, 0])
            k[0] += tm.kernel_ms
            k[1] += tm.kernels_launched
        return cnt, kms, launches

    for _ in range(max(args.warmup, 3)):
        first_counts, _, _ = step()

    sampler = None
    if info.rank == 0:
        sampler = ClockSampler(info.local_rank)
        sampler.start()
        time.sleep(0.3)
    ranks.barrier(dist, device)
    torch.cuda.synchronize()
    sb.stream_timer_begin(info.local_rank)  # CUDA event on the stream the scan kernels are launched on
    t0 = time.perf_counter()
    total_cnt, launches, kernel_ms = 0, 0, {}
    for _ in range(args.steps):
        c, kms, l = step()
        total_cnt += c
        launches += l
        for f, (ms_, n_) in kms.items():
            k = kernel_ms.setdefault(f, [0.0, 0])
            k[0] += ms_
            k[1] += n_
    event_ms = sb.stream_timer_end_ms(info.local_rank)  # records the end event and waits for it
    torch.cuda.synchronize()
    host_ms = (time.perf_counter() - t0) * 1e3
    elapsed_ms = event_ms
    ranks.barrier(dist, device)
    if sampler:
        time.sleep(0.2)
        sampler.stop()
    kernel_total = sum(v[0] for v in kernel_ms.values())
    sum_cnt, max_ms, max_kernel_ms = ranks.reduce_step(dist, device, total_cnt, elapsed_ms, kernel_total)
    _, max_host_ms, _ = ranks.reduce_step(dist, device, 0, host_ms, 0.0)

    # ---- the same step with every text byte read by every search (SB_FLAG_NO_SKIP: hor m=256 otherwise samples
    # one sector per stride and reads ~0.29 n), timed the same way: the strict read-once figure next to `value`
    step(sb.FLAG_NO_SKIP)
    ranks.barrier(dist, device)
    torch.cuda.synchronize()
    sb.stream_timer_begin(info.local_rank)
    for _ in range(args.steps):
        step(sb.FLAG_NO_SKIP)
    strict_ms = sb.stream_timer_end_ms(info.local_rank)
    _, max_strict_ms, _ = ranks.reduce_step(dist, device, 0, strict_ms, 0.0)

    # sanity outside the timed region: every text-extracted pattern is found at least once globally and
    # all algorithms agree on the count for the same m
    per_m = {}
    for a in algos:
        for m in ms:
            c, _, _ = ranks.reduce_step(dist, device, sb.search(a, text, pats[m], flags=args.flags).count, 0, 0)
            per_m.setdefault(m, set()).add(c)
    for m, s in per_m.items():
        if len(s) != 1 or min(s) < 1:
            raise RuntimeError(f"count mismatch across algorithms or missing occurrence for m={m}: {s}")

    bytes_per_step_rank = own * len(algos) * len(ms)
    bytes_per_step = n_total * len(algos) * len(ms)
    value = bytes_per_step * args.steps / (max_ms * 1e-3) / 1e9

    # ---- per-m kernel-only rates (CUDA events on the launching stream), rank 0's part
    per_m_rates = {}
    for m in ms:
        ts = []
        for _ in range(5):
            sb.search("bf", text, pats[m], flags=args.flags)
            ts.append(sb.last_timing().kernel_ms)
        per_m_rates[str(m)] = {"kernel": FILTER_NAMES.get(sb.last_timing().filter_kind, "?"),
                               "GBps_kernel_median": own / (float(np.median(ts)) * 1e-3) / 1e9,
                               "GBps_kernel_best": own / (min(ts) * 1e-3) / 1e9}

    # ---- the algorithm axis: the same searches with each algorithm's OWN table driving the scan on the device
    # (SB_FLAG_TABLE_SCAN: kmpNext / hbc / Shift-Or masks in shared memory, table_kernels.cuh); bf has no table
    table_scan = {}
    for a in list(algos) + [b for b in ("bm", "kr", "hash3", "qf23") if b not in algos]:  # (+ the other SURVEY 8(a) families)
        if a == "bf":
            continue
        try:
            row = {}
            for m in ms:
                r = sb.search(a, text, pats[m], flags=sb.FLAG_TABLE_SCAN)
                tm = sb.last_timing()
                if tm.filter_kind in range(601, 652) and r.count == sb.search(a, text, pats[m]).count:
                    row[str(m)] = own / (tm.kernel_ms * 1e-3) / 1e9
            table_scan[a] = row
        except Exception as e:  # an extra, never fatal
            table_scan[a] = {"error": repr(e)}

    # ---- (f)-3 extra: the driver's real shape, 500 patterns of one length, counted in ONE pass
    multi = None
    try:
        import oracle_lib

        mp = 16
        rngp = np.random.default_rng(SEED + 1)
        offs = rngp.integers(0, n_total - mp, 500)
        pats500 = np.stack([oracle_lib.textgen_np(int(o), mp, args.sigma, SEED) for o in offs])
        sb.search_multi("bf", text, pats500)
        ranks.barrier(dist, device)
        tm0 = time.perf_counter()
        cnts, _, _ = sb.search_multi("bf", text, pats500)
        dt = time.perf_counter() - tm0
        _, dt_ms, _ = ranks.reduce_step(dist, device, 0, dt * 1e3, 0.0)
        multi = {"patterns": 500, "m": mp, "ms": dt_ms, "effective_text_GBps": 500 * n_total / (dt_ms * 1e-3) / 1e9,
                 "kernel": "scan_multi_kernel<%d>" % (sb.last_timing().filter_kind - 200),
                 "note": "sb_search_multi64: one pass of the resident text for the whole pattern set"}
    except Exception as e:  # an extra, never fatal
        multi = {"error": repr(e)}

    # ---- roofline of the dominant kernel (largest share of device time inside the timed region)
    peak, peak_src = measured_peak()
    # the sub-linear scan_skip kernels read 32 of every `stride` bytes, so n/t is not bounded by the
    # read-once roofline; the roofline object describes the dominant READ-EVERY-BYTE kernel
    sublinear = lambda f: 100 < f < 200 or f >= 300  # scan_skip<G>, scan_packed*: they read less than the text
    linear = {f: v for f, v in kernel_ms.items() if not sublinear(f)} or kernel_ms
    dom = max(linear.items(), key=lambda kv: kv[1][0])
    dom_name = FILTER_NAMES.get(dom[0], str(dom[0]))
    avg_ms = dom[1][0] / max(1, dom[1][1])
    achieved = own / (avg_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "frac_of_nominal_8TBps": achieved / 8000.0,
                "traffic": recorded_traffic(dom_name), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": own, "avg_launch_ms": avg_ms,
                "share_of_step_device_time": dom[1][0] / max(kernel_total, 1e-9),
                "kernels": {FILTER_NAMES.get(f, str(f)): {"launches": v[1], "avg_ms": v[0] / max(1, v[1]),
                                                          "GBps": own / (v[0] / max(1, v[1]) * 1e-3) / 1e9,
                                                          "sublinear": sublinear(f)}
                            for f, v in kernel_ms.items()},
                "note": "scan_skip kernels sample one 32-byte sector per stride (stride ~ m-7): their text GB/s "
                        "may exceed the read-once peak; see profiles/ for their DRAM traffic"}

    # ---- e2e: the same searches through the host-buffer entry (legacy plugin path), H2D inside the timer
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(sb, ranks, dist, device, text, base, own, n_total, pats, algos, ms, args)
    text.free()

    # ---- config4: ONE 64 GiB sigma=4 text, strong scaling (BASELINE.json configs[3])
    config4 = None
    if not args.no_config4:
        try:
            config4 = run_config4(sb, ranks, dist, cpu_group, device, info, world, args)
        except Exception as e:  # reported on the line, never silently dropped
            config4 = {"error": repr(e)}

    # ---- CPU baseline (rank 0, at every N): reference search() on this box's host cores; the other ranks wait
    # on the host-side barrier (no GPU kernel spins meanwhile)
    cpu = None
    if info.rank == 0 and not args.no_cpu:
        try:
            cb = run_cpu_reference(args, algos, ms, 1, 0, n_total)
            cpu = {"value": cb["value"], "unit": UNIT, "cores": cb["cores"], "kind": cb["kind"], "sample": cb["sample"],
                   "best": cb["best"]}
        except Exception as e:  # the baseline is reported, never required
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(e)}
    ranks.host_barrier(cpu_group)

    if info.rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": max_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, algos, ms, n_total),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches * world,
            "clocks": sampler.summary() if sampler else None,
            "value_kernels_only": bytes_per_step_rank * args.steps / (max_kernel_ms * 1e-3) / 1e9 * world,
            "value_host_clock": bytes_per_step * args.steps / (max_host_ms * 1e-3) / 1e9,
            "value_read_every_byte": bytes_per_step * args.steps / (max_strict_ms * 1e-3) / 1e9,
            "per_m": per_m_rates, "occurrences_per_step": sum_cnt // args.steps, "multi_pattern": multi,
            "table_scan_GBps": table_scan,
            "config4": config4,
            "timing": "CUDA events on the library's launching stream bracketing the K timed steps (between "
                      "barrier+synchronize pairs), max over ranks; every step submits its 16 searches asynchronously "
                      "and collects them (sb_search64_submit / sb_search_collect); value_host_clock = same region by "
                      "the host clock; value_kernels_only = sum of per-kernel event times; value_read_every_byte = the same "
                      "step with SB_FLAG_NO_SKIP (no sub-linear sampling: every search reads all n bytes)",
        }
        print_json(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


C4_SEED = 64


def run_config4(sb, ranks, dist, cpu_group, device, info, world, args):
    """BASELINE.json configs[3]: ONE sigma=4 text of --c4-gib GiB (default 64), part r of N on rank r (m-1 halo),
    m in {8, 64, 1024}; strong scaling: the text is the same at every N.  Every (mode, algorithm, m) is timed as
    --c4-reps scans submitted back to back (CUDA events on the scan stream, max over ranks)."""
    import torch

    import oracle_lib

    n = int(args.c4_gib * GIB)
    sigma, ms4, reps = 4, (8, 64, 1024), max(1, args.c4_reps)
    pats = {m: oracle_lib.textgen_np(n // 2 - m // 2, m, sigma, C4_SEED) for m in ms4}  # straddle a seam for even N
    modes = (("byte_kernels", sb.FLAG_NO_PACK | sb.FLAG_NO_SKIP), ("default", 0))
    combos = [(mode, fl, a, m) for mode, fl in modes for a in ("bf", "hor") for m in ms4]
    peak, _ = measured_peak()

    t0 = time.perf_counter()
    text = sb.Text.generate_part(n, sigma, C4_SEED, world, info.rank, info.local_rank)
    gen_s = time.perf_counter() - t0
    results, counts = [], {}
    try:
        for mode, fl, a, m in combos:
            sb.search(a, text, pats[m], flags=fl)  # warm-up (plan and aux block cached, clocks up)
            ranks.barrier(dist, device)
            torch.cuda.synchronize()
            sb.stream_timer_begin(info.local_rank)
            tickets = [sb.submit(a, text, pats[m], flags=fl) for _ in range(reps)]
            got = [sb.collect(tk) for tk in tickets]
            tm = sb.last_timing()
            ev_ms = sb.stream_timer_end_ms(info.local_rank)
            if len({r.count for r in got}) != 1:
                raise RuntimeError(f"config4: repeated scans disagree ({a}, m={m})")
            cnt, max_ms, max_k = ranks.reduce_step(dist, device, got[0].count, ev_ms, float(np.median([r.run_ms for r in got])))
            counts.setdefault(m, set()).add(cnt)
            gbs = n * reps / (max_ms * 1e-3) / 1e9
            results.append({"mode": mode, "algo": a, "m": m, "count": cnt, "ms_per_scan": max_ms / reps,
                            "kernel_ms_median_max_over_ranks": max_k, "text_GBps": gbs,
                            "frac_of_measured_hbm": gbs / (peak * world), "frac_of_nominal_8TBps": gbs / (8000.0 * world),
                            "kernel": FILTER_NAMES.get(tm.filter_kind, str(tm.filter_kind))})
    finally:
        text.free()
    for m, cs in counts.items():
        if len(cs) != 1 or min(cs) < 1:
            raise RuntimeError(f"config4: modes/algorithms disagree on the count for m={m}: {cs}")

    # ---- the same text through the in-process partitioner, one process driving all N devices (rank 0)
    ranks.host_barrier(cpu_group)  # every rank has freed its part
    inproc = None
    if info.rank == 0:
        try:
            t0 = time.perf_counter()
            with sb.Text.generate(n, sigma, C4_SEED, devices=list(range(world))) as whole:
                gen2_s = time.perf_counter() - t0
                rows = []
                for mode, fl, a, m in combos:
                    sb.search(a, whole, pats[m], flags=fl)
                    t1 = time.perf_counter()
                    tickets = [sb.submit(a, whole, pats[m], flags=fl) for _ in range(reps)]
                    got = [sb.collect(tk) for tk in tickets]
                    wall = time.perf_counter() - t1
                    rows.append({"mode": mode, "algo": a, "m": m, "count": got[0].count,
                                 "same_as_per_rank_sum": {got[0].count} == counts[m],
                                 "wall_ms_per_scan": wall * 1e3 / reps,
                                 "kernel_ms_max_over_devices_median": float(np.median([r.run_ms for r in got])),
                                 "text_GBps_wall": n * reps / wall / 1e9})
                inproc = {"shards": whole.shards, "text_generation_s": round(gen2_s, 3), "results": rows,
                          "api": "sb_text_generate(devices=0..N-1) + sb_search64_submit/collect in ONE process: "
                                 "one scan stream per device, per-shard counts summed on the host"}
                if not all(r["same_as_per_rank_sum"] for r in rows):
                    raise RuntimeError("config4: in-process partitioner and per-rank parts disagree")
        except Exception as e:
            inproc = {"error": repr(e)}
    ranks.host_barrier(cpu_group)
    return {"workload": f"BASELINE configs[3]: ONE synthetic uniform rand{sigma} text of {args.c4_gib:g} GiB, "
                        f"{args.c4_gib / world:g} GiB per GPU with an m-1 halo, m in {{8,64,1024}}, bf and hor",
            "scaling": "strong", "n_gpus": world, "text_bytes_total": n, "seed": C4_SEED, "reps": reps,
            "text_generation_s": round(gen_s, 3),
            "modes": {"byte_kernels": "SB_FLAG_NO_PACK|SB_FLAG_NO_SKIP: every text byte is read once (the HBM "
                                      "roofline applies)",
                      "default": "the library's own choice (2-bit packed copy / sampling kernels: less than n bytes read)"},
            "results": results, "inprocess": inproc}


def run_e2e(sb, ranks, dist, device, text, base, own, n_total, pats, algos, ms, args):
    """End to end with HOST buffers: this rank's part of the text sits in pinned host memory and every timed step
    copies it to the device inside the timed region.  Two forms of the same 16 searches:
      value     sb_search_host_batch - ONE call for the step, the text crosses PCIe once and every 64 MiB chunk is
                scanned by all 16 jobs (the shape of the driver's loop: many searches over one text segment);
      per_call  sb_search_host - the body of the legacy per-algorithm search() plugin, one call and one full copy
                of the text PER SEARCH (what an unmodified SMART driver without the resident worker pays)."""
    import torch

    mmax = max(ms)
    ln = min(n_total - base, own + mmax - 1)
    sb.set_device(int(str(device).split(":")[1]) if ":" in str(device) else 0)  # this rank's GPU and PCIe link
    pin = sb.pinned_buffer(ln)
    try:
        chunk = 256 << 20
        for off in range(0, ln, chunk):  # fill the pinned buffer with this rank's part of the text
            k = min(chunk, ln - off)
            pin.array[off:off + k] = text.download(base + off, k)
        steps = max(1, min(args.steps, 3))
        nsearch = len(algos) * len(ms)
        world = 1 if dist is None else dist.get_world_size()
        jobs = [(a, pats[m], args.flags) for a in algos for m in ms]

        def step_batch():
            # the buffer is this rank's part + the halo of the longest pattern; every job counts the part's own starts
            res = sb.search_host_batch(jobs, pin.array[:ln], starts=own)
            return sum(r.count for r in res)

        def step_calls():
            c = 0
            for a in algos:
                for m in ms:
                    hi = min(ln, own + m - 1)
                    c += sb.search_host(a, pats[m], pin.array[:hi], flags=args.flags).count
            return c

        def timed(step):
            step()
            ranks.barrier(dist, device)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(steps):
                step()
            torch.cuda.synchronize()
            ms_el = (time.perf_counter() - t0) * 1e3
            ranks.barrier(dist, device)
            _, max_ms, _ = ranks.reduce_step(dist, device, 0, ms_el, 0)
            return n_total * nsearch * steps / (max_ms * 1e-3) / 1e9

        if step_batch() != step_calls():
            raise RuntimeError("e2e: the batch entry and the per-search entry disagree")
        v_batch = timed(step_batch)
        v_calls = timed(step_calls)
        return {"value": v_batch, "unit": UNIT,
                "h2d_bytes_per_step": int((ln + 4352 * nsearch) * world),
                "d2h_bytes_per_step": 32 * nsearch * world, "steps": steps,
                "api": "sb_search_host_batch: pinned host text, ONE call per step for the 16 searches; 64 MiB chunks, "
                       "each copied once and scanned by every job, copy of chunk k+1 overlapped with the scans of chunk k",
                "per_call": {"value": v_calls, "unit": UNIT,
                             "h2d_bytes_per_step": int(sum(min(ln, own + m - 1) + 4352 for m in ms) * len(algos) * world),
                             "api": "sb_search_host (the body of the legacy per-algorithm search() plugins): one call "
                                    "and one copy of the text per search"}}
    finally:
        pin.free()


def ensure_built(impl):
    """Binaries are git-ignored; a fresh checkout builds them (nvcc/gcc are in the image). Not a fallback:
    a failed build stops the bench.  The reference arm needs (and loads) only the checker under oracle/."""
    if impl != "reference":
        import smart_b200

        if not os.path.exists(smart_b200.library_path()):
            smart_b200.build_library()
    if not os.path.exists(os.path.join(ROOT, "oracle", "libsmart_oracle.so")):
        subprocess.run(["make", "-j8", "-C", os.path.join(ROOT, "oracle"), "oracle", "ref"], check=True, capture_output=True)


def main():
    args = parse_args()
    ms = [int(v) for v in args.ms.split(",")]
    algos = args.algos.split(",")
    # stdout must carry exactly ONE JSON line: libraries (e.g. the NCCL version banner) print to fd 1, so
    # everything else is sent to stderr and the JSON line is written to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    global print_json
    print_json = lambda obj: os.write(real_stdout, (json.dumps(obj) + "\n").encode())
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        ensure_built(args.impl)
    elif args.impl != "reference":  # the other ranks wait for local rank 0's build (normally it is prebuilt)
        import smart_b200

        for _ in range(600):
            if os.path.exists(smart_b200.library_path()):
                break
            time.sleep(1.0)
    if args.impl == "reference":
        return main_reference(args, algos, ms)
    return main_ours(args, algos, ms)


print_json = lambda obj: print(json.dumps(obj))


if __name__ == "__main__":
    sys.exit(main())
