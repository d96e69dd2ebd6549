"""bench.py's JSON contract: the reference arm runs on the CPU here; our arm is checked on the GPU box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def run_bench(*args, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                       timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


def test_reference_arm_json(monkeypatch):
    # the reference arm must not need (or map) the product library: point it at a path that does not exist
    monkeypatch.setenv("SMART_B200_LIB", "/nonexistent/libsmartb200.so")
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sample-mib", "8")
    assert BASE_KEYS <= set(d)
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True
    assert d["metric"] == "text GB/s scanned (HBM-resident)" and d["dtype"] == "u8" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    best = d["cpu_baseline"]["best"]  # the reference's fastest CPU code (EPSM) beside the same-algorithm number
    if d["cpu_baseline"]["kind"] == "reference":
        assert best["algo"] == "epsm" and best["value"] > 0


def test_reference_arm_other_ranks_exit_quietly():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=60, cwd=ROOT, env={**os.environ, "RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.gpu
def test_our_arm_json_small():
    d = run_bench("--gib-per-gpu", "0.25", "--steps", "2", "--warmup", "3", "--cpu-sample-mib", "8", "--c4-gib", "1",
                  "--c4-reps", "3")
    assert BASE_KEYS | {"roofline", "clocks"} <= set(d)
    assert "impl" not in d and d["n_gpus"] == 1 and d["scaling"] == "weak" and d["data"] == "synthetic"
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert rf["kernel"].startswith("scan_") and rf["achieved"] > 100
    assert d["gpu_launches"] == 2 * 16 and d["value"] > 100
    e = d["e2e"]
    # e2e: the batch entry copies the text once per step, the per-call entry once per search
    assert e["value"] > 1 and e["h2d_bytes_per_step"] >= (1 << 28) and e["d2h_bytes_per_step"] > 0
    assert e["per_call"]["h2d_bytes_per_step"] >= 16 * (1 << 28) and 1 < e["per_call"]["value"] < e["value"]
    assert e["value"] < d["value"]  # the end-to-end path pays the host->device copy
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] > 0 and cb["cores"] >= 1 and "sample" in cb
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    c4 = d["config4"]  # the strong-scaling leg: one sigma=4 text, byte kernels and default kernels, same counts
    assert c4["scaling"] == "strong" and c4["text_bytes_total"] == 1 << 30 and len(c4["results"]) == 12
    assert {r["mode"] for r in c4["results"]} == {"byte_kernels", "default"}
    for m in (8, 64, 1024):
        assert len({r["count"] for r in c4["results"] if r["m"] == m}) == 1
    assert all(r["same_as_per_rank_sum"] for r in c4["inprocess"]["results"])
    assert d["value"] <= d["value_kernels_only"] * 1.001  # the step includes the gaps between its kernels
