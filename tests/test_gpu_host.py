"""GPU-side checks of the C host programs through the real plugin protocol (SysV shm + one process per
search): the reference's own `test` binary drives OUR per-algorithm executables, our driver runs them with
the resident worker, and the worker keeps the text in HBM across the whole loop."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin")
REF = os.path.join(ROOT, "oracle", "_ref")

pytestmark = pytest.mark.gpu


def run(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, timeout=240, **kw)


@pytest.fixture()
def worker(tmp_path):
    sock = str(tmp_path / "w.sock")
    env = {**os.environ, "SMART_B200_SOCK": sock}
    r = run([os.path.join(BIN, "smartd")], env=env)
    assert r.returncode == 0 and r.stdout.startswith("ready"), (r.stdout, r.stderr)
    assert stats(env) == (0, 0)  # fail fast if the worker is not really serving
    yield env
    run([os.path.join(BIN, "smartd"), "-stop"], env=env)


def stats(env):
    r = run([os.path.join(BIN, "smartd"), "-stats"], env=env)
    m = re.search(r"searches (\d+) text_uploads (\d+)", r.stdout)
    assert m, f"worker not reachable: {r.stdout} {r.stderr}"
    return int(m.group(1)), int(m.group(2))


def test_standalone_mode_prints_reference_format():
    # main.h:123-135: "<algo> P m T n" -> "found %d occurrences"
    for algo, want in (("bf", 4), ("kmp", 4), ("qf23", 4), ("hash5", -1)):
        r = run([os.path.join(BIN, "algos", algo), "aba", "3", "ababababab", "10"],
                env={**os.environ, "SMART_B200_NO_WORKER": "1"})
        assert r.returncode == 0 and r.stdout.strip() == f"found {want} occurrences", (algo, r.stdout, r.stderr)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "test")), reason="oracle/_ref/test not built")
def test_reference_harness_drives_our_executables(tmp_path, worker):
    """The UNMODIFIED reference harness (src/test.c) runs its 20 cases against our executables."""
    (tmp_path / "source").mkdir()
    os.symlink(os.path.join(BIN, "algos"), tmp_path / "source" / "bin")
    for a in ("bf", "kmp", "mp", "hor", "bm", "qs", "so", "sa", "kr", "tunedbm", "raita", "hash3", "hash8", "qf23", "qf63",
              "smith", "tbm", "ag", "ac", "bndm", "br", "tvsbs", "fjs", "ssabs", "ts",
              "skip", "kmpskip", "fs", "graspm",
              # second batch: a few of every family
              "aoso2", "blim", "ufndmq4", "aut", "tw", "smoa", "sbndm", "bxs4", "dbww", "bom", "ebom", "bsdm5", "wfr",
              "twfrq3", "lwfr4", "skip5", "zt", "tvsbs-w2", "fs-w1", "ssef"):
        r = run([os.path.join(REF, "test"), a], cwd=tmp_path, env=worker)
        assert "Test passed" in r.stdout and r.returncode == 0, (a, r.stdout[-300:], r.stderr[-300:])
    searches, uploads = stats(worker)
    assert searches >= 29 * 10  # not-applicable cases never reach the worker
    assert uploads >= 29  # tiny texts change with every case


def test_our_harness_strict_on_our_executables(worker):
    for a in ("bf", "kmp", "hor", "so"):
        r = run([os.path.join(BIN, "test"), a, "-strict"], env=worker)
        assert r.returncode == 0 and "20 cases, 0 skipped" in r.stdout, (a, r.stdout, r.stderr)


def test_driver_keeps_text_resident_and_matches_reference(tmp_path):
    data = tmp_path / "data" / "mini"
    data.mkdir(parents=True)
    text = open(os.path.join(ROOT, "tests", "golden", "bible64k.bin"), "rb").read()
    (data / "a.txt").write_bytes(text)
    (data / "index.txt").write_text("#a.txt#\n")
    sock = str(tmp_path / "d.sock")
    env = {**os.environ, "SMART_B200_SOCK": sock}
    common = ["-text", "mini", "-data", str(tmp_path / "data"), "-results", str(tmp_path / "results"), "-pset", "6",
              "-plen", "2", "32", "-short", "-occ", "-seed", "11"]
    ours = run([os.path.join(BIN, "smart"), *common, "-algo", "bf,kmp,hor,so,hash3,qf23"], cwd=tmp_path, env=env)
    assert ours.returncode == 0, ours.stderr
    assert "Resident GPU worker started" in ours.stdout
    assert ".[ERROR]" not in ours.stdout and ".[OUT]" not in ours.stdout
    assert ours.stdout.count(".[--]") == 2  # hash3 and qf23 at m=2 (hash3.c:31, qf23.c:44)
    occ = lambda s: re.findall(r"\] (\w+) \.+.*?occ (\d+)", s)
    got = occ(ours.stdout)
    assert len(got) == 16 * 6 - 2
    if os.path.exists(os.path.join(REF, "bin", "bf")):
        ref = run([os.path.join(BIN, "smart"), *common, "-algo", "bf,kmp,hor,so,hash3,qf23", "-bin",
                   os.path.join(REF, "bin"), "-noworker"], cwd=tmp_path, env=env)
        assert occ(ref.stdout) == got  # same seed, same patterns: identical occurrence columns
    # the worker is stopped by the driver
    r = run([os.path.join(BIN, "smartd"), "-stats"], env=env)
    assert r.returncode != 0


def test_worker_uploads_text_once_per_text(tmp_path, worker):
    data = tmp_path / "data"
    for name, blob in (("t1", b"abcabcabd" * 5000), ("t2", b"xyzxyzxyw" * 5000)):
        (data / name).mkdir(parents=True)
        (data / name / "f.txt").write_bytes(blob)
        (data / name / "index.txt").write_text("#f.txt#\n")
    r = run([os.path.join(BIN, "smart"), "-text", "t1-t2", "-data", str(data), "-pset", "5", "-plen", "4", "8",
             "-algo", "bf,kmp", "-noworker", "-seed", "3"], cwd=tmp_path, env=worker)
    assert r.returncode == 0 and r.stdout.count(".[OK]") == 2 * 2 * 2, r.stdout
    searches, uploads = stats(worker)
    assert searches == 2 * 2 * 2 * 5
    assert uploads == 2  # one per text, although the shm key and n are identical for t1 and t2


def test_driver_batch_mode_matches_per_pattern_mode(tmp_path):
    """-batch: one pass of the resident text per (algorithm, length) for the whole pattern set; the
    occurrence columns must equal the process-per-search run with the same seed."""
    data = tmp_path / "data" / "mini"
    data.mkdir(parents=True)
    text = open(os.path.join(ROOT, "tests", "golden", "bible64k.bin"), "rb").read()
    (data / "a.txt").write_bytes(text)
    (data / "index.txt").write_text("#a.txt#\n")
    env = {**os.environ, "SMART_B200_SOCK": str(tmp_path / "b.sock")}
    common = ["-text", "mini", "-data", str(tmp_path / "data"), "-pset", "40", "-plen", "2", "32", "-short", "-occ",
              "-seed", "23", "-algo", "bf,kmp,hor,hash3"]
    single = run([os.path.join(BIN, "smart"), *common], cwd=tmp_path, env=env)
    batch = run([os.path.join(BIN, "smart"), *common, "-batch"], cwd=tmp_path, env=env)
    assert single.returncode == 0 and batch.returncode == 0, (single.stderr, batch.stderr)
    occ = lambda s: re.findall(r"\] (\w+) \.+\[(\w+|--)\](?:.*?occ (\d+))?", s)
    a, b = occ(single.stdout), occ(batch.stdout)
    assert len(a) == 16 * 4 and a == b, [(x, y) for x, y in zip(a, b) if x != y][:5]
    assert sum(1 for x in b if x[1] == "--") == 1  # hash3 at m=2


def test_driver_pipeline_mode_matches_per_pattern_mode(tmp_path):
    """-pipeline: the pattern set of a length goes to the worker in ONE request (SB_REQ_SEARCH_SET) and is searched
    as independent scans submitted asynchronously (sb_search64_submit / sb_search_collect); same verdicts and
    occurrence columns as the process-per-search run with the same seed, -1 domains and a second worker included."""
    data = tmp_path / "data" / "mini"
    data.mkdir(parents=True)
    text = open(os.path.join(ROOT, "tests", "golden", "bible64k.bin"), "rb").read()
    (data / "a.txt").write_bytes(text)
    (data / "index.txt").write_text("#a.txt#\n")
    env = {**os.environ, "SMART_B200_SOCK": str(tmp_path / "p.sock")}
    def args(npat):
        return [os.path.join(BIN, "smart"), "-text", "mini", "-data", str(tmp_path / "data"), "-pset", str(npat), "-plen", "2",
                "32", "-short", "-occ", "-seed", "29", "-algo", "bf,kmp,hor,hash3,sbndm2", "-pre"]

    single = run(args(30), cwd=tmp_path, env=env)
    piped30 = run([*args(30), "-pipeline"], cwd=tmp_path, env=env)
    piped = run([*args(300), "-pipeline"], cwd=tmp_path, env=env)  # more patterns per set than tickets in flight
    assert single.returncode == 0 and piped30.returncode == 0 and piped.returncode == 0, (single.stderr, piped.stderr)
    occ = lambda s: re.findall(r"\] ([\w-]+) \.+\[(\w+|--)\](?:.*?occ (\d+))?", s)
    a, b, c = occ(single.stdout), occ(piped30.stdout), occ(piped.stdout)
    assert len(a) == 16 * 5 and a == b, [(x, y) for x, y in zip(a, b) if x != y][:5]
    assert [x[:2] for x in c] == [x[:2] for x in a] and all(x[1] in ("OK", "--") for x in c)
    assert sum(1 for x in b if x[1] == "--") == 1  # hash3 at m=2
    # a worker that is already running is shared, not replaced (smartd exits 7) and not stopped by the second driver
    started = run([os.path.join(BIN, "smartd")], env=env)
    assert started.returncode == 0 and "ready" in started.stdout
    again = run([os.path.join(BIN, "smartd")], env=env)
    assert again.returncode == 7 and "already running" in again.stderr
    shared = run([*args(30), "-pipeline"], cwd=tmp_path, env=env)
    assert "already running (shared" in shared.stdout and occ(shared.stdout) == a
    assert run([os.path.join(BIN, "smartd"), "-stats"], env=env).returncode == 0  # still alive after the driver left
    assert run([os.path.join(BIN, "smartd"), "-stop"], env=env).returncode == 0


def test_plain_c_client_of_the_abi():
    """tests/c/abi_demo.c: a C program against include/smart_b200.h only (no Python in the loop)."""
    r = run([os.path.join(BIN, "abi_demo")])
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert r.stdout.count(": ok") == 4 and "FAIL" not in r.stdout


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "smart")), reason="oracle/_ref/smart not built")
def test_reference_driver_drives_our_executables(tmp_path, worker):
    """BASELINE configs[0] end to end with the UNMODIFIED reference driver (src/smart.c): it loads the text,
    creates the shm segments, samples patterns and launches OUR executables from ./source/bin; every cell must
    come back [OK] (a count of 0 would print [ERROR], a failure [--], smart.c:376-377)."""
    (tmp_path / "source").mkdir()
    os.symlink(os.path.join(BIN, "algos"), tmp_path / "source" / "bin")
    (tmp_path / "source" / "algorithms.h").write_text("#1 #bf \n#1 #kmp \n#1 #hor \n#1 #so \n#0 #bm \n#1 #hash3 \n")
    (tmp_path / "results").mkdir()
    d = tmp_path / "data" / "englishTexts"
    d.mkdir(parents=True)
    (d / "bible.txt").write_bytes(open(os.path.join(ROOT, "tests", "golden", "bible64k.bin"), "rb").read())
    (d / "index.txt").write_text("#bible.txt#\n")
    r = run([os.path.join(REF, "smart"), "-text", "englishTexts", "-pset", "4", "-plen", "4", "16", "-short", "-occ"],
            cwd=tmp_path, env=worker)
    out = r.stdout.replace("\b", "")
    assert "Text buffer of dimension 65536 byte" in out, (out[-500:], r.stderr[-300:])
    assert out.count("[OK]") == 5 * 7, out[-1500:]   # 5 enabled algorithms x m = 4,6,...,16
    assert "[ERROR]" not in out and "[--]" not in out and "[OUT]" not in out
    occ = [int(v) for v in re.findall(r"occ (\d+)", out)]
    assert len(occ) == 35 and min(occ) >= 1
    # all algorithms of one length saw the same pattern set, so they report the same mean occurrence count
    for k in range(0, 35, 5):
        assert len(set(occ[k:k + 5])) == 1
    searches, uploads = stats(worker)
    assert searches == 35 * 4 and uploads == 1   # 140 searches, the 64 KiB text uploaded once
