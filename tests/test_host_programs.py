"""The C host side (plugin shim, driver, harness, textgen) checked on the CPU by CROSS-DRIVING: our
driver and our harness speak the reference's shm + argv protocol, so they can run the reference's own
algorithm executables (oracle/_ref/bin, built from the unmodified sources).  GPU-side cross-driving (the
reference's `test` binary driving OUR executables) lives in test_gpu_host.py."""
import hashlib
import os
import re
import subprocess

import pytest

import oracle_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin")
REFBIN = os.path.join(ROOT, "oracle", "_ref", "bin")

needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REFBIN, "bf")), reason="oracle/_ref/bin not built")


@pytest.fixture(scope="module", autouse=True)
def built():
    subprocess.run(["make", "-C", os.path.join(ROOT, "smart_b200", "host")], check=True, capture_output=True)


def run(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, timeout=300, **kw)


def test_all_plugin_executables_exist():
    import smart_b200 as sb

    for a in sb.ALGORITHMS:
        assert os.access(os.path.join(BIN, "algos", a), os.X_OK), a
    for p in ("smart", "test", "textgen", "smartd", "abi_demo"):
        assert os.access(os.path.join(BIN, p), os.X_OK), p


def test_shim_argument_errors_match_reference_messages():
    # main.h:48-51 and :124-127
    r = run([os.path.join(BIN, "algos", "bf"), "shared", "1", "2"])
    assert r.returncode == 1 and "five parameters needed when used with shared memory" in r.stdout
    r = run([os.path.join(BIN, "algos", "bf"), "abc"])
    assert r.returncode == 1 and "four parameters needed in standard mode" in r.stdout
    # missing segments: shmget fails, exit 1 (main.h:60-63)
    r = run([os.path.join(BIN, "algos", "bf"), "shared", "991", "3", "992", "10", "993", "994", "995"])
    assert r.returncode == 1 and "shmget" in r.stderr


def test_plugin_fails_loudly_without_gpu():
    import smart_b200 as sb

    if sb.device_count() > 0:
        pytest.skip("a GPU is present")
    r = run([os.path.join(BIN, "algos", "kmp"), "aba", "3", "ababababab", "10"], env={**os.environ, "SMART_B200_NO_WORKER": "1"})
    assert r.returncode != 0 and "found" not in r.stdout  # never a made-up count
    # not-applicable is decided before any CUDA call, exactly like hash3.c:31
    r = run([os.path.join(BIN, "algos", "hash3"), "ab", "2", "ababababab", "10"])
    assert r.returncode == 0 and "found -1 occurrences" in r.stdout
    for args in (["-fg"], []):
        r = run([os.path.join(BIN, "smartd"), *args], env={**os.environ, "SMART_B200_SOCK": "/tmp/sb_test_nogpu.sock"})
        assert r.returncode != 0 and "no CUDA device" in r.stderr


@needs_ref
def test_our_harness_drives_reference_executables():
    for a in ("bf", "kmp", "hor", "so", "bm", "hash3", "qf23", "raita"):
        r = run([os.path.join(BIN, "test"), a, "-bin", REFBIN])
        assert r.returncode == 0 and "Test passed" in r.stdout, (a, r.stdout, r.stderr)
    # the harness must catch a wrong algorithm: sbndm2 is wrong on a^40 in a^64
    if not os.path.exists(os.path.join(REFBIN, "sbndm2")):
        subprocess.run(["gcc", "-O3", "-msse4", "-w", "-o", os.path.join(REFBIN, "sbndm2"),
                        "/root/reference/src/algos/sbndm2.c", "-lm"], check=False)
    if os.path.exists(os.path.join(REFBIN, "sbndm2")):
        r = run([os.path.join(BIN, "test"), "sbndm2", "-bin", REFBIN])
        assert r.returncode == 1 and "test failed on case n.16" in r.stdout
    # -strict: raita answers -1 for m=1 (raita.c:38) so it is reported, not hidden
    r = run([os.path.join(BIN, "test"), "raita", "-bin", REFBIN, "-strict"])
    assert r.returncode == 1 and "returned -1" in r.stdout


@needs_ref
def test_our_driver_drives_reference_executables(tmp_path):
    data = tmp_path / "data" / "mini"
    data.mkdir(parents=True)
    text = open(os.path.join(ROOT, "tests", "golden", "bible64k.bin"), "rb").read()
    (data / "part1.txt").write_bytes(text[:30000])
    (data / "part2.txt").write_bytes(text[30000:])
    (data / "index.txt").write_text("#part1.txt#\n#part2.txt#\n")
    r = run([os.path.join(BIN, "smart"), "-text", "mini", "-data", str(tmp_path / "data"), "-results",
             str(tmp_path / "results"), "-pset", "3", "-plen", "2", "16", "-short", "-algo", "bf,kmp,hash3,raita",
             "-bin", REFBIN, "-occ", "-pre", "-dif", "-std", "-txt", "-xml", "-tex", "-html", "-seed", "5", "-noworker"],
            cwd=tmp_path)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    assert "Text buffer of dimension 65536 byte" in out
    assert out.count("Searching for a set of 3 patterns with length") == 8  # m = 2,4,...,16
    assert out.count(".[OK]") == 8 * 4 - 1  # hash3 is not applicable to m=2 -> [--] (hash3.c:31)
    assert out.count(".[--]") == 1 and ".[ERROR]" not in out
    # same seed -> same pattern sets -> same occurrence columns
    r2 = run([os.path.join(BIN, "smart"), "-text", "mini", "-data", str(tmp_path / "data"), "-results",
              str(tmp_path / "results"), "-pset", "3", "-plen", "2", "16", "-short", "-algo", "bf,kmp", "-bin", REFBIN,
              "-occ", "-seed", "5", "-noworker"], cwd=tmp_path)
    occ = lambda s: re.findall(r"\[\d/\d\] (\w+) .*occ (\d+)", s)
    a, b = occ(out), occ(r2.stdout)
    assert [v for k, v in a if k in ("BF", "KMP")] == [v for k, v in b]
    assert all(v == w for (k, v), (k2, w) in zip(a[0::4], a[1::4]))  # BF and KMP agree
    txt = list((tmp_path / "results").glob("EXP*/mini.txt"))
    assert len(txt) == 1
    rows = txt[0].read_text().splitlines()
    assert [row.split()[0] for row in rows] == ["BF", "KMP", "HASH3", "RAITA"]
    assert rows[2].split()[1] == "-"  # hash3 @ m=2: no result
    xml = list((tmp_path / "results").glob("EXP*/mini.xml"))[0].read_text()
    assert xml.startswith("<RESULTS>") and xml.count("<ALGO>") == 4 and xml.count("<NAME>HASH3</NAME>") == 1
    assert xml.count("<DATA>-</DATA>") == 1 and xml.count("<SEARCH>") == 8 * 4 and "<BEST>" in xml  # (output.h:225-228)
    # -tex (output.h:153-194) and -html (+ index.html, output.h:635-745)
    tex = list((tmp_path / "results").glob("EXP*/mini.tex"))[0].read_text()
    assert tex.startswith("\\begin{tabular}{|l|llllllll|}") and tex.endswith("\\end{tabular}")
    assert "$m$ & $2$ & $4$" in tex and tex.count("\\textsc{") == 4 and "\\textsc{HASH3} & - &" in tex
    html = list((tmp_path / "results").glob("EXP*/mini.html"))[0].read_text()
    assert html.count("<tr>") == 5 and "<td><b>HASH3</b></td><td>--</td>" in html and html.count("<b><u>") == 8
    index = list((tmp_path / "results").glob("EXP*/index.html"))[0].read_text()
    assert index.count('<a href="mini.html">') == 1 and index.rstrip().endswith("</html>")
    # -simple mode (smart.c:496-507)
    r = run([os.path.join(BIN, "smart"), "-simple", "aba", "ababababab", "-algo", "bf", "-bin", REFBIN, "-occ",
             "-noworker"], cwd=tmp_path)
    assert "occ 4" in r.stdout


@needs_ref
def test_report_files_equal_the_reference_drivers(tmp_path):
    """-txt / -xml / -tex written by OUR driver against the files the UNMODIFIED reference driver (oracle/_ref/smart)
    writes for the same text, lengths and algorithms (both drive the reference's executables here).  Timings differ
    and the reference seeds its patterns with time(NULL) (smart.c:435), so numbers are masked; everything else - rows,
    columns, names, the "-" cells, the XML schema including outputXML's doubled DATA node (output.h:225-228) - must be
    byte-identical."""
    refsmart = os.path.join(ROOT, "oracle", "_ref", "smart")
    if not os.path.exists(refsmart):
        pytest.skip("oracle/_ref/smart not built")
    text = open(os.path.join(ROOT, "tests", "golden", "bible64k.bin"), "rb").read()
    # the reference only accepts the buffer names of its settings table (smart.c:80-87): use one of them
    data = tmp_path / "data" / "englishTexts"
    data.mkdir(parents=True)
    (data / "part1.txt").write_bytes(text[:30000])
    (data / "part2.txt").write_bytes(text[30000:])
    (data / "index.txt").write_text("#part1.txt#\n#part2.txt#\n")
    # the layout the reference sources hard-code: ./source/bin/<algo>, ./source/algorithms.h, ./results (smart.c:142)
    (tmp_path / "source").mkdir()
    os.symlink(REFBIN, tmp_path / "source" / "bin")
    (tmp_path / "source" / "algorithms.h").write_text("".join(f"#1 #{a} \n" for a in ("bf", "kmp", "hash3", "raita")))
    (tmp_path / "results").mkdir()
    r = run([refsmart, "-text", "englishTexts", "-pset", "3", "-plen", "2", "16", "-short", "-txt", "-tex"], cwd=tmp_path)
    assert r.returncode == 0 and "OUTPUT RUNNING TIMES" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    ours = tmp_path / "ours"
    ours.mkdir()
    r = run([os.path.join(BIN, "smart"), "-text", "englishTexts", "-data", str(tmp_path / "data"), "-results",
             str(ours / "results"), "-pset", "3", "-plen", "2", "16", "-short", "-algo", "bf,kmp,hash3,raita", "-bin", REFBIN,
             "-txt", "-xml", "-tex", "-seed", "5", "-noworker"], cwd=ours)
    assert r.returncode == 0, r.stderr

    def masked(path):
        s = re.sub(r"EXP\d+", "EXP", open(path).read())
        return re.sub(r"\d+\.\d+", "N", s)

    for ext in ("txt", "xml", "tex"):
        theirs = list((tmp_path / "results").glob(f"EXP*/englishTexts.{ext}"))
        mine = list((ours / "results").glob(f"EXP*/englishTexts.{ext}"))
        assert len(theirs) == 1 and len(mine) == 1, ext
        assert masked(mine[0]) == masked(theirs[0]), ext


def test_select_edits_a_registry_like_the_reference_tool(tmp_path):
    """bin/select: the command line of src/select.c on the '#<0|1> #<name>' registry format."""
    sel = os.path.join(BIN, "select")
    reg = str(tmp_path / "algorithms.h")
    common = ["-file", reg, "-bin", os.path.join(BIN, "algos")]
    r = run([sel, "-init", *common])
    assert r.returncode == 0 and "143 algorithms registered" in r.stdout
    lines = open(reg).read().splitlines()
    assert len(lines) == 143 and lines == sorted(lines) and all(ln.startswith("#0 #") for ln in lines)
    assert run([sel, "-show", *common]).stdout.count("\n") == 144
    r = run([sel, "bf", "kmp", *common])
    assert "The bf algorithm has been selected" in r.stdout and "The kmp algorithm has been selected" in r.stdout
    assert run([sel, "-which", *common]).stdout.split() == ["The", "list", "of", "selected", "algorithms:", "-bf", "-kmp"]
    assert "has been deselected" in run([sel, "kmp", *common]).stdout
    assert "#1 #bf " in open(reg).read() and "#0 #kmp " in open(reg).read()
    run([sel, "-all", *common])
    assert open(reg).read().count("#1 #") == 143
    run([sel, "-none", *common])
    assert open(reg).read().count("#1 #") == 0
    assert "no parameter nosuch" in run([sel, "nosuch", *common]).stdout
    # -add: the executable must exist and pass the harness (select.c:105-126)
    assert "already in the set" in run([sel, "-add", "bf", *common]).stdout
    assert "does not exist" in run([sel, "-add", "ghost", *common]).stdout
    extra = tmp_path / "bin"
    extra.mkdir()
    os.symlink(os.path.join(BIN, "algos", "bf"), extra / "mine")
    ok = run([sel, "-add", "mine", "-file", reg, "-bin", str(extra), "-test", "/bin/true"])
    assert "added succesfully" in ok.stdout and "#0 #mine " in open(reg).read()
    os.symlink(os.path.join(BIN, "algos", "bf"), extra / "broken")
    bad = run([sel, "-add", "broken", "-file", reg, "-bin", str(extra), "-test", "/bin/false"])
    assert "unable to add the algorithm broken" in bad.stdout and "broken" not in open(reg).read()
    # our driver reads the file it wrote (function.h:62-77 format)
    run([sel, "hor", *common])
    r = run([os.path.join(BIN, "smart"), "-simple", "aba", "ababababab", "-alist", reg, "-bin", REFBIN, "-occ", "-noworker"],
            cwd=tmp_path)
    assert "HOR" in r.stdout and "occ 4" in r.stdout


def test_driver_argument_errors():
    r = run([os.path.join(BIN, "smart")])
    assert "No parameter given" in r.stdout
    r = run([os.path.join(BIN, "smart"), "-pset", "3"])
    assert "No filename given" in r.stdout
    r = run([os.path.join(BIN, "smart"), "-text", "x", "-simple", "a", "b"])
    assert "Both parameters -simple and -text defined" in r.stdout
    r = run([os.path.join(BIN, "smart"), "-bogus"])
    assert "Error in input parameters" in r.stdout


def test_textgen_reproduces_reference_files(tmp_path):
    """md5 of SMART's rand2/rand4; -only keeps the rand() stream position."""
    r = run([os.path.join(BIN, "textgen"), "-data", str(tmp_path), "-only", "4"])
    assert r.returncode == 0
    f = tmp_path / "rand4" / "rand4.txt"
    assert hashlib.md5(f.read_bytes()).hexdigest() == "c371ad58eda30fa18f35a77b1e775ece"
    assert (tmp_path / "rand4" / "index.txt").read_text().strip() == "#rand4.txt#"
    assert not (tmp_path / "rand2").exists()
