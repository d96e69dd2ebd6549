#!/bin/bash
# BASELINE.json configs[0]: the SMART driver on the bundled corpus (englishTexts first), text-extracted
# patterns m = 2..32 (-short), BF/KMP/Horspool/Shift-Or (+BM, HASH3, QF23): once with OUR executables and the
# resident GPU worker, once with the REFERENCE executables on the host CPU, same seed => same pattern sets.
# Needs data_local/ = a git-ignored copy of the reference data: mkdir data_local && cp -r /root/reference/data/{englishTexts,italianTexts,frenchTexts,chineseTexts,protein,midimusic} data_local/
set -e
cd "$(dirname "$0")/../.."
W=${1:-gpurun_out/corpus_demo}
mkdir -p "$W"
TEXTS=englishTexts-italianTexts-frenchTexts-chineseTexts-protein-midimusic
ALGOS=bf,kmp,hor,so,bm,hash3,qf23
clean() { tr '\b' ' ' | sed 's/ \{4,\}\[[0-9]*%\]//g'; }
for mode in short large; do
  flag=$([ $mode = short ] && echo -short || echo "")
  ARGS="-text $TEXTS -data data_local -results $W/results -pset 50 -tsize 7 -occ -pre -seed 11 $flag -algo $ALGOS"
  bin/smart $ARGS | clean > "$W/ours_$mode.txt"
  bin/smart $ARGS -batch | clean > "$W/ours_batch_$mode.txt"
  if [ -x oracle/_ref/bin/bf ]; then bin/smart $ARGS -bin oracle/_ref/bin -noworker | clean > "$W/reference_$mode.txt"; fi
done
python - "$W" <<'PY'
import re, sys
w = sys.argv[1]
def cells(p):
    return re.findall(r"\] (\w+) \.+\[(\w+|--)\]([^\n]*)", open(p).read())
occ = lambda c: c[2].split("occ")[-1].strip()
for mode in ("short", "large"):
    a, b, r = (cells(f"{w}/{k}_{mode}.txt") for k in ("ours", "ours_batch", "reference"))
    both = [(x, y, z) for x, y, z in zip(a, b, r) if x[1] == "OK" and z[1] == "OK"]
    print(mode, "cells:", len(a), len(b), len(r), "finished by both:", len(both),
          "ours == reference:", all(occ(x) == occ(z) for x, y, z in both),
          "batch == reference:", all(occ(y) == occ(z) for x, y, z in both),
          "| reference not-OK cells:", sum(z[1] != "OK" for z in r), "ours not-OK:", sum(x[1] != "OK" for x in a))
PY
