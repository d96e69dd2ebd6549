#!/bin/bash
# End-to-end run of the C driver on the GPU box: regenerate SMART's rand4 / rand250 buffers with textgen,
# run the same experiment (same seed => same pattern sets) with OUR per-algorithm executables (resident GPU
# worker) and with the REFERENCE's executables (oracle/_ref/bin, CPU), and keep both transcripts.
set -e
cd "$(dirname "$0")/../.."
W=${1:-gpurun_out/driver_demo}
mkdir -p "$W"
bin/textgen -data "$W/data" -only 4 > /dev/null
bin/textgen -data "$W/data" -only 250 > /dev/null
ARGS="-text rand4-rand250 -data $W/data -results $W/results -pset 20 -tsize 5 -occ -pre -seed 7 -txt"
bin/smart $ARGS -algo bf,kmp,hor,so,bm,hash3,qf23 | tr '\b' ' ' | sed 's/ \{4,\}\[[0-9]*%\]//g' > "$W/ours.txt"
if [ -x oracle/_ref/bin/bf ]; then
  bin/smart $ARGS -algo bf,kmp,hor,so,bm,hash3,qf23 -bin oracle/_ref/bin -noworker | tr '\b' ' ' | sed 's/ \{4,\}\[[0-9]*%\]//g' > "$W/reference.txt"
fi
grep -c "\[OK\]" "$W/ours.txt" "$W/reference.txt" || true
python - "$W" <<'PY'
import re, sys
w = sys.argv[1]
def cells(p):
    return re.findall(r"\] (\w+) \.+\[(\w+|--)\]([^\n]*)", open(p).read())
a, b = cells(f"{w}/ours.txt"), cells(f"{w}/reference.txt")
occ = lambda c: c[2].split("occ")[-1].strip()
both = [(x, y) for x, y in zip(a, b) if x[1] == "OK" and y[1] == "OK"]
print("cells:", len(a), len(b), "finished by both:", len(both), "identical occurrence columns:",
      all(occ(x) == occ(y) for x, y in both), "| reference [OUT]/[--] cells:", sum(y[1] != "OK" for y in b),
      "ours:", sum(x[1] != "OK" for x in a))
PY
