#!/bin/bash
# run on the GPU box: kernel GB/s of every tuning variant on a few representative cases
cd "$(dirname "$0")/.."
for lib in smart_b200/lib/variants/*.so; do
  name=$(basename $lib .so)
  SMART_B200_LIB=$PWD/$lib python tools/sweep.py --gib 4 --sigmas 256 --ms 4,16,64,1024 --default-only --reps 7 --out gpurun_out/tune_${name}_256.jsonl > /dev/null 2>&1
  SMART_B200_LIB=$PWD/$lib python tools/sweep.py --gib 4 --sigmas 4 --ms 8,16,64 --default-only --reps 7 --out gpurun_out/tune_${name}_4.jsonl > /dev/null 2>&1
  echo "$name: $(cat gpurun_out/tune_${name}_256.jsonl gpurun_out/tune_${name}_4.jsonl | python -c 'import sys,json; print(" ".join("s%d/m%d:f%d:%.0f(%.0f)"%(r["sigma"],r["m"],r["filter"],r["GBps"],r["GBps_best"]) for r in map(json.loads,sys.stdin)))')"
done
