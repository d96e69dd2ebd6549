#!/bin/bash
# A/B of two builds of the library on the SAME box (kernel-only GB/s, default kernels): usage tools/ab_hotloop.sh
for rep in 1 2; do
for v in headloop cur; do
  if [ $v = cur ]; then L=smart_b200/lib/libsmartb200.so; else L=smart_b200/lib/libsmartb200_$v.so; fi
  SMART_B200_LIB=$L python tools/sweep.py --gib 4 --sigmas 256 --default-only --ms 4,5,8,16,64 --out gpurun_out/ab_${v}_${rep}_256.jsonl > /dev/null 2>&1
  SMART_B200_LIB=$L python tools/sweep.py --gib 4 --sigmas 4 --default-only --flags 8192 --algo bf --ms 9,12,14,16 --out gpurun_out/ab_${v}_${rep}_4.jsonl > /dev/null 2>&1
done; done
