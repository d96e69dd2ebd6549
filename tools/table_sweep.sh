#!/bin/bash
# Kernel-only rates of every table-driven scan (SB_FLAG_TABLE_SCAN = 16384) on 1 GiB, sigma 256 and 4.
# Usage: tools/table_sweep.sh [out.jsonl]
out=${1:-gpurun_out/sweep_table_scans.jsonl}
rm -f "$out"
for a in kmp mp so sa kr aut ac hor bm qs raita tunedbm ssabs smith fs graspm tbm fjs skip br tvsbs bom ebom bndm sbndm2 \
         hash3 hash5 hash8 qf23 qf34 qf63; do
    python "$(dirname "$0")/sweep.py" --flags 16384 --algo $a --default-only --gib 1 --sigmas 256,4 --ms 8,32,64,1024 \
        --reps 3 --append --out "$out" > /dev/null || echo "sweep failed for $a"
done
wc -l "$out"
