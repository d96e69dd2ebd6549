#!/usr/bin/env python
"""Summarise ncu outputs into small committed files under profiles/.
  python tools/ncu_summary.py launches <launches.csv> <out.json>
  python tools/ncu_summary.py full <report.ncu-rep> <out.json>     (runs `ncu -i ... --page raw --csv`)
"""
import collections
import csv
import io
import json
import subprocess
import sys

KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'launch__local_load_bytes', 'smsp__inst_executed_op_local_st.sum', 'local_store_bytes']


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    h = rows[hdr]
    ki, vi = h.index('Kernel Name'), h.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) > vi:
            name = r[ki].split('(')[0].replace('void ', '').replace('sb::', '')
            agg.setdefault(name, []).append(float(r[vi].replace(',', '')))
    tot = sum(sum(v) for v in agg.values())
    res = [{"kernel": k, "launches": len(v), "avg_us": sum(v) / len(v) / 1e3, "share_of_gpu_time": sum(v) / tot}
           for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))]
    json.dump({"source": path, "note": "ncu --metrics gpu__time_duration.sum --clock-control none: cold-cache, "
               "serialised launches; compare shares, not absolutes", "kernels": res}, open(out, "w"), indent=1)
    for r in res:
        print(r)


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        ent = {}
        for k in KEEP:
            if k in h:
                ent[k] = (r[h.index(k)] + " " + units[h.index(k)]).strip()
        res.append(ent)
    json.dump({"source": rep, "launches": res}, open(out, "w"), indent=1)
    for e in res:
        print(json.dumps(e, indent=0)[:1500])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
