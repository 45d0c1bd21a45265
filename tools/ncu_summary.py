#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu -i ... --page raw --csv) into the handful of metrics DESIGN.md and
profiles/ quote.  Usage: python tools/ncu_summary.py file.ncu-rep [...] > summary.json"""
import csv
import io
import json
import subprocess
import sys

WANT = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'launch__registers_per_thread', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.per_cycle_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'smsp__thread_inst_executed_per_inst_executed.ratio']
STALL = 'smsp__average_warps_issue_stalled_'


def summarise(path):
    txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {}
        for i, h in enumerate(hdr):
            if h in WANT:
                d[h] = r[i] + (' ' + units[i] if units[i] else '')
            elif h.startswith(STALL) and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h:
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v >= 0.05:
                    d.setdefault('stalls_per_issue', {})[h[len(STALL):-len('_per_issue_active.ratio')]] = round(v, 3)
        out.append(d)
    return out


if __name__ == '__main__':
    print(json.dumps({p: summarise(p) for p in sys.argv[1:]}, indent=1))
