#!/usr/bin/env python
"""Summarise an .ncu-rep: per-kernel headline metrics, and instructions per source line for one launch.
usage: ncu_lines.py <rep> [launch-index [top-n]]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_pipe_lsu.sum']
for i, r in enumerate(rows[2:]):
    print(i, ' | '.join(f"{r[h.index(w)]}" for w in want if w in h))
if len(sys.argv) > 2:
    k = int(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-skip', str(k), '--launch-count', '1', '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    agg = {}; f = ''
    for r in csv.reader(io.StringIO(src)):
        if len(r) < 10:
            if r and r[0] == 'File Path': f = r[1].split('/')[-1]
            continue
        if r[0] in ('Line No', ''): continue
        try: agg[(f, int(r[0]))] = [r[1][:120], int(r[7]), int(r[6])]
        except ValueError: pass
    tot = sum(v[1] for v in agg.values())
    print('total warp-instructions', tot)
    for kk, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{kk[0][:14]:14s} {kk[1]:4d} {v[1]:9d} {100 * v[1] / tot:5.1f}% samp {v[2]:4d}  {v[0]}")
