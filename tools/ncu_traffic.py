#!/usr/bin/env python
"""Turn an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,... --csv` log of
bench.py into per-family DRAM traffic of ONE picture (from one DMVR-16x16 launch to the next).  usage: ncu_traffic.py in.csv out.json"""
import csv, re, collections, json, sys
src, dst = sys.argv[1], sys.argv[2]
rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
h = rows[0]; col = {n: h.index(n) for n in ('Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit', 'ID', 'Grid Size', 'Block Size')}
L = collections.OrderedDict()
for r in rows[1:]:
    d = L.setdefault(r[col['ID']], {'name': re.sub(r'\(.*', '', r[col['Kernel Name']]), 'grid': r[col['Grid Size']], 'block': r[col['Block Size']]})
    d[r[col['Metric Name']]] = float(r[col['Metric Value']].replace(',', '')); d['u_' + r[col['Metric Name']]] = r[col['Metric Unit']]
ids = list(L)
starts = [k for k, i in enumerate(ids) if L[i]['name'].startswith('void mc_kernel<3, 4>') and '(64' in L[i]['block']]
a, b = starts[0], starts[1]
def famof(n):
    for key, f in (('alf', 'alf'), ('mc_', 'mc'), ('k1_', 'k1'), ('lf_', 'lf'), ('sao', 'sao')):
        if key in n: return f
    return 'bucket'
B = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}; T = {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1, 'msecond': 1e3}
tot = collections.OrderedDict(); lines = []
for k in range(a, b):
    d = L[ids[k]]; f = famof(d['name'])
    us = d['gpu__time_duration.sum'] * T[d['u_gpu__time_duration.sum']]
    rd = d['dram__bytes_read.sum'] * B[d['u_dram__bytes_read.sum']]; wr = d['dram__bytes_write.sum'] * B[d['u_dram__bytes_write.sum']]
    t = tot.setdefault(f, {'us_serial': 0, 'dram_read_bytes': 0, 'dram_write_bytes': 0, 'warp_inst_M': 0, 'launches': 0})
    t['us_serial'] += us; t['dram_read_bytes'] += rd; t['dram_write_bytes'] += wr; t['warp_inst_M'] += d['smsp__inst_executed.sum'] / 1e6; t['launches'] += 1
    lines.append({'kernel': d['name'], 'grid': d['grid'], 'block': d['block'], 'us': round(us, 1), 'dram_read_MB': round(rd / 1e6, 2), 'dram_write_MB': round(wr / 1e6, 2),
                  'warp_inst_M': round(d['smsp__inst_executed.sum'] / 1e6, 2), 'issue_active_pct': round(d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0), 1)})
json.dump({'source': src + ' (one 4K picture; launches serialised by ncu, caches not flushed between launches)',
           'per_family': {f: {k: round(v, 3) for k, v in t.items()} for f, t in tot.items()}, 'launches': lines}, open(dst, 'w'), indent=1)
for f, t in tot.items(): print(f, {k: round(v, 2) for k, v in t.items()})
