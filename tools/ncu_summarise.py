#!/usr/bin/env python
"""Turns `ncu -i gpurun_out/r02_full_pictures.ncu-rep --page raw --csv` (one B and one I picture, tools/ncu_round2.sh) into profiles/r02_ncu_full_pictures.md and
profiles/r02_traffic.json.  usage: ncu_summarise.py raw.csv <commit of the capture> [first launch id of the I picture]"""
import csv, json, collections, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
commit = sys.argv[2] if len(sys.argv) > 2 else "?"
hdr, unit, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def g(r, k):
    try: return float(r[ix[k]].replace(',', ''))
    except Exception: return 0.0
def scaled(r, k):
    m = {'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6, 'byte': 1e-6, 'kbyte': 1e-3, 'mbyte': 1, 'gbyte': 1e3}
    return g(r, k) * m.get(unit[ix[k]].lower(), 1)
def fam_of(n): return 'mc' if 'mc_' in n else 'k1' if 'k1_' in n else 'intra' if 'intra' in n else 'alf' if 'alf_' in n else 'lf' if 'lf_' in n else 'sao'
stall = [h for h in hdr if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued')]
dthr = [h for h in hdr if 'dram__throughput.avg.pct_of_peak_sustained_elapsed' in h][0]
lthr = [h for h in hdr if 'lts__throughput.avg.pct_of_peak_sustained_elapsed' in h][0]
# the I picture starts with the first K1 launch after the first ALF launch
names = [r[ix['Kernel Name']] for r in data]
first_alf = next(i for i, n in enumerate(names) if 'alf_' in n)
split = int(sys.argv[3]) if len(sys.argv) > 3 else next(i for i in range(first_alf, len(names)) if 'k1_' in names[i] or 'mc_' in names[i])
lines = [f"# ncu --set full --clock-control none, one B picture (launches 0-{split - 1}) and one I picture ({split}-{len(data) - 1}) of the bench workload, 3840x2160 10 bit",
         "# (tools/ncu_round2.sh -> tools/prof_pictures.py, summarised by tools/ncu_summarise.py; the .ncu-rep is not committed).",
         "# Per-launch times are serialised and cold-cache (ncu flushes caches between replays): shares, not absolutes; the bench's per-family ms is the live figure.",
         f"# Commit of the capture: {commit}.",
         "id | kernel | grid x block | us | DRAM rd MB | DRAM wr MB | regs | smem KB/block | warps active % | issue active % | SM thr % | L1 hit % | L2 hit % | DRAM thr % | L2 thr % | top stall reasons (pc samples)"]
pic = {'B': collections.OrderedDict(), 'I': collections.OrderedDict()}
for r in data:
    name, i = r[ix['Kernel Name']], int(r[ix['ID']])
    st = sorted(((g(r, h), h.replace('smsp__pcsamp_warps_issue_stalled_', '')) for h in stall), reverse=True)
    tot = sum(v for v, _ in st) or 1
    top = ", ".join("%s %.0f%%" % (n, 100 * v / tot) for v, n in st[:4])
    us, rd, wr = scaled(r, 'gpu__time_duration.sum'), scaled(r, 'dram__bytes_read.sum'), scaled(r, 'dram__bytes_write.sum')
    sm = g(r, 'launch__shared_mem_per_block_dynamic') + g(r, 'launch__shared_mem_per_block_static')
    f = pic['B' if i < split else 'I'].setdefault(fam_of(name), [0, 0, 0]); f[0] += us; f[1] += rd; f[2] += wr
    lines.append("%d | %s | %s x %s | %.1f | %.2f | %.2f | %d | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %s" % (
        i, name[:46], r[ix['Grid Size']], r[ix['Block Size']], us, rd, wr, g(r, 'launch__registers_per_thread'), sm, g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'),
        g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'), g(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'), g(r, 'l1tex__t_sector_hit_rate.pct'),
        g(r, 'lts__t_sector_hit_rate.pct'), g(r, dthr), g(r, lthr), top))
lines.append("")
for k in ('B', 'I'):
    lines.append("## %s picture, per family: serialised us | DRAM read MB | DRAM write MB" % k)
    for f, (us, rd, wr) in pic[k].items(): lines.append("%s | %.1f | %.2f | %.2f" % (f, us, rd, wr))
open(os.path.join(ROOT, 'profiles', 'r02_ncu_full_pictures.md'), 'w').write("\n".join(lines) + "\n")
tr = {'width': 3840, 'height': 2160, 'source': f'ncu --set full, tools/ncu_round2.sh, commit {commit}; one B and one I picture of the bench workload',
      'per_family': {f: {'dram_read_bytes': int(v[1] * 1e6), 'dram_write_bytes': int(v[2] * 1e6), 'serialised_us': round(v[0], 1)} for f, v in pic['B'].items()},
      'per_family_I_picture': {f: {'dram_read_bytes': int(v[1] * 1e6), 'dram_write_bytes': int(v[2] * 1e6), 'serialised_us': round(v[0], 1)} for f, v in pic['I'].items()}}
json.dump(tr, open(os.path.join(ROOT, 'profiles', 'r02_traffic.json'), 'w'), indent=1)
print("\n".join(lines[-14:]))
