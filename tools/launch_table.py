#!/usr/bin/env python
"""Per-launch table (last run of the command) from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log."""
import csv, collections, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 4
byid = collections.OrderedDict()
for x in csv.DictReader(lines):
    d = byid.setdefault(x['ID'], {'name': x['Kernel Name'].split('(')[0], 'grid': x['Grid Size'], 'block': x['Block Size']})
    v = float(x['Metric Value'].replace(',', ''))
    u = x['Metric Unit']
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1, 'us': 1e3, 'ms': 1e6}.get(u, 1)
    d[x['Metric Name']] = v * scale
ids = list(byid)
n = len(ids) // runs
print(f"# {len(ids)} launches captured, {n} per run; last run:")
tot = 0
for i in ids[-n:]:
    d = byid[i]
    t = d['gpu__time_duration.sum']; tot += t
    r = d.get('dram__bytes_read.sum', 0); w = d.get('dram__bytes_write.sum', 0)
    print(f"{d['name'][:34]:34s} grid {d['grid']:>16s} block {d['block']:>12s} {t/1e6:8.3f} ms  dram rd {r/1e9:7.3f} GB wr {w/1e9:7.3f} GB  {(r+w)/t:7.1f} GB/s")
print(f"# sum {tot/1e6:.3f} ms (serialised, cold-cache: compare shares)")
