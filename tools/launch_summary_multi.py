"""Summarise an ncu --csv launch list with several metrics per launch: per kernel name the count, average duration, share of
the total, and average DRAM read / write bytes and achieved DRAM rate."""
import csv, sys
from collections import defaultdict
lines = open(sys.argv[1]).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(lines[start:]))
per = defaultdict(dict)          # launch id -> metric -> value
name = {}
for r in rows:
    n = r['Kernel Name'].replace('vqb200::', '')
    n = n[:n.index('(')] if '(' in n else n
    name[r['ID']] = n[:58]
    v = float(r['Metric Value'].replace(',', ''))
    u = r['Metric Unit']
    if u == 'Mbyte': v *= 1e6
    elif u == 'Kbyte': v *= 1e3
    elif u == 'Gbyte': v *= 1e9
    elif u == 'us': v *= 1e3
    elif u == 'ms': v *= 1e6
    per[r['ID']][r['Metric Name']] = v
agg = defaultdict(lambda: defaultdict(list))
for i, m in per.items():
    for k, v in m.items():
        agg[name[i]][k].append(v)
tot = sum(sum(a['gpu__time_duration.sum']) for a in agg.values())
print(f"{'kernel':60s} {'n':>4s} {'avg us':>9s} {'share':>6s} {'rd MB':>9s} {'wr MB':>9s} {'DRAM GB/s':>9s} {'L2 red sectors':>14s}")
for n, a in sorted(agg.items(), key=lambda kv: -sum(kv[1]['gpu__time_duration.sum'])):
    t = a['gpu__time_duration.sum']
    avg = sum(t) / len(t)
    rd = sum(a.get('dram__bytes_read.sum', [0])) / len(t)
    wr = sum(a.get('dram__bytes_write.sum', [0])) / len(t)
    red = sum(a.get('lts__t_sectors_op_red.sum', [0])) / len(t)
    print(f"{n:60s} {len(t):4d} {avg/1e3:9.1f} {100*sum(t)/tot:5.1f}% {rd/1e6:9.1f} {wr/1e6:9.1f} {(rd+wr)/avg:9.0f} {red:14.0f}")
