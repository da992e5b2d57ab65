"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, average and share."""
import csv, sys
from collections import defaultdict
lines = open(sys.argv[1]).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(lines[start:]))
d = defaultdict(list)
for r in rows:
    n = r['Kernel Name'].replace('vqb200::', '')
    n = n[:n.index('(')] if '(' in n else n
    d[n[:60]].append(float(r['Metric Value'].replace(',', '')))
tot = sum(sum(v) for v in d.values())
for n, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{n:62s} n={len(v):4d} avg={sum(v)/len(v)/1e3:10.1f} us  total={sum(v)/1e3:10.1f} us  {100*sum(v)/tot:5.1f} %")
