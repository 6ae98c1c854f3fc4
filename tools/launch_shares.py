#!/usr/bin/env python
"""Per-kernel totals and shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split('(')[0][:100]
    tot[name][0] += 1
    tot[name][1] += float(r[-1])
s = sum(v[1] for v in tot.values())
print('kernel,launches,total_ns,share')
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f'"{k}",{v[0]},{v[1]:.0f},{v[1] / s:.4f}')
