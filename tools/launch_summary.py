"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
order = []
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r'\(.*', '', r[ki]).replace('wb::<unnamed>::', '').replace('void ', '')
    order.append((name, float(r[vi].replace(',', '')) / 1e3))
agg = collections.OrderedDict()
for n, t in order:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(v[1] for v in agg.values())
print("kernel,launches,total_us,share")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%s,%d,%.1f,%.4f" % (k, v[0], v[1], v[1] / tot))
print("TOTAL,%d,%.1f,1.0" % (len(order), tot))
if len(sys.argv) > 2:
    print("\n# launches in order (name, us)")
    for n, t in order:
        print("%s,%.1f" % (n, t))
