#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total us, share."""
import csv, sys, re, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]); v = float(r[vi].replace(",", ""))
    v = v / 1000.0 if r[ui] in ("ns", "nsecond") else v
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"total {tot/1000:.1f} ms over {sum(a[0] for a in agg.values())} launches\n")
print("| kernel | launches | us | share |\n|---|---:|---:|---:|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k[:110]}` | {a[0]} | {a[1]:.1f} | {100*a[1]/tot:.1f}% |")
