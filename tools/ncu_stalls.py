"""Top warp-stall locations (SASS) of an ncu capture taken with --import-source on.
usage: ncu_stalls.py report.ncu-rep|source.csv[.gz] [top] [ctx_line ...]"""
import csv, gzip, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
if rep.endswith(".gz"):
    raw = gzip.open(rep, "rt").read()
elif rep.endswith(".csv"):
    raw = open(rep).read()
else:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); h = rows[1]; data = rows[2:]
si = h.index("Warp Stall Sampling (All Samples)"); src = h.index("Source"); ie = h.index("Instructions Executed")
tot = sum(int(r[si] or 0) for r in data if len(r) > si)
print(rows[0][1][:90], "total samples", tot)
for t in sorted([(int(r[si] or 0), i, r[ie], r[src].strip()) for i, r in enumerate(data) if len(r) > si], reverse=True)[:top]:
    print(f"{100*t[0]/tot:5.1f}% line {t[1]:4d} exec {t[2]:>8s} {t[3][:100]}")
g = collections.Counter()
for r in data:
    try: g[int(r[ie])] += int(r[si] or 0)
    except Exception: pass
print("samples by exec-count class:", sorted(g.items(), key=lambda kv: -kv[1])[:8])
for c in sys.argv[3:]:
    c = int(c); print("---- context", c)
    for i in range(max(0, c - 14), c + 3): print(i, data[i][si], data[i][ie], data[i][src].strip()[:110])
