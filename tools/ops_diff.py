"""Side-by-side per-op times of LDPROF dumps (last call of each file). usage: ops_diff.py a.txt b.txt ..."""
import sys
def load(fn):
    calls = open(fn).read().split("=== call")
    rows = [l.split() for l in calls[-1].splitlines() if l.startswith("LDPROF") and "total" not in l]
    return [float(r[3]) for r in rows]
cols = [load(f) for f in sys.argv[1:]]
n = max(len(c) for c in cols)
for i in range(n):
    print(f"{i:3d} " + " ".join(f"{c[i]:9.1f}" if i < len(c) else " " * 9 for c in cols))
print("sum " + " ".join(f"{sum(c):9.1f}" for c in cols))
