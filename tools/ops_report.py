"""Per-op profile (LDPROF lines of tools/gpu_profile_ops.py) -> markdown with per-level sums.  usage: ops_report.py ops.txt [title]"""
import re, sys
txt = open(sys.argv[1]).read()
calls = txt.split("=== call")
last = calls[-1]
ops = [(int(m.group(1)), int(m.group(2)), float(m.group(3))) for m in re.finditer(r"LDPROF\s+(\d+) n=(\d+)\s+([\d.]+) us", last)]
tot = re.search(r"LDPROF total\s+([\d.]+) us over (\d+) ops \((.*)\)", last)
levels = [("time MLP + init conv", 0, 1), ("256x256 down (2 ResnetBlocks + LinearAttention + down conv)", 2, 11), ("128x128 down", 12, 21),
          ("64x64 down", 22, 31), ("32x32 (down3, mid, cond fusion, up0; 3 full attentions)", 32, 75), ("64x64 up", 76, 87),
          ("128x128 up", 88, 97), ("256x256 up + final block (final conv folded in)", 98, 10 ** 6)]
print(f"One UNet forward, {tot.group(3)}, bf16, eager launches with CUDA events around every op of the plan (last of {len(calls) - 1} calls).\n")
print("| level | ops | us |\n|---|---|---:|")
for name, a, b in levels:
    s = sum(u for i, n, u in ops if a <= i <= b)
    hi = min(b, ops[-1][0])
    print(f"| {name} | {a}-{hi} | {s:.0f} |")
print(f"| **total** | | **{float(tot.group(1)):.0f}** |\n")
print("```\nop  kernels   us")
for i, n, u in ops: print(f"{i:3d} {n:3d} {u:9.1f}")
print("```")
