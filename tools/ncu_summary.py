"""One-line-per-metric summary of `ncu --page raw --csv` exports (development aid; feeds profiles/*.md).
usage: ncu_summary.py raw.csv [raw2.csv ...]"""
import csv, sys
KEYS = [
    ("gpu__time_duration.sum", "us", 1e-3),
    ("launch__grid_size", "CTAs", 1), ("launch__block_size", "threads", 1), ("launch__registers_per_thread", "regs", 1),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem)", 1), ("sm__warps_active.avg.pct_of_peak_sustained_active", "% warps active", 1),
    ("dram__bytes_read.sum", "MB read", 1e-6), ("dram__bytes_write.sum", "MB written", 1e-6),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "% DRAM", 1),
    ("lts__t_sector_hit_rate.pct", "% L2 hit", 1),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "% tensor pipe", 1),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "% smem pipe (tensor operand reads)", 1),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "% smem pipe (ld/st.shared)", 1),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "% XU (MUFU) pipe", 1),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "% FMA pipe", 1),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "% ALU pipe", 1),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "% issue slots", 1),
    ("smsp__inst_executed.sum", "M warp instructions", 1e-6),
]
def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}
cols = [load(p) for p in sys.argv[1:]]
print("| metric | " + " | ".join(p.split("/")[-1].replace(".raw.csv", "") for p in sys.argv[1:]) + " |")
print("|---|" + "---:|" * len(cols))
print("| kernel | " + " | ".join(c["Kernel Name"][0][:60] for c in cols) + " |")
for k, unit, sc in KEYS:
    out = []
    for c in cols:
        if k not in c: out.append("-"); continue
        v, u = c[k]
        try:
            f = float(v.replace(",", ""))
            if k == "gpu__time_duration.sum": f = f * (1e-3 if u in ("ns", "nsecond") else 1.0 if u in ("us", "usecond") else 1e3 if u in ("ms", "msecond") else 1e-3)
            elif k.startswith("dram__bytes"): f = f * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
            else: f = f * sc
            out.append(f"{f:.1f}")
        except ValueError:
            out.append(v)
    print(f"| `{k}` ({unit}) | " + " | ".join(out) + " |")
