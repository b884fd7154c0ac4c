#!/usr/bin/env python
"""Benchmark of the B200-native LocalDiffusion sampler (driver contract: see DESIGN.md §Measurement).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm (oracle port) on host cores

A "step" is one pass of the hot path over one batch: the full T=1000-timestep local-diffusion
sampling (IND/OOD branches, fusion at start_timestep=2, 1998 UNet image-forwards per image) of
B=16 synthetic 256x256 conditional images per GPU (BASELINE.json configs[1]; weak scaling:
global batch 16*N, which at N=8 is configs[2]).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "local_diffusion_sampling_256x256_T1000_images_per_sec"
UNIT = "img/s"
GF_PER_IMAGE_FORWARD = 75.35  # BASELINE.md §3, un-hoisted reference count (UNet 69.53 + cond encoder 5.81)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="images per GPU (development override)")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--timesteps", type=int, default=1000)
    ap.add_argument("--start-timestep", type=int, default=2)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


# ------------------------------------------------------------------------------------------------
# CPU leg: the reference's algorithm (oracle port, torch CPU fp32) on the host cores, bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_sample_rate(args, cores, sample_batch=8):
    """img/s of the reference's CPU algorithm for the bench workload, extrapolated from one branched
    step + one fused-phase step at `sample_batch` images (a full run is hours)."""
    import torch

    from oracle import ld_oracle as lo
    from tests import util
    from tests.golden import cases

    torch.set_num_threads(cores)
    S, T, s = args.size, args.timesteps, args.start_timestep
    sd = util.cpu_state_dict(util.make_model("mri"))
    hp = util.hp_of("mri")
    cond, mask = cases.mri_like(sample_batch, S)
    x = cases.noise_tape(sample_batch, S, 1)[0]
    cfg = cases.base_config("mri", s)
    smp = lo.Sampler(cfg, sd, hp, image_size=S, timesteps=T)
    z = lambda: x.clone()
    with torch.no_grad():
        t0 = time.perf_counter()
        smp._p_sample([x, x], mask, cases.MRI_MIN_MAX, cond, T - 1, z)      # branched step (2 UNet forwards)
        t_br = time.perf_counter() - t0
        cfg["branch_out"] = False
        t0 = time.perf_counter()
        smp._p_sample(x, mask, cases.MRI_MIN_MAX, cond, 1, z)               # fused-phase step (1 UNet forward)
        t_si = time.perf_counter() - t0
    total = (T - s) * t_br + s * t_si
    return sample_batch / total, f"extrapolated from 1 branched + 1 single step at B={sample_batch}, {S}x{S}: {t_br:.2f}s + {t_si:.2f}s"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vals = []
    sample = ""
    for i in range(args.warmup + args.steps):
        v, sample = cpu_sample_rate(args, cores)
        if i >= args.warmup:
            vals.append(v)
        if i == 0 and 1.0 / max(v, 1e-12) > 0:  # keep the whole run within minutes: one warm-up is enough on CPU
            pass
    value = statistics.median(vals)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * args.batch / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, world):
    return {
        "workload": f"BASELINE configs[1]: {args.size}x{args.size} single-channel conditional translation (synthetic T1-like + OOD blob), "
                    f"mri Unet(dim=32), T={args.timesteps} DDPM steps, start_timestep={args.start_timestep}, batch {args.batch}/GPU",
        "global_batch": args.batch * world, "image": args.size, "timesteps": args.timesteps,
        "unet_forwards_per_image": 2 * (args.timesteps - args.start_timestep) + args.start_timestep,
        "parallelism": f"batch-sharded x{world}, final all-gather only",
        "l2": "working set (activations + noise tape) >> 126 MB L2, no flush needed",
    }


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from localdiffusion_hallucination_b200 import GaussianDiffusion, _lib, parallel
    from tests import util
    from tests.golden import cases

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    S, T, s, B = args.size, args.timesteps, args.start_timestep, args.batch
    GB = B * world
    model = util.make_model("mri", args.precision, device=dev)
    gd = GaussianDiffusion(cases.base_config("mri", s), model, image_size=S, timesteps=T, objective="pred_x0").to(dev)
    cond_g, mask_g = cases.mri_like(GB, S)
    lo_, hi_ = parallel.shard_bounds(GB, rank, world)
    cond_h, mask_h = cond_g[lo_:hi_].contiguous().pin_memory(), mask_g[lo_:hi_].contiguous().pin_memory()
    cond_d, mask_d = cond_h.to(dev), mask_h.to(dev)
    mm = cases.MRI_MIN_MAX
    # device-resident noise tape in the reference's draw order (seed 10); the same tape is reused every step
    tape = gd.make_noise_tape((B, 1, S, S), T, dev)
    h = model.engine()
    lib = _lib.lib()

    def step_resident():
        out = gd.sample(cond_d, None, batch_size=B, mask=mask_d, min_max_val=mm, noise=tape)
        return parallel.gather_rows(out, GB)

    host_out = torch.empty(B, 1, S, S).pin_memory()

    def step_e2e():
        c = cond_h.to(dev, non_blocking=True)
        m = mask_h.to(dev, non_blocking=True)
        out = gd.sample(c, None, batch_size=B, mask=m, min_max_val=mm, noise=None)  # noise drawn on device, seed 10
        host_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host_out

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, k):
        sync_all()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(k):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        sync_all()
        return float(ms)

    for _ in range(args.warmup):
        step_resident()
    l0 = lib.ld_launch_count(h)
    with ClockSampler(local) as cs:
        ms = timed(step_resident, args.steps)
    launches = (lib.ld_launch_count(h) - l0) * world
    clocks = cs.summary()
    value = GB * args.steps / (ms / 1000.0)
    e2e = None
    if not args.no_e2e:
        step_e2e()
        ke = min(args.steps, 2)
        t0 = time.perf_counter()
        ms_e = timed(step_e2e, ke)
        e2e = {"value": GB * ke / (ms_e / 1000.0), "unit": UNIT, "h2d_bytes_per_step": int(cond_h.numel() * 4 * 2 * world),
               "d2h_bytes_per_step": int(host_out.numel() * 4 * world), "wall_ms_per_step": 1000.0 * (time.perf_counter() - t0) / ke,
               "note": "pinned host cond+mask -> device, noise drawn on device (seed 10, reference order), result -> pinned host"}

    if rank == 0:
        hbm, tf_burst, tf_sus, src = measured_peaks()
        fwd = 2 * (T - s) + s
        tflops = value * fwd * GF_PER_IMAGE_FORWARD / 1000.0
        roof = dominant_kernel_roofline(lib, dev, B, S, hbm, src)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic", "config": workload_config(args, world),
            "unet_tflops": tflops, "unet_tflops_frac_of_sustained_peak": tflops / (tf_sus * world),
            "flop_count": f"{GF_PER_IMAGE_FORWARD} GF per image-forward (reference, cond encoder not hoisted) x {fwd} forwards/image",
            "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e, "roofline": roof,
            "peaks": {"hbm_gbs": hbm, "bf16_tflops_burst": tf_burst, "bf16_tflops_sustained": tf_sus, "source": src},
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, sample = cpu_sample_rate(args, cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def dominant_kernel_roofline(lib, dev, B, S, hbm, src):
    """Roofline of the dominant kernel, timed alone with CUDA events on its launch stream: the tcgen05
    3x3 convolution at full resolution, 32->32 channels (16 of the 45 UNet convs, the HBM-bound ones).
    Algorithmic bytes per launch = N*H*W*(Cin+Cout)*2 (bf16 in + out, each touched once) + weights."""
    import ctypes as C

    import torch

    if not hasattr(lib, "ld_debug_conv_time"):
        return None
    N = 2 * B
    ms = C.c_float(0)
    rc = lib.ld_debug_conv_time(2, 32, 0, N, S, S, 0, 32, 3, 20, C.byref(ms), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    if rc != 0:
        return None
    byts = N * S * S * (32 + 32) * 2 + 9 * 32 * 32 * 2
    ach = byts / (ms.value / 1000.0) / 1e9
    traffic, tsrc = None, None
    tp = os.path.join(ROOT, "profiles", "r1v_conv32_traffic.json")
    if N == 32 and S == 256 and os.path.isfile(tp):  # the ncu capture was taken on exactly this launch shape
        t = json.load(open(tp))
        traffic, tsrc = t["traffic_bytes_per_launch"], t["source"]
    return {"kernel": "conv_tc_kernel<32,3,32> (3x3, 32->32 ch, %dx%dx%d; TMA in, tcgen05, TMA out)" % (N, S, S), "bound": "hbm",
            "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": traffic, "traffic_source": tsrc,
            "ms_per_launch": ms.value, "peak_source": src, "algorithmic_bytes_per_launch": byts}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
