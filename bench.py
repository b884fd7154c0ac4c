#!/usr/bin/env python
"""Benchmark of the B200-native LocalDiffusion sampler (driver contract: see DESIGN.md §6).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm (oracle port) on host cores
    python bench.py --config c4|c5 [--start-timestep s]       # the other BASELINE configurations (development / SCALE runs)

A "step" is one pass of the hot path over one batch: the full T-timestep local-diffusion sampling (IND/OOD branches, fusion at
`start_timestep`, `2(T-s)+s` UNet image-forwards per image) of B synthetic conditional images per GPU.  Default = BASELINE.json
configs[1]: 256x256, T=1000, s=2, B=16 per GPU (weak scaling: global batch 16*N, which at N=8 is configs[2]).

The product arm imports only the package (`localdiffusion_hallucination_b200`); `oracle/` is imported by the `cpu_baseline`
leg and by `--impl reference` alone, after the GPU measurements are complete.
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

UNIT = "img/s"
# BASELINE.json configs: (model, image size, T, start_timestep, images per GPU)
CONFIGS = {
    "c2": dict(model="mri", size=256, timesteps=1000, start_timestep=2, batch=16,
               label="BASELINE configs[1]: 256x256 single-channel conditional translation (synthetic T1-like + OOD blob), mri Unet(dim=32)"),
    "c4": dict(model="mri_attn8", size=256, timesteps=1000, start_timestep=2, batch=8,
               label="BASELINE configs[3]: attention-heavy Unet(full_attn=(F,F,T,T), attn_heads=8), 256x256 (64x64 and 32x32 token grids)"),
    "c4s": dict(model="mri_attn8", size=128, timesteps=1000, start_timestep=2, batch=16,
                label="BASELINE configs[3]: attention-heavy Unet(full_attn=(F,F,T,T), attn_heads=8), 128x128 (32x32 and 16x16 token grids)"),
    "c5": dict(model="mri", size=512, timesteps=1000, start_timestep=500, batch=4,
               label="BASELINE configs[4]: branching-timestep sweep point, 512x512, mri Unet(dim=32), 4 images per GPU (batch 32 on 8 GPUs)"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=None, help="images per GPU (development override)")
    ap.add_argument("--size", type=int, default=None)
    ap.add_argument("--timesteps", type=int, default=None)
    ap.add_argument("--start-timestep", type=int, default=None)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-roofline", action="store_true", help="skip the stand-alone probes of the dominant kernel (launch-list captures)")
    a = ap.parse_args()
    c = CONFIGS[a.config]
    a.model = c["model"]
    for k in ("batch", "size", "timesteps", "start_timestep"):
        if getattr(a, k) is None:
            setattr(a, k, c[k])
    a.label = c["label"]
    return a


def metric_name(a):
    return f"local_diffusion_sampling_{a.size}x{a.size}_T{a.timesteps}_images_per_sec"


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


def workload_config(a, world):
    fwd = 2 * (a.timesteps - a.start_timestep) + a.start_timestep
    return {
        "workload": f"{a.label}, T={a.timesteps} DDPM steps, start_timestep={a.start_timestep}, batch {a.batch}/GPU",
        "config": a.config, "global_batch": a.batch * world, "image": a.size, "timesteps": a.timesteps,
        "start_timestep": a.start_timestep, "unet_forwards_per_image": fwd,
        "parallelism": f"batch-sharded x{world}, final all-gather only",
        "l2": "working set (activations + noise tape) >> 126 MB L2, no flush needed",
    }


# ------------------------------------------------------------------------------------------------
# CPU leg: the reference's algorithm (oracle port, torch CPU fp32) on the host cores, bounded sample
# ------------------------------------------------------------------------------------------------
class CpuLeg:
    """The reference's CPU algorithm (oracle port, torch fp32 on all host cores) on a bounded sample of the bench workload at the
    TRUE batch: `n_branched` branched steps (2 UNet forwards each) + 1 fused-phase step (1 forward), extrapolated with
    (T-s)*t_branched + s*t_single (SURVEY.md §8d prescribes 3 + 1; a full run is hours)."""

    def __init__(self, a, cores):
        import torch

        from localdiffusion_hallucination_b200 import workload as wl
        from oracle import ld_oracle as lo

        torch.set_num_threads(cores)
        self.a, self.wl, self.lo, self.torch = a, wl, lo, torch
        m = wl.make_model(a.model)
        self.sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
        kw = wl.MODEL_KW[a.model]
        self.hp = lo.UnetHP(dim=kw["dim"], init_dim=kw["init_dim"], dim_mults=kw.get("dim_mults", (1, 2, 4, 8)),
                            full_attn=kw.get("full_attn", (False, False, False, True)), heads=kw.get("attn_heads", 4), mode=kw["mode"])
        self.cond, self.mask = wl.mri_like(a.batch, a.size)
        self.x = wl.noise_tape(a.batch, a.size, 1)[0]
        self.t_single = None

    def _one(self, kind):
        a, wl, torch = self.a, self.wl, self.torch
        cfg = wl.base_config("mri", a.start_timestep)
        smp = self.lo.Sampler(cfg, self.sd, self.hp, image_size=a.size, timesteps=a.timesteps)
        z = lambda: self.x.clone()
        with torch.no_grad():
            t0 = time.perf_counter()
            if kind == "branched":
                smp._p_sample([self.x, self.x], self.mask, wl.MRI_MIN_MAX, self.cond, a.timesteps - 1, z)   # 2 UNet forwards
            else:
                cfg["branch_out"] = False
                smp._p_sample(self.x, self.mask, wl.MRI_MIN_MAX, self.cond, 1, z)                           # 1 UNet forward
        return time.perf_counter() - t0

    def rate(self, n_branched=3):
        a = self.a
        t_br = statistics.median([self._one("branched") for _ in range(n_branched)])
        fresh = self.t_single is None
        if fresh:
            self.t_single = self._one("single")
        total = (a.timesteps - a.start_timestep) * t_br + a.start_timestep * self.t_single
        return a.batch / total, (f"extrapolated from {n_branched} branched + {1 if fresh else 0} single step(s) at the bench batch B={a.batch}, "
                                 f"{a.size}x{a.size}: {t_br:.2f}s / branched step, {self.t_single:.2f}s / single step, (T-s)*t_br + s*t_si")


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    leg = CpuLeg(a, cores)
    leg.rate(1)   # one warm-up pass is enough on the CPU (keeps the whole run within minutes); it also times the single step
    leg.t_single = None   # re-timed inside the first (3 + 1) step
    vals, sample = [], ""
    for i in range(a.steps):
        # the first timed step is the prescribed 3 + 1 sample; later steps re-time one branched step (the single step is 0.1 % of a run)
        v, sm = leg.rate(3 if i == 0 else 1)
        sample = sample or sm
        vals.append(v)
    value = statistics.median(vals)
    line = {
        "metric": metric_name(a), "value": value, "unit": UNIT, "impl": "reference", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1000.0 * a.batch / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "config": workload_config(a, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample + f"; median of {a.steps} such steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist

    from localdiffusion_hallucination_b200 import GaussianDiffusion, _lib, parallel
    from localdiffusion_hallucination_b200 import workload as wl

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    S, T, s, B = a.size, a.timesteps, a.start_timestep, a.batch
    GB = B * world
    model = wl.make_model(a.model, a.precision, device=dev)
    gd = GaussianDiffusion(wl.base_config("mri", s), model, image_size=S, timesteps=T, objective="pred_x0").to(dev)
    cond_g, mask_g = wl.mri_like(GB, S)
    lo_, hi_ = parallel.shard_bounds(GB, rank, world)
    cond_h, mask_h = cond_g[lo_:hi_].contiguous().pin_memory(), mask_g[lo_:hi_].contiguous().pin_memory()
    cond_d, mask_d = cond_h.to(dev), mask_h.to(dev)
    mm = wl.MRI_MIN_MAX
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

    for _ in range(a.warmup):
        step_resident()
    l0 = lib.ld_launch_count(h)
    with ClockSampler(local) as cs:
        ms = timed(step_resident, a.steps)
    launches = (lib.ld_launch_count(h) - l0) * world
    clocks = cs.summary()
    value = GB * a.steps / (ms / 1000.0)
    e2e = None
    if not a.no_e2e:
        step_e2e()
        t0 = time.perf_counter()
        ms_e = timed(step_e2e, a.steps)   # every one of the --steps steps, host buffers in and out
        e2e = {"value": GB * a.steps / (ms_e / 1000.0), "unit": UNIT, "steps": a.steps, "h2d_bytes_per_step": int(cond_h.numel() * 4 * 2 * world),
               "d2h_bytes_per_step": int(host_out.numel() * 4 * world), "wall_ms_per_step": 1000.0 * (time.perf_counter() - t0) / a.steps,
               "note": "pinned host cond+mask -> device, noise drawn on device (seed 10, reference order), result -> pinned host"}

    if rank == 0:
        hbm, tf_burst, tf_sus, src = measured_peaks()
        fwd = 2 * (T - s) + s
        gf_fwd = wl.forward_gflop(a.model, S)   # reference count, cond encoder included (75.35 GF for the mri model at 256x256)
        tflops = value * fwd * gf_fwd / 1000.0
        # mixed roofline of SURVEY.md §8d: sum over layers of max(flops / tensor peak, min bytes / HBM bandwidth), measured peaks
        ideal_us = wl.mixed_roofline_us(a.model, S, tf_sus, hbm)
        us_per_fwd = 1e6 / (value / world * fwd)
        line = {
            "metric": metric_name(a), "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": a.precision, "data": "synthetic", "config": workload_config(a, world),
            "unet_tflops": tflops, "unet_tflops_frac_of_sustained_peak": tflops / (tf_sus * world),
            "flop_count": f"{gf_fwd:.2f} GF per image-forward (reference count, cond encoder not hoisted) x {fwd} forwards/image",
            "mixed_roofline": {"ideal_us_per_image_forward": ideal_us, "achieved_us_per_image_forward": us_per_fwd, "frac": ideal_us / us_per_fwd,
                               "ceiling_img_s_per_gpu": 1e6 / (ideal_us * fwd),
                               "definition": "sum over layers of max(flops/sustained bf16 peak, min bytes/measured HBM bandwidth), "
                                             "layer table in localdiffusion_hallucination_b200/workload.py (SURVEY.md 8d)"},
            "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e,
            "roofline": None if a.no_roofline else dominant_kernel_roofline(lib, dev, a, hbm, src),
            "peaks": {"hbm_gbs": hbm, "bf16_tflops_burst": tf_burst, "bf16_tflops_sustained": tf_sus, "source": src},
        }
        if world == 1 and not a.no_cpu_baseline:
            del gd, model
            torch.cuda.empty_cache()
            cores = os.cpu_count() or 1
            v, sample = CpuLeg(a, cores).rate(3)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def dominant_kernel_roofline(lib, dev, a, hbm, src):
    """Roofline of the dominant kernel family -- the tcgen05 3x3 convolution with 32 output channels at full resolution
    (the largest share of device time, profiles/*_launch_shares.md) -- timed alone with CUDA events on its launch stream, in every
    variant the sampler launches it in, weighted by how often one UNet forward launches each at full resolution:
      plain x1 (ups[-1] conv), +GroupNorm statistics x2 (block1 of the two down ResnetBlocks), +normalise-on-load and statistics x5
      (every block2), dual 3x3 + 1x1 over a 64-channel virtual concat with statistics x3 (block1 + res_conv of the up / final blocks).
    Algorithmic bytes per launch = N*H*W*(Cin+Cout_total)*2 (bf16 in + out, each touched once) + weights."""
    import ctypes as C

    import torch

    if a.model != "mri":
        return None
    N, S = 2 * a.batch, a.size
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    variants = [("plain", 0, 32, 0, 1, 1), ("stats", 1, 32, 0, 1, 2), ("normalise_on_load+stats", 2, 32, 0, 1, 5), ("dual+stats", 3, 32, 32, 2, 3)]
    rows, tot_b, tot_t = [], 0.0, 0.0
    for name, var, c0, c1, nout, weight in variants:
        ms = C.c_float(0)
        if lib.ld_debug_conv_variant_time(var, c0, c1, N, S, S, 32, 20, C.byref(ms), st) != 0:
            return None
        byts = N * S * S * (c0 + c1 + 32 * nout) * 2 + (9 + (1 if var == 3 else 0)) * (c0 + c1) * 32 * 2
        rows.append({"variant": name, "launches_per_forward": weight, "ms_per_launch": ms.value, "bytes_per_launch": byts,
                     "gbs": byts / (ms.value / 1e3) / 1e9, "frac": byts / (ms.value / 1e3) / 1e9 / hbm})
        tot_b += weight * byts
        tot_t += weight * ms.value / 1e3
    ach = tot_b / tot_t / 1e9
    traffic, tsrc = None, None
    tp = os.path.join(ROOT, "profiles", "conv32_traffic.json")
    if N == 32 and S == 256 and os.path.isfile(tp):  # ncu captures taken on exactly these launch shapes
        t = json.load(open(tp))
        traffic, tsrc = t["traffic_bytes_per_launch_weighted"], t["source"]
    return {"kernel": "conv_tc_kernel<32,3,32> (3x3 tcgen05 conv, 32 output channels, %dx%dx%d; swizzled TMA in, register or TMA stores out), launch-weighted over its "
                      "four in-situ variants" % (N, S, S), "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
            "traffic": traffic, "traffic_source": tsrc, "variants": rows, "peak_source": src,
            "algorithmic_bytes_per_launch": tot_b / sum(v[5] for v in variants)}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
