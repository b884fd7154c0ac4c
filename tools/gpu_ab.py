"""In-process A/B of engine options on the bench workload: alternates variants call by call (power / clock drift hits both alike)
and prints the median ms per timestep of each.  usage: gpu_ab.py OPT=v0,v1[,..] [T] [rounds] [B] [S]   e.g. gpu_ab.py pdl=0,1 100 5"""
import statistics, sys
import torch
sys.path.insert(0, ".")
from localdiffusion_hallucination_b200 import GaussianDiffusion, _lib
from localdiffusion_hallucination_b200 import workload as wl

opt, vals = sys.argv[1].split("=")
vals = [int(v) for v in vals.split(",")]
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 5
B = int(sys.argv[4]) if len(sys.argv) > 4 else 16
S = int(sys.argv[5]) if len(sys.argv) > 5 else 256
dev = torch.device("cuda:0")
lib = _lib.lib()
m = wl.make_model("mri", "bf16", device=dev)
gd = GaussianDiffusion(wl.base_config("mri", 2), m, image_size=S, timesteps=T, objective="pred_x0").to(dev)
cond, mask = wl.mri_like(B, S)
cond, mask = cond.to(dev), mask.to(dev)
tape = gd.make_noise_tape((B, 1, S, S), T, dev)
h = m.engine()
res = {v: [] for v in vals}
outs = {}
for r in range(rounds + 1):
    for v in vals:
        _lib.check(lib.ld_set_option(h, opt.encode(), v))
        gd.sample(cond, None, batch_size=B, mask=mask, min_max_val=wl.MRI_MIN_MAX, noise=tape)   # builds plans / graphs for this variant
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        out = gd.sample(cond, None, batch_size=B, mask=mask, min_max_val=wl.MRI_MIN_MAX, noise=tape)
        e1.record(); torch.cuda.synchronize()
        if r > 0: res[v].append(e0.elapsed_time(e1) / T)
        if v in outs and r == rounds: print(f"{opt}={v}: max |diff| between two runs of the same variant:", float((outs[v] - out).abs().max()))
        outs[v] = out
for v in vals:
    print(f"{opt}={v}: median {statistics.median(res[v]):.4f} ms/timestep  (min {min(res[v]):.4f}, max {max(res[v]):.4f}, n={len(res[v])})")
a, b = outs[vals[0]], outs[vals[-1]]
print("max |diff| between first and last variant:", float((a - b).abs().max()))
