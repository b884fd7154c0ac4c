"""The fused posterior / clamp / composite step at C2 size (B=16, 256x256), a few launches, for ncu captures and event timing
(development aid).  usage: gpu_step_one.py [B] [S] [kind]"""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
from localdiffusion_hallucination_b200 import GaussianDiffusion, Unet, _lib
lib = _lib.lib(); dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
kind = int(sys.argv[3]) if len(sys.argv) > 3 else 0
torch.manual_seed(0)
m = Unet(dim=32, init_dim=32, dim_mults=(1, 2, 4), full_attn=(False, False, True), mode="mnist", precision="fp32").to(dev)
cfg = dict(branch_out=True, start_intermediate=True, start_timestep=2, mask_x=True, mask_cond=False, ood_AD=True, ood_confidence=False,
           classifier=False, use_gt=False, use_gt_timestep=100, data="mri")
gd = GaussianDiffusion(cfg, m, image_size=S, timesteps=50, objective="pred_x0").to(dev)
h = m.engine(); gd._push_schedule(h)
n = B * S * S
g = torch.Generator().manual_seed(3)
t = [torch.randn(n, generator=g).to(dev) for _ in range(7)]
mask = (torch.rand(n, generator=g) > 0.7).float().to(dev)
sd = _lib.SampleDesc(); sd.mask_x, sd.ood_uses_cond, sd.cond_in_floor, sd.min_val, sd.max_val = 1, 0, 0.95, 0.0, 2.0
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(4):
    ev0.record()
    rc = lib.ld_posterior_step(h, kind, 7, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(), t[4].data_ptr(), mask.data_ptr(),
                               t[5].data_ptr(), C.byref(sd), n, st)
    ev1.record(); torch.cuda.synchronize()
print("rc", rc, "n", n, "last call (incl. staging copies + prep) ms", ev0.elapsed_time(ev1))
