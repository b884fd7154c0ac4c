"""Per-op device times of one UNet forward inside its real launch sequence (development aid).
usage: LD_PROFILE_OPS=60 python tools/gpu_profile_ops.py [N] [S] [model] [calls]
(prints `LDPROF` lines on stderr for every call; without LD_PROFILE_OPS it just runs `calls` forwards, e.g. under ncu)"""
import os
import sys

import torch

sys.path.insert(0, ".")
from localdiffusion_hallucination_b200 import Unet  # noqa: E402

MODEL_KW = {
    "mri": dict(dim=32, init_dim=32, mode="mri"),
    "mri_attn8": dict(dim=32, init_dim=32, mode="mri", full_attn=(False, False, True, True), attn_heads=8),
    "mnist": dict(dim=32, init_dim=32, dim_mults=(1, 2, 4), full_attn=(False, False, True), mode="mnist"),
}
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
name = sys.argv[3] if len(sys.argv) > 3 else "mri"
calls = int(sys.argv[4]) if len(sys.argv) > 4 else 3
torch.manual_seed(0)
m = Unet(**MODEL_KW[name], precision="bf16").to("cuda:0").eval()
g = torch.Generator().manual_seed(0)
x = torch.randn(N, 1, S, S, generator=g).cuda()
cond = torch.rand(N, 1, S, S, generator=g).cuda() * 4
t = torch.full((N,), 500, dtype=torch.long).cuda()
for it in range(calls):
    print(f"=== call {it}", file=sys.stderr, flush=True)
    y = m(x, cond, t)
    torch.cuda.synchronize()
    print(f"launches so far {m.launch_count()}", file=sys.stderr, flush=True)
print("ok", float(y.abs().mean()))
