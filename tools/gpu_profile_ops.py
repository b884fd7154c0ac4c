"""Per-op device times of one UNet forward inside its real launch sequence (development aid).
usage: LD_PROFILE_OPS=60 python tools/gpu_profile_ops.py [N] [S]   (prints `LDPROF` lines on stderr for the last call)"""
import os
import sys

import torch

sys.path.insert(0, ".")
os.environ.setdefault("LD_PROFILE_OPS", "60")
from tests import util  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
m = util.make_model("mri", "bf16", device="cuda:0")
g = torch.Generator().manual_seed(0)
x = torch.randn(N, 1, S, S, generator=g).cuda()
cond = torch.rand(N, 1, S, S, generator=g).cuda() * 4
t = torch.full((N,), 500, dtype=torch.long).cuda()
for it in range(3):
    print(f"=== call {it}", file=sys.stderr, flush=True)
    y = m(x, cond, t)
    torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
