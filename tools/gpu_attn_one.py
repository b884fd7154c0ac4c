"""Flash attention (ld_debug_attention) at bench size, for ncu captures (development aid). usage: gpu_attn_one.py [N] [n] [heads]"""
import sys, time, torch
sys.path.insert(0, ".")
from localdiffusion_hallucination_b200 import _lib
lib = _lib.lib(); torch.zeros(1, device="cuda")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
heads = int(sys.argv[3]) if len(sys.argv) > 3 else 4
g = torch.Generator().manual_seed(1)
qkv = torch.randn(N, n, 3 * heads * 32, generator=g).cuda()
out = torch.empty(N, n, heads * 32, device="cuda")
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rc = lib.ld_debug_attention(qkv.data_ptr(), N, n, heads, out.data_ptr(), None)
    torch.cuda.synchronize()
print(f"attn N={N} n={n} heads={heads} rc={rc} wall {1e3*(time.perf_counter()-t0):.2f} ms (incl. converts)")
