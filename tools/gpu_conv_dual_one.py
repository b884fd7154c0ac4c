"""One dual-output conv (block1.proj 64->32 over a virtual concat + res_conv) at full size, for ncu captures (development aid)."""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
from localdiffusion_hallucination_b200 import _lib
lib = _lib.lib(); dev = torch.device("cuda:0")
N, H, W, C0, C1, Cout, G = 32, 256, 256, 32, 32, 32, 8
g = torch.Generator().manual_seed(0)
x0 = torch.randn(N, H, W, C0, generator=g).to(dev); x1 = torch.randn(N, H, W, C1, generator=g).to(dev)
w3 = (torch.randn(Cout, C0 + C1, 3, 3, generator=g) / 24.0).contiguous(); w1 = (torch.randn(Cout, C0 + C1, generator=g) / 8.0).contiguous()
b3, b1 = torch.zeros(Cout), torch.zeros(Cout)
out, out2 = torch.empty(N, H, W, Cout, device=dev), torch.empty(N, H, W, Cout, device=dev)
stats = torch.zeros(N, G, 2, dtype=torch.float64, device=dev)
for _ in range(3):
    rc = lib.ld_debug_conv_dual(x0.data_ptr(), C0, x1.data_ptr(), C1, N, H, W, w3.data_ptr(), b3.data_ptr(), w1.data_ptr(), b1.data_ptr(), Cout,
                                stats.data_ptr(), G, out.data_ptr(), out2.data_ptr(), None)
torch.cuda.synchronize(); print("rc", rc)
