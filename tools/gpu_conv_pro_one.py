"""One normalise-on-load + statistics conv (ld_debug_conv_fused) at full size, for ncu captures (development aid)."""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
from localdiffusion_hallucination_b200 import _lib
lib = _lib.lib(); dev = torch.device("cuda:0")
N, H, W, C0, Cout, G = 32, 256, 256, 32, 32, 8
g = torch.Generator().manual_seed(0)
x = torch.randn(N, H, W, C0, generator=g).to(dev)
w = (torch.randn(Cout, C0, 3, 3, generator=g) / (C0 * 9) ** 0.5).contiguous(); b = torch.zeros(Cout)
xg = x.double().view(N, H * W, G, C0 // G)
st = torch.stack([xg.sum(dim=(1, 3)), (xg * xg).sum(dim=(1, 3))], dim=-1).contiguous()
gamma, beta, film = torch.ones(C0).to(dev), torch.zeros(C0).to(dev), torch.zeros(N, 2 * C0).to(dev)
out = torch.empty(N, H, W, Cout, device=dev); stats = torch.zeros(N, G, 2, dtype=torch.float64, device=dev)
for _ in range(3):
    rc = lib.ld_debug_conv_fused(x.data_ptr(), C0, N, H, W, w.data_ptr(), b.data_ptr(), Cout, st.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                 film.data_ptr(), 2 * C0, G, 1, stats.data_ptr(), G, out.data_ptr(), None)
torch.cuda.synchronize(); print("rc", rc)
