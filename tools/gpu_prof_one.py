"""Launch one kernel family a few times (development aid for `ncu --set full` captures and isolated timings).
usage: gpu_prof_one.py conv|la [N]"""
import ctypes as C
import sys
import time

import torch

sys.path.insert(0, ".")
from localdiffusion_hallucination_b200 import _lib  # noqa

lib = _lib.lib()
torch.zeros(1, device="cuda")
what = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 32
if what == "conv":
    ms = C.c_float(0)
    for (c0, c1, hw, co, ks, up) in [(32, 0, 256, 32, 3, 0), (64, 0, 64, 64, 3, 0), (128, 0, 64, 128, 3, 0), (32, 0, 256, 384, 1, 0)]:
        rc = lib.ld_debug_conv_time(2, c0, c1, N, hw, hw, up, co, ks, 5, C.byref(ms), None)
        by = N * hw * hw * (c0 + c1 + co) * 2
        print(f"conv C{c0}+{c1}->{co} {ks}x{ks} @{hw} N={N}: {ms.value*1000:8.1f} us {by/ms.value/1e6:8.1f} GB/s rc={rc}")
else:
    for (Cc, hw) in [(32, 256), (32, 128), (64, 64), (128, 64)]:
        HW = hw * hw
        g = torch.Generator().manual_seed(1)
        x = torch.randn(N, HW, Cc, generator=g).cuda()
        wqkv = (torch.randn(384, Cc, generator=g) / Cc ** 0.5).contiguous()
        gn, g2 = torch.ones(Cc), torch.ones(Cc)
        wout = (torch.randn(Cc, 128, generator=g) / 128 ** 0.5).contiguous()
        bout = torch.zeros(Cc)
        out = torch.empty_like(x)
        for it in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rc = lib.ld_debug_linattn(x.data_ptr(), Cc, N, HW, wqkv.data_ptr(), gn.data_ptr(), wout.data_ptr(), bout.data_ptr(),
                                      g2.data_ptr(), out.data_ptr(), None)
            torch.cuda.synchronize()
        print(f"la C={Cc} @{hw} N={N}: rc={rc} wall {1e3*(time.perf_counter()-t0):.2f} ms (includes pack + converts)")
