"""GPU: individual kernels through the C-ABI test hooks against plain PyTorch fp32 ops."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from localdiffusion_hallucination_b200 import _lib
from tests import util

pytestmark = pytest.mark.gpu

# (C0, C1, Cout, ks, N, H, W, up, residual)
CONV_CASES = [
    (32, 0, 32, 3, 2, 32, 32, 0, False),
    (64, 0, 64, 3, 1, 16, 24, 0, True),
    (64, 32, 64, 3, 2, 16, 16, 0, False),     # virtual concat, 32-channel chunks
    (128, 128, 128, 3, 1, 8, 8, 0, False),    # virtual concat, 64-channel chunks
    (64, 0, 32, 3, 1, 32, 32, 1, False),      # nearest x2 then 3x3 (ddpm.py:114-118)
    (32, 0, 384, 1, 2, 16, 16, 0, False),     # to_qkv (ddpm.py:227)
    (128, 0, 32, 1, 1, 20, 12, 0, True),      # 1x1 + residual, ragged pixel count
    (256, 256, 256, 3, 1, 8, 8, 0, False),    # conv_fusion block1 shape (ddpm.py:380)
    (32, 0, 32, 3, 1, 20, 12, 0, False),      # H, W not multiples of the tile
    (96 - 32, 32, 64, 1, 1, 8, 8, 0, False),  # res_conv over a concat
]


def run_conv(kernel, x0, x1, w, b, res, up, H, W):
    lib = _lib.lib()
    N, Hin, Win, C0 = x0.shape
    C1 = x1.shape[3] if x1 is not None else 0
    Cout, _, ks, _ = w.shape
    out = torch.empty(N, H, W, Cout, device=x0.device)
    wh, bh = w.cpu().contiguous(), b.cpu().contiguous()
    rc = lib.ld_debug_conv(kernel, x0.data_ptr(), C0, x1.data_ptr() if x1 is not None else None, C1, N, Hin, Win, up, H, W,
                           wh.data_ptr(), bh.data_ptr(), Cout, ks, res.data_ptr() if res is not None else None, out.data_ptr(),
                           C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc)
    return out


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("kernel", [0, 1, 2])
def test_conv_kernels_match_torch(case, kernel):
    C0, C1, Cout, ks, N, H, W, up, use_res = case
    g = torch.Generator().manual_seed(C0 * 7 + Cout + ks + H)
    Hin, Win = (H // 2, W // 2) if up else (H, W)
    dev = torch.device("cuda:0")
    x0 = torch.randn(N, Hin, Win, C0, generator=g).to(dev)
    x1 = torch.randn(N, Hin, Win, C1, generator=g).to(dev) if C1 else None
    w = (torch.randn(Cout, C0 + C1, ks, ks, generator=g) / ((C0 + C1) * ks * ks) ** 0.5).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    res = torch.randn(N, H, W, Cout, generator=g).to(dev) if use_res else None
    out = run_conv(kernel, x0, x1, w, b, res, up, H, W)

    def ref(rnd):
        q = (lambda t: t.bfloat16().float()) if rnd else (lambda t: t)
        x = q(torch.cat([x0] + ([x1] if x1 is not None else []), dim=3)).permute(0, 3, 1, 2)
        if up:
            x = F.interpolate(x, scale_factor=2, mode="nearest")
        y = F.conv2d(x.double(), q(w).double(), b.double(), padding=ks // 2).permute(0, 2, 3, 1)
        if res is not None:
            y = y + q(res).double()
        return y

    if kernel == 0:
        assert util.rel_err(out, ref(False)) < 1e-5
    else:  # bf16 storage: compare against the same op on bf16-rounded operands; only the output rounding remains
        assert util.rel_err(out, ref(True)) < 4e-3
