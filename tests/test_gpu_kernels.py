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
    (128, 0, 64, 3, 3, 24, 40, 1, False),     # up-sampling conv, two 64-channel chunks, ragged tiles (H, W = 24, 40)
    (256, 0, 128, 3, 1, 16, 16, 1, False),    # up-sampling conv of the deepest level (four chunks, direct stores)
    (32, 0, 32, 3, 2, 20, 12, 1, False),      # up-sampling conv, 32-channel chunk, H, W not multiples of the tile
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


@pytest.mark.parametrize("case", [c for c in CONV_CASES if c[7] == 1] + [(64, 0, 32, 3, 2, 64, 48, 1, False), (128, 0, 64, 3, 1, 40, 24, 1, False)])
def test_folded_upsampling_conv_matches_torch(case):
    """nearest x2 + 3x3 (ddpm.py:114-118) as ONE low-resolution convolution with 4 * Cout parity channels and a pixel-shuffle
    epilogue (conv_tc_pack_up2): compared with torch on the up-sampled image.  The four 2x2 filters are summed in fp32 and rounded
    to bf16 once, so the reference is the un-rounded-weight convolution with a bf16-level tolerance."""
    C0, C1, Cout, ks, N, H, W, up, use_res = case
    g = torch.Generator().manual_seed(C0 * 5 + Cout + H)
    dev = torch.device("cuda:0")
    x0 = torch.randn(N, H // 2, W // 2, C0, generator=g).to(dev)
    w = (torch.randn(Cout, C0, 3, 3, generator=g) / (C0 * 9) ** 0.5).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    out = run_conv(3, x0, None, w, b, None, 1, H, W)
    x = F.interpolate(x0.bfloat16().float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=1).permute(0, 2, 3, 1)
    assert util.rel_err(out, ref) < 5e-3
    old = run_conv(2, x0, None, w, b, None, 1, H, W)      # the replicate-on-load path it replaces
    assert util.rel_err(out, old) < 5e-3


# (C0, Cout, N, H, W, pro_G (0 = no prologue), act, film, stats_G (0 = none))
FUSED_CASES = [
    (32, 32, 2, 32, 32, 8, 1, True, 8),      # ResnetBlock block2 at dim 32 (ddpm.py:174-185)
    (64, 64, 1, 20, 12, 8, 1, True, 8),      # ragged tiles
    (128, 128, 2, 16, 16, 8, 1, False, 8),   # streamed weights, 16 channels per group
    (256, 256, 1, 8, 8, 8, 1, True, 8),      # 32 channels per group
    (32, 32, 3, 16, 24, 16, 2, False, 16),   # BasicBlock (unet_model.py:20-25): 16 groups, ReLU
    (32, 64, 2, 16, 16, 0, 0, False, 16),    # statistics only
    (64, 32, 40, 16, 8, 8, 1, True, 0),      # prologue only, many images per persistent CTA
]


@pytest.mark.parametrize("case", FUSED_CASES)
def test_conv_fused_groupnorm_prologue_and_stats(case):
    C0, Cout, N, H, W, pG, act, use_film, sG = case
    lib = _lib.lib()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(C0 + 3 * Cout + H + pG)
    x = (torch.randn(N, H, W, C0, generator=g) * 1.7 + 0.3).bfloat16().float()
    w = (torch.randn(Cout, C0, 3, 3, generator=g) / (C0 * 9) ** 0.5)
    b = torch.randn(Cout, generator=g)
    gamma, beta = torch.rand(C0, generator=g) + 0.5, torch.randn(C0, generator=g) * 0.2
    film = torch.randn(N, 2 * C0, generator=g) * 0.3
    xin = x
    args = dict(ps=None, ga=None, be=None, fi=None)
    if pG:
        xg = x.double().view(N, H * W, pG, C0 // pG)
        st = torch.stack([xg.sum(dim=(1, 3)), (xg * xg).sum(dim=(1, 3))], dim=-1).contiguous()  # [N, G, 2]
        cnt = H * W * (C0 // pG)
        mean, var = st[..., 0] / cnt, st[..., 1] / cnt - (st[..., 0] / cnt) ** 2
        rstd = 1.0 / torch.sqrt(var + 1e-5)
        xn = (xg - mean[:, None, :, None]) * rstd[:, None, :, None]
        xn = xn.view(N, H, W, C0) * gamma.double() + beta.double()
        if use_film:
            xn = xn * (film[:, None, None, :C0].double() + 1.0) + film[:, None, None, C0:].double()
        xn = torch.nn.functional.silu(xn) if act == 1 else (torch.relu(xn) if act == 2 else xn)
        xin = xn.float().bfloat16().float()
        args = dict(ps=st.to(dev), ga=gamma.to(dev), be=beta.to(dev), fi=film.to(dev) if use_film else None)
    ref = F.conv2d(xin.permute(0, 3, 1, 2).double(), w.bfloat16().double(), b.double(), padding=1).permute(0, 2, 3, 1)
    out = torch.empty(N, H, W, Cout, device=dev)
    stats = torch.full((N, max(sG, 1), 2), 7.0, dtype=torch.float64, device=dev)
    xd = x.to(dev)
    ptr = lambda t: t.data_ptr() if t is not None else None
    rc = lib.ld_debug_conv_fused(xd.data_ptr(), C0, N, H, W, w.contiguous().data_ptr(), b.contiguous().data_ptr(), Cout,
                                 ptr(args["ps"]), ptr(args["ga"]), ptr(args["be"]), ptr(args["fi"]), 2 * C0, pG, act,
                                 stats.data_ptr() if sG else None, sG, out.data_ptr(),
                                 C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc)
    # prologue output is re-rounded to bf16 (and SiLU uses tanh.approx): one extra bf16 ulp on some operands
    assert util.rel_err(out, ref) < (8e-3 if pG else 4e-3)
    if sG:
        rg = ref.view(N, H * W, sG, Cout // sG)
        want = torch.stack([rg.sum(dim=(1, 3)), (rg * rg).sum(dim=(1, 3))], dim=-1)
        got = stats.cpu()
        scale = want[..., 1].sqrt().unsqueeze(-1) * (H * W * (Cout // sG)) ** 0.5  # |sum| <= sqrt(n * sumsq)
        assert float(((got[..., 0] - want[..., 0]).abs() / scale[..., 0]).max()) < 2e-3
        assert float(((got[..., 1] - want[..., 1]).abs() / want[..., 1]).max()) < 5e-3


# (C0, C1, Cout, N, H, W, stats_G)
DUAL_CASES = [
    (32, 32, 32, 2, 32, 32, 8),     # up-path ResnetBlock(64 -> 32): virtual concat of two 32-channel sources
    (64, 32, 64, 1, 20, 12, 8),     # ResnetBlock(96 -> 64), ragged tiles, KC = 32 packing
    (64, 0, 32, 3, 16, 24, 8),      # single source, KC = 64
    (64, 64, 64, 37, 16, 8, 8),     # many images per persistent CTA
]


@pytest.mark.parametrize("case", DUAL_CASES)
def test_conv_dual_block1_and_res_conv(case):
    """block1.proj + GroupNorm statistics and res_conv of the same input from one launch (ddpm.py:207,212)."""
    C0, C1, Cout, N, H, W, sG = case
    lib = _lib.lib()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(C0 + 3 * C1 + 5 * Cout + H)
    Cin = C0 + C1
    x = (torch.randn(N, H, W, Cin, generator=g) * 1.3 + 0.2).bfloat16().float()
    w3 = torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5
    w1 = torch.randn(Cout, Cin, generator=g) / Cin ** 0.5
    b3, b1 = torch.randn(Cout, generator=g), torch.randn(Cout, generator=g)
    xc = x.permute(0, 3, 1, 2).double()
    ref = F.conv2d(xc, w3.bfloat16().double(), b3.double(), padding=1).permute(0, 2, 3, 1)
    ref2 = F.conv2d(xc, w1.bfloat16().double()[:, :, None, None], b1.double()).permute(0, 2, 3, 1)
    x0 = x[..., :C0].contiguous().to(dev)
    x1 = x[..., C0:].contiguous().to(dev) if C1 else None
    out, out2 = torch.empty(N, H, W, Cout, device=dev), torch.empty(N, H, W, Cout, device=dev)
    stats = torch.full((N, sG, 2), 7.0, dtype=torch.float64, device=dev)
    rc = lib.ld_debug_conv_dual(x0.data_ptr(), C0, x1.data_ptr() if C1 else None, C1, N, H, W, w3.contiguous().data_ptr(),
                                b3.data_ptr(), w1.contiguous().data_ptr(), b1.data_ptr(), Cout, stats.data_ptr(), sG,
                                out.data_ptr(), out2.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc)
    assert util.rel_err(out, ref) < 4e-3
    assert util.rel_err(out2, ref2) < 4e-3
    rg = ref.view(N, H * W, sG, Cout // sG)
    want = torch.stack([rg.sum(dim=(1, 3)), (rg * rg).sum(dim=(1, 3))], dim=-1)
    got = stats.cpu()
    scale = want[..., 1].sqrt().unsqueeze(-1) * (H * W * (Cout // sG)) ** 0.5
    assert float(((got[..., 0] - want[..., 0]).abs() / scale[..., 0]).max()) < 2e-3
    assert float(((got[..., 1] - want[..., 1]).abs() / want[..., 1]).max()) < 5e-3


# (C, N, HW)
LINATTN_CASES = [(32, 2, 1024), (64, 1, 784), (128, 2, 256), (32, 3, 4096), (64, 5, 1000),
                 (32, 6, 65536),   # 256 x 256: ~20 stages per CTA, every ring wraps several times (a 3-deep raw ring hung here)
                 (64, 20, 4096)]   # many images: flat / sliced tile lists cross image boundaries


# 8 heads (BASELINE configs[3]): two head groups of four; (C, N, HW)
LINATTN8_CASES = [(32, 2, 1024), (64, 3, 4096), (32, 5, 16384), (64, 1, 1000), (32, 3, 65536), (64, 20, 4096)]


@pytest.mark.parametrize("C_,N,HW,heads", [c + (4,) for c in LINATTN_CASES] + [c + (8,) for c in LINATTN8_CASES])
def test_fused_linear_attention_matches_torch(C_, N, HW, heads):
    """attn(x) + x of LinearAttention (ddpm.py:214-251, 425) through the fused tcgen05 kernels."""
    lib = _lib.lib()
    dev = torch.device("cuda:0")
    hid = heads * 32
    g = torch.Generator().manual_seed(C_ + N + HW + heads)
    x = (torch.randn(N, HW, C_, generator=g) * 1.3).bfloat16().float()
    wqkv = torch.randn(3 * hid, C_, generator=g) / C_ ** 0.5
    gn = torch.rand(C_, generator=g) + 0.5
    wout = torch.randn(C_, hid, generator=g) / hid ** 0.5
    bout = torch.randn(C_, generator=g) * 0.1
    g2 = torch.rand(C_, generator=g) + 0.5
    rd = dev if N * HW > 200000 else torch.device("cpu")   # plain torch fp64 reference (on the GPU for the big cases)
    xd = x.double().to(rd)
    xn = F.normalize(xd, dim=-1) * gn.double().to(rd) * C_ ** 0.5
    q, k, v = (xn @ wqkv.double().to(rd).T).view(N, HW, 3, heads, 32).unbind(dim=2)
    q = q.softmax(dim=-1) * 32 ** -0.5
    k = k.softmax(dim=1)
    ctx = torch.einsum("nphd,nphe->nhde", k, v)
    o = torch.einsum("nhde,nphd->nphe", ctx, q).reshape(N, HW, hid) @ wout.double().to(rd).T + bout.double().to(rd)
    attn = (F.normalize(o, dim=-1) * g2.double().to(rd) * C_ ** 0.5).cpu()
    xd = xd.cpu()
    del q, k, v, o, xn
    out = torch.empty(N, HW, C_, device=dev)
    rc = lib.ld_debug_linattn_h(x.to(dev).data_ptr(), C_, N, HW, heads, wqkv.contiguous().data_ptr(), gn.data_ptr(), wout.contiguous().data_ptr(),
                                bout.data_ptr(), g2.data_ptr(), out.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc)
    got = out.cpu().double() - xd   # attention branch (the residual is exact up to the output rounding)
    assert util.rel_err(got, attn) < 2.5e-2


# (N, n tokens, heads)
ATTN_CASES = [(2, 1024, 4), (1, 64, 4), (3, 49, 4), (1, 300, 8), (2, 4096, 4)]


@pytest.mark.parametrize("N,n,heads", ATTN_CASES)
def test_flash_attention_matches_torch(N, n, heads):
    """softmax(q k^T d^-0.5) v of attend.py:98-113 through the tcgen05 kernel (no [n x n] matrix)."""
    lib = _lib.lib()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(N + n + heads)
    qkv = (torch.randn(N, n, 3 * heads * 32, generator=g) * 1.5).bfloat16().float()
    q, k, v = qkv.double().view(N, n, 3, heads, 32).unbind(dim=2)
    sim = torch.einsum("nihd,njhd->nhij", q, k) * 32 ** -0.5
    ref = torch.einsum("nhij,njhd->nihd", sim.softmax(dim=-1), v).reshape(N, n, heads * 32)
    out = torch.empty(N, n, heads * 32, device=dev)
    rc = lib.ld_debug_attention(qkv.to(dev).data_ptr(), N, n, heads, out.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc)
    assert util.rel_err(out, ref) < 1e-2


@pytest.mark.parametrize("N,H,W,Cout", [(2, 32, 32, 32), (1, 20, 28, 32), (3, 64, 64, 64), (1, 256, 256, 32), (2, 300, 36, 64),
                                        (1, 18, 30, 32)])   # W % 4 != 0: the im2col form
def test_init_conv7_matches_torch(N, H, W, Cout):
    """7x7 single-channel init_conv (ddpm.py:319) on tcgen05: the fp32 state enters as bf16 hi + lo, weights as bf16."""
    lib = _lib.lib()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(N + H + W + Cout)
    x = torch.randn(N, H, W, generator=g) * 2.0
    w = torch.randn(Cout, 1, 7, 7, generator=g) / 7.0
    b = torch.randn(Cout, generator=g)
    out = torch.empty(N, H, W, Cout, device=dev)
    rc = lib.ld_debug_conv7(x.to(dev).data_ptr(), N, H, W, w.contiguous().data_ptr(), b.data_ptr(), Cout, out.data_ptr(),
                            C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc)
    ref = F.conv2d(x[:, None].double(), w.bfloat16().double(), b.double(), padding=3).permute(0, 2, 3, 1)
    assert util.rel_err(out, ref) < 4e-3   # only the bf16 output rounding (+ 2^-17 of the hi/lo split) remains
