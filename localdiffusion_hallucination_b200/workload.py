"""Seeded synthetic workloads of the sampler: model constructor arguments of the BASELINE configurations,
the reference's sampler-relevant config keys, conditional images / masks / noise tapes, and the per-layer work model the
bench's roofline figures are computed from.

Everything is generated on the CPU with explicit `torch.Generator`s so the same tensors can be rebuilt on any box without
shipping them.  Used by bench.py, __graft_entry__.smoke() and (re-exported through tests/golden/cases.py) by the tests; nothing
here touches `oracle/`.
"""

import math

import torch

MODEL_KW = {
    "mnist": dict(dim=32, init_dim=32, dim_mults=(1, 2, 4), full_attn=(False, False, True), mode="mnist"),
    "mri": dict(dim=32, init_dim=32, mode="mri"),
    "mri_attn8": dict(dim=32, init_dim=32, mode="mri", full_attn=(False, False, True, True), attn_heads=8),
}


def base_config(data="mnist", start_timestep=2, **kw):
    """Mirror of the sampler-relevant keys of the reference's config.yaml:18-35."""
    d = dict(branch_out=True, start_intermediate=True, start_timestep=start_timestep, mask_x=True, mask_cond=False,
             ood_AD=True, ood_confidence=False, classifier=False, use_gt=False, use_gt_timestep=100, data=data)
    d.update(kw)
    return d


def cond_uniform(B, S, seed=1, scale=2.0):
    g = torch.Generator().manual_seed(seed)
    return scale * torch.rand(B, 1, S, S, generator=g)


def mask_left_columns(B, S, cols=8):
    """Analogue of the manual mask of test.py:379-381 (left columns == 1)."""
    m = torch.zeros(B, 1, S, S)
    m[:, :, :, :cols] = 1.0
    return m


def noise_tape(B, S, steps, seed=10):
    """x_T followed by one draw per step, the order the reference consumes them (ddpm.py:935, 852/857)."""
    g = torch.Generator().manual_seed(seed)
    return torch.stack([torch.randn(B, 1, S, S, generator=g) for _ in range(steps)])


def _blur(x, sigma):
    k = int(3 * sigma) * 2 + 1
    ax = torch.arange(k, dtype=torch.float32) - k // 2
    w = torch.exp(-0.5 * (ax / sigma) ** 2)
    w = (w / w.sum()).view(1, 1, 1, k)
    x = torch.nn.functional.conv2d(x, w, padding=(0, k // 2))
    return torch.nn.functional.conv2d(x, w.transpose(2, 3), padding=(k // 2, 0))


def mri_like(B, S, seed=2, max_t1=4.02):
    """Synthetic 'T1' conditional image with an injected OOD blob + soft anomaly mask.

    cond: blurred noise mapped to [0, max_t1] inside a centred ellipse, 0 outside, +1.5 inside a disc.
    mask: soft values in [0,1) outside the disc and exactly 1.0 inside it (like test.py:301-304).
    """
    g = torch.Generator().manual_seed(seed)
    n = torch.randn(B, 1, S, S, generator=g)
    sm = _blur(n, max(1.0, S / 32.0))
    sm = (sm - sm.amin(dim=(2, 3), keepdim=True)) / (sm.amax(dim=(2, 3), keepdim=True) - sm.amin(dim=(2, 3), keepdim=True) + 1e-8)
    yy, xx = torch.meshgrid(torch.arange(S, dtype=torch.float32), torch.arange(S, dtype=torch.float32), indexing="ij")
    cy = cx = (S - 1) / 2
    ell = (((yy - cy) / (0.42 * S)) ** 2 + ((xx - cx) / (0.36 * S)) ** 2) <= 1.0
    cond = sm * max_t1 * ell
    g3 = torch.Generator().manual_seed(seed + 1)
    mask = torch.zeros(B, 1, S, S)
    for b in range(B):
        r = int(torch.randint(max(2, S * 12 // 256), max(3, S * 32 // 256) + 1, (1,), generator=g3))
        oy = int(torch.randint(int(0.3 * S), int(0.7 * S), (1,), generator=g3))
        ox = int(torch.randint(int(0.3 * S), int(0.7 * S), (1,), generator=g3))
        disc = ((yy - oy) ** 2 + (xx - ox) ** 2) <= r * r
        cond[b, 0] = torch.where(disc, torch.clamp(cond[b, 0] + 1.5, max=max_t1), cond[b, 0])
        soft = torch.exp(-(((yy - oy) ** 2 + (xx - ox) ** 2).sqrt() - r).clamp(min=0) / (0.05 * S)) * 0.98
        mask[b, 0] = torch.where(disc, torch.ones(()), soft ** 2)
    return cond, mask


MRI_MIN_MAX = (0.0, 4096.0 / 386.31912016662903, 0.0)  # test.py:24-29 with config.yaml:59-60 (translate_zero)
MNIST_MIN_MAX = (0.0, 2.0)  # test.py:30-33


def make_model(name, precision="fp32", seed=0, device=None, **opts):
    """Product `Unet` of a named configuration with the seed-`seed` default initialisation (bit-identical to what the
    reference constructs after `torch.manual_seed(seed)`, tests/golden/make_golden.py)."""
    from .unet import Unet

    torch.manual_seed(seed)
    m = Unet(**MODEL_KW[name], precision=precision)
    if opts:
        m.set_engine_options(**opts)
    if device is not None:
        m = m.to(device)
    return m.eval()


# ----------------------------------------------------------------------------------------------------------------------
# work model: algorithmic FLOPs and minimum HBM bytes of every layer of one image-forward
# ----------------------------------------------------------------------------------------------------------------------
def unet_layers(name, S, include_cond=True):
    """[(layer, flops, min_bytes)] of one image-forward of `Unet.forward` (ddpm.py:404-451) at S x S.

    FLOPs: 2 per MAC, convolutions and batched matmuls only (the reference count of BASELINE.md §3).  Minimum HBM bytes: every
    convolution reads its input and writes its output once in bf16 and its weights once (ideal conv-level fusion: norms,
    activations, residual adds, concats and re-sampling folded into the neighbouring convolutions); an attention block reads
    its input and writes its output once, except LinearAttention, whose soft-max over all n pixels of an image forces a second
    read of the input (6 C bytes per pixel)."""
    kw = MODEL_KW[name]
    dim, init_dim = kw["dim"], kw.get("init_dim") or kw["dim"]
    mults = tuple(kw.get("dim_mults", (1, 2, 4, 8)))
    full = tuple(kw.get("full_attn", (False,) * (len(mults) - 1) + (True,)))
    heads, dh = kw.get("attn_heads", 4), 32
    hid = heads * dh
    dims = [init_dim] + [dim * m for m in mults]
    L = len(mults)
    out = []

    def conv(nm, r, ci, co, k, r_in=None):
        px, pin = r * r, (r_in or r) ** 2
        out.append((nm, 2.0 * px * ci * co * k * k, pin * ci * 2.0 + px * co * 2.0 + k * k * ci * co * 2.0))

    def res(nm, r, ci, co):
        conv(nm + ".block1", r, ci, co, 3)
        conv(nm + ".block2", r, co, co, 3)
        if ci != co:
            conv(nm + ".res_conv", r, ci, co, 1)

    def attn(nm, r, c, is_full):
        n = r * r
        if is_full:   # ddpm.py:271-282, attend.py:98-113
            fl = 2.0 * n * c * 3 * hid + 2 * (2.0 * n * n * dh * heads) + 2.0 * n * hid * c
            out.append((nm + ".attn", fl, n * c * 2.0 * 2))
        else:         # ddpm.py:234-251
            fl = 2.0 * n * c * 3 * hid + 2 * (2.0 * n * dh * dh * heads) + 2.0 * n * hid * c
            out.append((nm + ".linattn", fl, n * c * 2.0 * 3))

    out.append(("init_conv", 2.0 * S * S * 49 * init_dim, S * S * (4.0 + 2.0 * init_dim)))
    r = S
    for i in range(L):
        di, dn = dims[i], dims[i + 1]
        res(f"downs.{i}.0", r, di, di)
        res(f"downs.{i}.1", r, di, di)
        attn(f"downs.{i}.2", r, di, full[i])
        if i < L - 1:
            conv(f"downs.{i}.3", r // 2, 4 * di, dn, 1)   # pixel-unshuffle + 1x1 (ddpm.py:120-124): reads r*r*di
            r //= 2
        else:
            conv(f"downs.{i}.3", r, di, dn, 3)
    mid = dims[-1]
    res("mid_block1", r, mid, mid)
    attn("mid_attn", r, mid, True)
    res("mid_block2", r, mid, mid)
    if include_cond:   # ResUnet (unet_model.py:91-137): time-invariant; the reference recomputes it every forward (ddpm.py:434)
        rc = S
        deep = kw["mode"] in ("mri", "mvtec", "mvtecGray")
        blocks = [(1, 32, 32), (32, 32, 64), (64, 64, 128)] + ([(128, 128, 256)] if deep else [])
        for j, (ci, cm, co) in enumerate(blocks):
            conv(f"cond.{j}.a", rc, ci, cm, 3)
            conv(f"cond.{j}.b", rc, cm, co, 3)
            conv(f"cond.{j}.id", rc, ci, co, 3)
            if j < len(blocks) - 1:
                rc //= 2
    res("conv_fusion", r, 2 * mid, mid)
    for i in range(L):
        di, dn = dims[L - 1 - i], dims[L - i]
        res(f"ups.{i}.0", r, dn + di, dn)
        res(f"ups.{i}.1", r, dn + di, dn)
        attn(f"ups.{i}.2", r, dn, full[L - 1 - i])
        if i < L - 1:
            conv(f"ups.{i}.3", 2 * r, dn, di, 3, r_in=r)   # nearest x2 + 3x3 (ddpm.py:114-118): reads the low-resolution tensor
            r *= 2
        else:
            conv(f"ups.{i}.3", r, dn, di, 3)
    res("final_res_block", r, 2 * dim, dim)
    out.append(("final_conv", 2.0 * r * r * dim, r * r * (dim * 2.0 + 4.0)))
    return out


def mixed_roofline_us(name, S, tflops, gbs, include_cond=True):
    """Ideal time (us) of one image-forward: sum over layers of max(flops / tensor peak, bytes / HBM bandwidth)."""
    return sum(max(f / (tflops * 1e12), b / (gbs * 1e9)) for _, f, b in unet_layers(name, S, include_cond)) * 1e6


def forward_gflop(name, S, include_cond=True):
    return sum(f for _, f, _ in unet_layers(name, S, include_cond)) / 1e9
