"""Seeded synthetic inputs shared by the golden generator, the tests and bench.py.

Everything is generated on the CPU with explicit `torch.Generator`s so the same tensors can be
rebuilt on the GPU box without shipping them.
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


def weight_checksum(sd):
    """Order-independent fingerprint of a state_dict (fp64 sums) to pin seed-derived weights."""
    tot = 0.0
    sq = 0.0
    for k in sorted(sd):
        v = sd[k].double()
        tot += float(v.sum())
        sq += float((v * v).sum())
    return [len(sd), tot, sq]
