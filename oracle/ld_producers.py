"""CPU oracle for the two stages in front of the sampler.

TEST INFRASTRUCTURE ONLY (same rules as oracle/ld_oracle.py): plain torch CPU restatements, each citing the reference lines it
follows.  Pinning: `tests/golden/make_golden_producers.py` runs the reference's own `data.py` dataset classes, the reference's own
mask block of `test.py` (exec'd from the source lines where they lie, nothing copied) and `models.py`'s nearest-neighbour search on
seeded inputs and checks these restatements against them; the outputs are committed as `tests/golden/golden_producers.npz`.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ---- conditional-image producers ------------------------------------------------------------------------------------------
def mnist_pair(raw: Tensor):
    """data.py:814-836 (`MNIST.__getitem__`): raw [N,S,S] in 0..255 -> (hr, cond), both `2 * (x / 255)` (data.py:808-809);
    cond is the sub-sampled image bilinearly up-sampled back to S x S (align_corners=False).  NOTE the reference slices
    `img[:, ::2, ::2]` on the 4-D tensor `[1, 1, S, S]` (data.py:822-826), so only the ROWS are sub-sampled ("28x28 -> 14x14" in its
    comment is really 14 x 28): the degradation is vertical only.  Reproduced as is."""
    img = raw.float().unsqueeze(1)
    down = img[:, :, ::2, :]
    up = F.interpolate(down, size=(img.shape[-1], img.shape[-1]), mode="bilinear", align_corners=False)
    norm = lambda x: 2 * (x / 255.0)
    return norm(img), norm(up)


def center_crop(x: Tensor, size: int) -> Tensor:
    """torchvision CenterCrop on an image at least `size` wide (data.py:380-383)."""
    h, w = x.shape[-2:]
    top, left = int(round((h - size) / 2.0)), int(round((w - size) / 2.0))
    return x[..., top:top + size, left:left + size]


def mri_normalize(raw: Tensor, mean: float, std: float, translate_zero: bool, crop: int) -> Tensor:
    """data.py:400-414 (`MedDataset_png.normalize`) after the centre crop of `transform` (data.py:380-394): per image
    (x - mean) / std, then + |min| when translate_zero."""
    x = center_crop(raw.float(), crop)
    x = (x - mean) / std
    if translate_zero:
        mini = torch.abs(x.flatten(1).min(dim=1).values)
        x = x + mini.view(-1, *([1] * (x.dim() - 1)))
    return x


def min_max_val(config: dict, mode: str = "mri"):
    """test.py:17-37 (`set_min_max_val`): the clamp range handed to `sample()` as (min, max[, min_t1])."""
    if mode == "mri":
        if not config["translate_zero"]:
            mx = (4096 - config["mean_flair"]) / config["std_flair"]
            mn = (0 - config["mean_flair"]) / config["std_flair"]
            mn_t1 = (0 - config["mean_t1"]) / config["std_t1"]
        else:
            mn2 = (0 - config["mean_flair"]) / config["std_flair"]
            mn = 0.0
            mx = (4096 - config["mean_flair"]) / config["std_flair"]
            mx = mx + torch.abs(torch.tensor(mn2))   # an fp32 0-dim tensor in the reference: the sum is rounded to fp32
            mn_t1 = 0.0
        return float(mx), float(mn), float(mn_t1)
    return 2.0, 0.0   # mnist, mvtec


# ---- anomaly map -> masks (test.py:237-381) ----------------------------------------------------------------------------------
RULES = ("mnist_8to3", "mnist_8to5", "mri_t12flair", "mri_flair2t1", "mvtec_transistor", "mvtec_toothbrush", "mvtec_grid")


def masks_from_anomaly(amap: Tensor, rule: str, img_size: int | None = None, manual_cols: int = 0):
    """(mask_pred, binary_mask) from a PatchCore anomaly map [B,1,h,w]: bilinear resize to the image size for mnist / mvtec
    (test.py:254-255), dataset-specific threshold from the map's maximum, soft mask `((clip(a, lo, thr) - min) / (thr - min))**2`
    which is exactly 1.0 wherever a >= thr, all-ones masks when the score is below the gate; `manual_cols > 0` applies the manual
    left-columns mask that test.py:379-381 puts in place of the detector's."""
    a = amap.float()
    if img_size is not None and tuple(a.shape[-2:]) != (img_size, img_size):
        a = F.interpolate(a, size=(img_size, img_size), mode="bilinear", align_corners=False)
    if manual_cols > 0:
        m = torch.zeros_like(a)
        m[:, :, :, :manual_cols] = 1.0
        return m, m
    mx, sd = a.max(), a.std()
    thr = lo = None
    if rule == "mnist_8to3":
        if mx > 37.0:
            thr = 41.7 if mx > 44 else (38.2 if mx > 40.0 else 35.0)
            lo = thr - sd
    elif rule == "mnist_8to5":
        if mx > 58.5:
            thr = 61.0 if mx > 71.0 else (57.0 if mx > 65 else 55.0)
            lo = thr - sd
    elif rule == "mri_t12flair":
        if mx > 43:
            thr = mx - 12 if mx > 60 else (47 if mx > 51 else (44 if mx > 48.5 else 42))
            lo = thr - sd
    elif rule == "mri_flair2t1":
        if mx > 43:
            thr = 47 if mx > 60 else (43 if mx > 50 else 42)
            lo = thr - sd
    elif rule == "mvtec_transistor":
        if mx > 32:
            thr = 33.5 if mx > 40.0 else (mx - 2 * sd if mx > 36.8 else (mx - 1 * sd if mx > 35.0 else 29.5))
            lo = thr - 0.5 * sd
    elif rule == "mvtec_toothbrush":
        if mx > 35:
            thr = 40.0 if mx > 49 else 28.0
            lo = a.min()
    elif rule == "mvtec_grid":
        if mx > 27:
            thr = 35.0 if mx > 40 else (30.0 if mx > 35.0 else 26.5)
            lo = a.min()
    else:
        raise ValueError(rule)
    if thr is None:
        return torch.ones_like(a), torch.ones_like(a)
    binary = (a > thr).float()
    mp = torch.clip(a, min=lo, max=thr)
    mask = ((mp - mp.min()) / (thr - mp.min())) ** 2
    return mask, binary


# ---- PatchCore nearest neighbour (models.py:179-217) ---------------------------------------------------------------------------
def knn_min(embedding: Tensor, memory_bank: Tensor):
    """`euclidean_dist` + `nearest_neighbors(n_neighbors=1)`: sqrt(clamp(|x|^2 - 2 x.y^T + |y|^2, 0)), row minimum and its index."""
    xn = embedding.pow(2).sum(dim=-1, keepdim=True)
    yn = memory_bank.pow(2).sum(dim=-1, keepdim=True)
    d = (xn - 2 * torch.matmul(embedding, memory_bank.transpose(-2, -1)) + yn.transpose(-2, -1)).clamp_min_(0).sqrt_()
    return d.min(1)
