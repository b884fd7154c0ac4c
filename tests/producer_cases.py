"""Seeded inputs of the producer-stage tests (shared by the golden generator and the tests)."""
import torch

MRI_CFG = dict(mean_t1=104.43, std_t1=1019.2, mean_flair=96.08, std_flair=386.31912016662903, ProjectName="x")


def mnist_raw(n=6, S=28, seed=11):
    """uint8-valued digit-like images [n,S,S] (as floats, what `np2tensor` yields from the IDX arrays)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, 256, (n, S, S), generator=g).float()
    x[0] = 0.0
    x[1, 8:20, 10:18] = 255.0
    return x


def mri_raw(n=2, S=240, seed=12):
    g = torch.Generator().manual_seed(seed)
    t1 = torch.randint(0, 4096, (n, S, S), generator=g).float()
    fl = torch.randint(0, 4096, (n, S, S), generator=g).float()
    t1[0, :40] = 0.0
    return t1, fl


def _amap(h, w, amax, seed, blob=True):
    g = torch.Generator().manual_seed(seed)
    a = 20.0 + 6.0 * torch.rand(1, 1, h, w, generator=g)
    if blob:
        yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
        r2 = (yy - 0.4 * h) ** 2 + (xx - 0.6 * w) ** 2
        a = a + (amax - 23.0) * torch.exp(-r2 / (2 * (0.12 * h) ** 2))
    return a


def mask_cases():
    """(name, oracle rule, reference config, anomaly map, image size or None, manual override)"""
    out = []
    levels = {
        "mnist_8to3": (dict(data="mnist", mnist_cls="8to3", img_size=28), [30.0, 39.0, 42.0, 47.0], 42, 28),
        "mnist_8to5": (dict(data="mnist", mnist_cls="8to5", img_size=28), [50.0, 62.0, 68.0, 75.0], 42, 28),
        "mri_t12flair": (dict(data="mri", ProjectName="mri_t12flair", ood_detector=dict(seg=False)), [40.0, 46.0, 50.0, 55.0, 66.0], 64, None),
        "mri_flair2t1": (dict(data="mri", ProjectName="mri_flair2t1", ood_detector=dict(seg=False)), [40.0, 46.0, 55.0, 66.0], 64, None),
        "mvtec_transistor": (dict(data="mvtec", mvtec_path="/a/b/c/d/transistor/x", img_size=32), [30.0, 34.0, 36.0, 38.0, 43.0], 48, 32),
        "mvtec_toothbrush": (dict(data="mvtec", mvtec_path="/a/b/c/d/toothbrush/x", img_size=32), [30.0, 40.0, 52.0], 48, 32),
        "mvtec_grid": (dict(data="mvtec", mvtec_path="/a/b/c/d/grid/x", img_size=32), [25.0, 30.0, 37.0, 44.0], 48, 32),
    }
    for rule, (cfg, amaxes, h, S) in levels.items():
        for k, amax in enumerate(amaxes):
            # the resize (mnist / mvtec) smooths the peak: overshoot so that the resized maximum lands near the level
            a = _amap(h, h, amax * (1.04 if S else 1.0), seed=100 + 7 * k + len(rule))
            out.append((f"{rule}_{k}", rule, cfg, a, S, False))
    cfg, _, h, S = levels["mnist_8to3"]
    out.append(("manual", "mnist_8to3", cfg, _amap(h, h, 45.0, seed=5), S, True))
    return out


def knn_cases():
    g = torch.Generator().manual_seed(21)
    x1, b1 = torch.randn(256, 96, generator=g), torch.randn(1000, 96, generator=g)
    x2, b2 = torch.rand(200, 128, generator=g) * 3, torch.rand(777, 128, generator=g) * 3
    b2[5] = x2[17]   # an exact hit: distance 0 after the clamp
    return [("gauss", x1, b1), ("hit", x2, b2)]
