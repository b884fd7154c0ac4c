"""Golden fixtures for the stages in front of the sampler, generated from the reference's OWN code:

  * `data.MNIST.__getitem__` (data.py:814-836) and `data.MedDataset_png.__getitem__` (data.py:400-442) are instantiated and called;
  * the anomaly-map -> mask block of `test.py` (lines 246-247 resize, 251-375 rules, 379-381 manual override) lives inline in the
    script's `__main__`, so its source lines are read from /root/reference at generation time and exec'd in a scratch namespace
    (nothing is copied into this repository);
  * `models.PatchcoreModel.euclidean_dist` / `nearest_neighbors` (models.py:179-217) are called on synthetic banks.

    python tests/golden/make_golden_producers.py          # build container only

Every case is also pushed through oracle/ld_producers.py and asserted equal.  Inputs are rebuilt from seeds by
`tests/producer_cases.py`; only outputs (or sub-samples + checksums of large ones) are stored in golden_producers.npz.
"""
import os
import sys
import tempfile
import textwrap
import types

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ld_producers as lp  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402
from tests import producer_cases as pc  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def ref_lines(path, first, last, dedent):
    src = open(os.path.join(rh.REF_ROOT, path)).read().splitlines()[first - 1:last]
    return "\n".join(l[dedent:] if l.strip() else "" for l in src)


def run_reference_mask_block(amap, cfg, manual):
    """test.py:246-247 + 251-375 (+ 379-381) exec'd on `anomaly_map`."""
    ns = dict(torch=torch, F=F, config=cfg, anomaly_map=amap.clone(), cls=torch.zeros(1), print=lambda *a, **k: None)
    exec(ref_lines("test.py", 246, 247, 20), ns)
    exec(ref_lines("test.py", 251, 375, 16), ns)
    if manual:
        exec(ref_lines("test.py", 379, 381, 16), ns)
    return ns["mask_pred"], ns["binary_mask"], ns["anomaly_map"]


def main():
    assert rh.available()
    rh.load_reference()
    import data  # the reference's data.py (already importable through the harness)
    import models  # the reference's models.py

    G = {}
    # ---- MNIST pair -----------------------------------------------------------------------------------------------------
    raw = pc.mnist_raw()
    ds = data.MNIST({"augmentations": False}, [r.numpy() for r in raw], [0] * len(raw), train=False)
    hr, cond = [], []
    for i in range(len(raw)):
        a, b, _ = ds[i]
        hr.append(a); cond.append(b)
    hr, cond = torch.stack(hr), torch.stack(cond)
    o_hr, o_cond = lp.mnist_pair(raw)
    assert torch.equal(hr, o_hr) and torch.equal(cond, o_cond), "oracle mnist_pair differs from data.MNIST"
    G["mnist_hr"], G["mnist_cond"] = hr.numpy(), cond.numpy()
    print("mnist pair ok", tuple(cond.shape), float(cond.max()))

    # ---- MRI normalise + translate zero ------------------------------------------------------------------------------------
    from PIL import Image
    tmp = tempfile.mkdtemp()
    t1, fl = pc.mri_raw()
    files = []
    for i in range(t1.shape[0]):
        f = os.path.join(tmp, f"s{i}_flair.png")
        Image.fromarray(fl[i].numpy().astype(np.uint16)).save(f)
        Image.fromarray(t1[i].numpy().astype(np.uint16)).save(f.replace("flair", "t1"))
        np.save(f.replace("_flair.png", "_seg.npy"), np.zeros((240, 240), np.float32))
        files.append(f)
    for tz in (True, False):
        cfg = dict(pc.MRI_CFG, translate_zero=tz, augmentations=False)
        dsm = data.MedDataset_png(cfg, files, train=False, tumor=False, mode="flair")
        assert len(dsm) == t1.shape[0]
        outs_f, outs_t = [], []
        for i in range(len(dsm)):
            f_s, t_s, _ = dsm[i]
            outs_f.append(f_s); outs_t.append(t_s)
        outs_f, outs_t = torch.stack(outs_f), torch.stack(outs_t)
        o_f = lp.mri_normalize(fl.unsqueeze(1), cfg["mean_flair"], cfg["std_flair"], tz, 224)
        o_t = lp.mri_normalize(t1.unsqueeze(1), cfg["mean_t1"], cfg["std_t1"], tz, 224)
        assert torch.equal(outs_f, o_f) and torch.equal(outs_t, o_t), "oracle mri_normalize differs from data.MedDataset_png"
        tag = "tz" if tz else "raw"
        G[f"mri_{tag}_t1_sub4"] = outs_t[:, :, ::4, ::4].contiguous().numpy()
        G[f"mri_{tag}_t1_sums"] = np.array([float(outs_t.double().sum()), float((outs_t.double() ** 2).sum()), float(outs_t.min()), float(outs_t.max())])
    # min_max_val (test.py:17-37): the reference function is importable only through test.py's script body -> exec its def
    ns = dict(torch=torch)
    exec(ref_lines("test.py", 17, 37, 0), ns)
    for tz in (True, False):
        cfg = dict(pc.MRI_CFG, translate_zero=tz)
        got = ns["set_min_max_val"](cfg, "mri")
        mine = lp.min_max_val(cfg, "mri")
        assert all(float(a) == float(b) for a, b in zip(got, mine)), (got, mine)
        G[f"minmax_{'tz' if tz else 'raw'}"] = np.array([float(v) for v in got])
    assert ns["set_min_max_val"]({}, "mnist") == lp.min_max_val({}, "mnist")
    print("mri normalise ok")

    # ---- anomaly map -> masks -------------------------------------------------------------------------------------------------
    n_active = 0
    for name, rule, cfg, amap, img_size, manual in pc.mask_cases():
        mp, bm, a_res = run_reference_mask_block(amap, cfg, manual)
        o_mp, o_bm = lp.masks_from_anomaly(amap, rule, img_size, 7 if manual else 0)
        assert torch.equal(mp, o_mp) and torch.equal(bm, o_bm), name
        G[f"mask_{name}_pred"], G[f"mask_{name}_bin"] = mp.numpy(), bm.numpy()
        n_active += int(bool((mp != 1).any()))
        print(f"  mask {name}: amax {float(a_res.max()):.2f} ones {float((mp == 1).float().mean()):.3f} binary {float(bm.mean()):.3f}")
    assert n_active >= 12

    # ---- PatchCore kNN ------------------------------------------------------------------------------------------------------------
    for name, x, bank in pc.knn_cases():
        d = models.PatchcoreModel.euclidean_dist(x, bank)
        fake = types.SimpleNamespace(memory_bank=bank, euclidean_dist=models.PatchcoreModel.euclidean_dist)
        sc, loc = models.PatchcoreModel.nearest_neighbors(fake, x, 1)
        o_sc, o_loc = lp.knn_min(x, bank)
        assert torch.equal(sc, o_sc) and torch.equal(loc, o_loc) and torch.equal(d.min(1).values, sc)
        G[f"knn_{name}_score"], G[f"knn_{name}_loc"] = sc.numpy(), loc.numpy()
        print(f"  knn {name}: x {tuple(x.shape)} bank {tuple(bank.shape)} score range [{float(sc.min()):.3f}, {float(sc.max()):.3f}]")

    path = os.path.join(OUT, "golden_producers.npz")
    np.savez_compressed(path, **G)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
