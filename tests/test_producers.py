"""The stages in front of the sampler: oracle vs the reference's own outputs (CPU), device kernels vs
the oracle and the fixtures (GPU).  Fixtures: tests/golden/golden_producers.npz (tests/golden/make_golden_producers.py)."""
import os

import numpy as np
import pytest
import torch

from localdiffusion_hallucination_b200 import producers
from oracle import ld_producers as lp
from tests import producer_cases as pc
from tests import util

DEV = "cuda:0"


@pytest.fixture(scope="module")
def gp():
    return np.load(os.path.join(util.ROOT, "tests", "golden", "golden_producers.npz"), allow_pickle=False)


# ---- CPU: oracle against the reference's outputs ------------------------------------------------------------------------------
def test_oracle_mnist_pair(gp):
    hr, cond = lp.mnist_pair(pc.mnist_raw())
    assert np.array_equal(hr.numpy(), gp["mnist_hr"]) and np.array_equal(cond.numpy(), gp["mnist_cond"])


def test_oracle_mri_normalize_and_min_max(gp):
    t1, _ = pc.mri_raw()
    for tz, tag in ((True, "tz"), (False, "raw")):
        o = lp.mri_normalize(t1.unsqueeze(1), pc.MRI_CFG["mean_t1"], pc.MRI_CFG["std_t1"], tz, 224)
        assert np.array_equal(o[:, :, ::4, ::4].numpy(), gp[f"mri_{tag}_t1_sub4"])
        np.testing.assert_allclose([float(o.double().sum()), float((o.double() ** 2).sum()), float(o.min()), float(o.max())],
                                   gp[f"mri_{tag}_t1_sums"], rtol=1e-12)
        cfg = dict(pc.MRI_CFG, translate_zero=tz)
        assert list(lp.min_max_val(cfg, "mri")) == list(gp[f"minmax_{tag}"])
        assert list(producers.set_min_max_val(cfg, "mri")) == list(gp[f"minmax_{tag}"])   # host logic of the product
    assert producers.set_min_max_val({}, "mnist") == (2.0, 0.0)
    if True:  # translate_zero puts the minimum of every image at exactly 0
        o = lp.mri_normalize(t1.unsqueeze(1), pc.MRI_CFG["mean_t1"], pc.MRI_CFG["std_t1"], True, 224)
        assert float(o.flatten(1).min(dim=1).values.abs().max()) == 0.0


def test_oracle_masks(gp):
    for name, rule, cfg, amap, size, manual in pc.mask_cases():
        mp, bm = lp.masks_from_anomaly(amap, rule, size, 7 if manual else 0)
        assert np.array_equal(mp.numpy(), gp[f"mask_{name}_pred"]) and np.array_equal(bm.numpy(), gp[f"mask_{name}_bin"]), name


def test_oracle_knn(gp):
    for name, x, bank in pc.knn_cases():
        sc, loc = lp.knn_min(x, bank)
        assert np.array_equal(sc.numpy(), gp[f"knn_{name}_score"]) and np.array_equal(loc.numpy(), gp[f"knn_{name}_loc"])


def test_producers_refuse_to_run_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(Exception):
        producers.mnist_pair(pc.mnist_raw())


# ---- GPU: device kernels ----------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_mnist_pair_bit_exact(gp):
    hr, cond = producers.mnist_pair(pc.mnist_raw().to(DEV))
    assert np.array_equal(hr.cpu().numpy(), gp["mnist_hr"])
    assert np.array_equal(cond.cpu().numpy(), gp["mnist_cond"])     # uint8-valued inputs: every bilinear product is exact
    # arbitrary float inputs and an odd size: same formula, rounding order may differ from ATen's vectorised kernel
    g = torch.Generator().manual_seed(3)
    raw = 255 * torch.rand(3, 33, 33, generator=g)
    hr2, cond2 = producers.mnist_pair(raw.to(DEV))
    o_hr, o_cond = lp.mnist_pair(raw)
    assert torch.equal(hr2.cpu(), o_hr) and util.max_abs(cond2, o_cond) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("tz", [True, False])
def test_gpu_mri_normalize_bit_exact(gp, tz):
    t1, fl = pc.mri_raw()
    tag = "tz" if tz else "raw"
    got = producers.mri_normalize(t1.to(DEV), pc.MRI_CFG["mean_t1"], pc.MRI_CFG["std_t1"], tz, 224)
    want = lp.mri_normalize(t1.unsqueeze(1), pc.MRI_CFG["mean_t1"], pc.MRI_CFG["std_t1"], tz, 224)
    assert torch.equal(got.cpu(), want)
    assert np.array_equal(got.cpu()[:, :, ::4, ::4].numpy(), gp[f"mri_{tag}_t1_sub4"])
    # ragged crop offsets (odd margins) and a batch of one
    g = torch.Generator().manual_seed(5)
    raw = 4096 * torch.rand(1, 37, 41, generator=g)
    assert torch.equal(producers.mri_normalize(raw.to(DEV), 100.0, 900.0, tz, 32).cpu(), lp.mri_normalize(raw.unsqueeze(1), 100.0, 900.0, tz, 32))


@pytest.mark.gpu
def test_gpu_masks_match_reference(gp):
    for name, rule, cfg, amap, size, manual in pc.mask_cases():
        mp, bm = producers.masks_from_anomaly(amap.to(DEV), rule, size, 7 if manual else 0)
        ref_mp, ref_bm = torch.from_numpy(gp[f"mask_{name}_pred"]), torch.from_numpy(gp[f"mask_{name}_bin"])
        mp, bm = mp.cpu(), bm.cpu()
        assert torch.equal(bm, ref_bm), name                                     # binary mask: bit-exact
        assert torch.equal(mp == 1.0, ref_mp == 1.0), name                       # the `== 1.0` (OOD) region: bit-exact
        assert util.max_abs(mp, ref_mp) < 2e-5, name                             # soft region: std / resize rounding order


@pytest.mark.gpu
def test_gpu_mask_feeds_the_sampler_contract():
    """The produced mask is exactly what `sample()` partitions on: `mask >= 1` equals the binary mask (ddpm.py:672)."""
    name, rule, cfg, amap, size, manual = [c for c in pc.mask_cases() if c[0] == "mri_t12flair_3"][0]
    mp, bm = producers.masks_from_anomaly(amap.repeat(1, 1, 1, 1).to(DEV), rule, size)
    assert torch.equal((mp >= 1.0).float(), bm) and 0 < float(bm.mean()) < 1


@pytest.mark.gpu
def test_gpu_knn_matches_reference(gp):
    """PatchCore nearest neighbour (models.py:179-217) on tcgen05 with split-bf16 operands: fp32-level distances, same locations."""
    for name, x, bank in pc.knn_cases():
        sc, loc = producers.knn_min(x.to(DEV), bank.to(DEV))
        ref_sc, ref_loc = torch.from_numpy(gp[f"knn_{name}_score"]), torch.from_numpy(gp[f"knn_{name}_loc"])
        sc, loc = sc.cpu(), loc.cpu()
        assert util.max_abs(sc, ref_sc) < 0.1, name           # sqrt amplifies the rounding of d^2 near 0: the exact hit reads 0.05, not 0
        big = ref_sc > 0.1
        assert util.rel_err(sc[big], ref_sc[big]) < 1e-5, name
        # locations: identical wherever the runner-up is not within rounding distance of the minimum
        d = torch.cdist(x.double(), bank.double())
        top2 = d.topk(2, largest=False, dim=1).values
        clear = (top2[:, 1] - top2[:, 0]) > 1e-4
        assert torch.equal(loc[clear], ref_loc[clear]) and int(clear.sum()) > 0.9 * len(loc), name
    # PatchCore-sized problem: 784 patches of one image against a 16k-entry bank of 1536-d features; fp64 reference on the GPU
    g = torch.Generator().manual_seed(9)
    x, bank = torch.randn(784, 1536, generator=g).to(DEV), torch.randn(16385, 1536, generator=g).to(DEV)
    sc, loc = producers.knn_min(x, bank)
    d = torch.cdist(x.double(), bank.double())
    ref_sc, ref_loc = d.min(dim=1)
    assert util.rel_err(sc, ref_sc) < 1e-5
    top2 = d.topk(2, largest=False, dim=1).values
    clear = (top2[:, 1] - top2[:, 0]) > 1e-3
    assert torch.equal(loc[clear].cpu(), ref_loc[clear].cpu())
