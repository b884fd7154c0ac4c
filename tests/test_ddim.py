"""DDIM branch sampler (`ddim_sample`, ddpm.py:979-1075): oracle vs the reference's golden outputs (CPU) and the
CUDA path through the C ABI vs the same vectors (GPU)."""
import os

import numpy as np
import pytest
import torch

from localdiffusion_hallucination_b200 import GaussianDiffusion, Unet
from oracle import ld_oracle as lo
from tests import util
from tests.golden import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = list(cases.DDIM_CASES)


@pytest.fixture(scope="module")
def gd():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_ddim.npz"), allow_pickle=False)


def inputs(key):
    c = cases.DDIM_CASES[key]
    if key == "d2":
        cond, mask = cases.mri_like(c["B"], 64)
        return "mri", 64, cond, mask, cases.MRI_MIN_MAX
    S, B = 32, c["B"]
    cond = cases.cond_uniform(8, S)[:B]
    mask = torch.ones(B, 1, S, S) if key == "d1ones" else cases.mask_left_columns(B, S)
    return "mnist", S, cond, mask, cases.MNIST_MIN_MAX


@pytest.mark.parametrize("key", KEYS)
def test_oracle_ddim_matches_reference_output(gd, key):
    c = cases.DDIM_CASES[key]
    name, S, cond, mask, mm = inputs(key)
    cfg = cases.base_config(c["data"], c["s"], **c.get("cfg", {}))
    torch.manual_seed(0)
    sd = {k: v.detach().clone() for k, v in Unet(**cases.MODEL_KW[name]).state_dict().items()}
    smp = lo.Sampler(cfg, sd, util.hp_of(name), image_size=S, timesteps=c["T"], beta_schedule=c["sched"])
    with torch.no_grad():
        out = lo.ddim_sample(smp, cond, mask, mm, list(cases.noise_tape(c["B"], S, c["steps"])), c["steps"], c["eta"])
    assert isinstance(out, list) == bool(int(gd[f"{key}_pair"]))
    out = torch.stack(out) if isinstance(out, list) else out
    assert util.max_abs(out, torch.from_numpy(gd[f"{key}_out"])) < 5e-5
    assert repr(cfg) == str(gd[f"{key}_cfg_after"])
    assert smp.unet_calls == int(gd[f"{key}_unet_calls"])


def test_ddim_schedule_is_pinned(gd):
    """Host-side (time, coefficient) table: times of ddpm.py:984-986, fusion step of ddpm.py:987,1022."""
    c = cases.DDIM_CASES["d1eta"]
    g = GaussianDiffusion(cases.base_config(c["data"], c["s"]), Unet(**cases.MODEL_KW["mnist"]), image_size=32, timesteps=c["T"],
                          sampling_timesteps=c["steps"], beta_schedule=c["sched"], objective="pred_x0", ddim_sampling_eta=c["eta"])
    times, coefs, fuse = g.ddim_schedule()
    assert times == gd["d1eta_times"].tolist() and times[0] == c["T"] - 1 and len(times) == c["steps"]
    assert np.array_equal(coefs.numpy(), gd["d1eta_coefs"])
    assert fuse == int(gd["d1eta_fuse"]) == c["steps"] - 1 - c["s"]
    assert float(coefs[:-1, 4].min()) > 0.0  # eta > 0: every step with a successor is stochastic


def run_gpu(key, precision):
    c = cases.DDIM_CASES[key]
    name, S, cond, mask, mm = inputs(key)
    m = util.make_model(name, precision, device="cuda:0")
    cfg = cases.base_config(c["data"], c["s"], **c.get("cfg", {}))
    g = GaussianDiffusion(cfg, m, image_size=S, timesteps=c["T"], sampling_timesteps=c["steps"], beta_schedule=c["sched"],
                          objective="pred_x0", ddim_sampling_eta=c["eta"]).to("cuda:0")
    out = g.sample(cond, None, batch_size=c["B"], mask=mask, min_max_val=mm, noise=cases.noise_tape(c["B"], S, c["steps"]))
    return out, cfg, mm


@pytest.mark.gpu
@pytest.mark.parametrize("key", KEYS)
def test_gpu_ddim_fp32_matches_reference_golden(gd, key):
    out, cfg, mm = run_gpu(key, "fp32")
    assert isinstance(out, list) == bool(int(gd[f"{key}_pair"]))  # never fused: the reference returns [x_out, x_in]
    out = torch.stack(out) if isinstance(out, list) else out
    ref = torch.from_numpy(gd[f"{key}_out"])
    assert tuple(out.shape) == tuple(ref.shape)
    assert util.max_abs(out, ref) < 2e-3 * mm[1]
    assert util.psnr(out, ref, mm[1]) > 60.0
    assert repr(cfg) == str(gd[f"{key}_cfg_after"])


@pytest.mark.gpu
@pytest.mark.parametrize("key", ["d1", "d1eta", "d2"])
def test_gpu_ddim_bf16_psnr(gd, key):
    out, cfg, mm = run_gpu(key, "bf16")
    out = torch.stack(out) if isinstance(out, list) else out
    assert util.psnr(out, torch.from_numpy(gd[f"{key}_out"]), mm[1]) > 40.0
