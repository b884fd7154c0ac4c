"""CPU: the oracle restatement against the golden fixtures produced by the real reference."""
import numpy as np
import pytest
import torch

from oracle import ld_oracle as lo
from oracle import ref_harness as rh
from tests import util
from tests.golden import cases


def _sd(name):
    return util.cpu_state_dict(util.make_model(name))


def test_seeded_weights_match_reference_fingerprint(golden):
    for name in cases.MODEL_KW:
        got = np.array(cases.weight_checksum(_sd(name)))
        np.testing.assert_allclose(got, golden[f"wsum_{name}"], rtol=1e-12, atol=1e-9)


@pytest.mark.parametrize("sched,T", [("sigmoid", 1000), ("sigmoid", 50), ("linear", 100), ("cosine", 200)])
def test_schedule_buffers_bit_exact(golden, sched, T):
    b = lo.diffusion_buffers(sched, T)
    for k in ("betas", "posterior_mean_coef1", "posterior_mean_coef2", "posterior_log_variance_clipped",
              "sqrt_recip_alphas_cumprod", "loss_weight"):
        assert np.array_equal(b[k].numpy(), golden[f"sched_{sched}_{T}_{k}"]), (sched, T, k)


def test_schedule_check_values_from_survey():
    # SURVEY.md §8(a7) probe values of the reference's buffers
    b = lo.diffusion_buffers("sigmoid", 1000)
    assert abs(float(b["betas"][0]) - 3.0027920729e-4) < 1e-10 and float(b["betas"][-1]) == pytest.approx(0.999)
    assert float(b["posterior_mean_coef1"][0]) == 1.0
    assert float(b["posterior_mean_coef1"][-1]) == pytest.approx(0.0173114669, rel=1e-6)
    assert float(b["posterior_mean_coef2"][-1]) == pytest.approx(0.0316132903, rel=1e-6)
    assert float(b["posterior_log_variance_clipped"][0]) == pytest.approx(-46.0517006, rel=1e-6)
    assert float(b["posterior_log_variance_clipped"][1]) == pytest.approx(-8.80093479, rel=1e-6)
    b = lo.diffusion_buffers("linear", 100)
    assert float(b["betas"][1]) == pytest.approx(3.0101009e-3, rel=1e-6)
    assert float(b["posterior_mean_coef2"][-1]) == pytest.approx(0.89442265, rel=1e-6)


@pytest.mark.parametrize("name,S,B,ts", [("mnist", 32, 2, [99, 3]), ("mri", 64, 2, [999, 17]), ("mri_attn8", 64, 1, [500])])
def test_unet_forward_matches_reference_output(golden, name, S, B, ts):
    x = cases.noise_tape(B, S, 1)[0]
    cond = cases.cond_uniform(B, S)
    with torch.no_grad():
        y = lo.unet_forward(_sd(name), util.hp_of(name), x, cond, torch.tensor(ts))
    assert util.max_abs(y, torch.from_numpy(golden[f"unet_{name}_out"])) < 2e-5


SAMPLER_CASES = {
    # key: (model, data, S, T, s, B, cfg overrides, schedule, mask kind)
    "c1mri": ("mnist", "mri", 32, 24, 5, 4, {}, "sigmoid", "cols"),
    "c1s0": ("mnist", "mri", 32, 40, 0, 2, {}, "linear", "cols"),
    "c1pair": ("mnist", "mri", 32, 8, 2, 2, dict(start_intermediate=False), "sigmoid", "cols"),
    "c1ones": ("mnist", "mri", 32, 8, 2, 2, {}, "sigmoid", "ones"),
    "c1nobranch": ("mnist", "mri", 32, 8, 2, 2, dict(branch_out=False), "sigmoid", "cols"),
    "c2s": ("mri", "mri", 64, 12, 3, 2, {}, "sigmoid", "mri"),
    "c1": ("mnist", "mnist", 32, 100, 2, 8, {}, "sigmoid", "cols"),
}


def sampler_inputs(key):
    name, data, S, T, s, B, over, sched, mk = SAMPLER_CASES[key]
    if mk == "mri":
        cond, mask = cases.mri_like(B, S)
        mm = cases.MRI_MIN_MAX
    else:
        cond = cases.cond_uniform(8, S)[:B]
        mask = torch.ones(B, 1, S, S) if mk == "ones" else cases.mask_left_columns(B, S)
        mm = cases.MNIST_MIN_MAX
    return name, data, S, T, s, B, over, sched, cond, mask, mm


@pytest.mark.parametrize("key", ["c1mri", "c1s0", "c1pair", "c1ones", "c1nobranch", "c2s"])
def test_sampler_matches_reference_output(golden, key):
    name, data, S, T, s, B, over, sched, cond, mask, mm = sampler_inputs(key)
    cfg = cases.base_config(data, s, **over)
    smp = lo.Sampler(cfg, _sd(name), util.hp_of(name), image_size=S, timesteps=T, beta_schedule=sched, trace=[])
    with torch.no_grad():
        out = smp.sample(cond, mask, mm, list(cases.noise_tape(B, S, T)))
    assert util.max_abs(out, torch.from_numpy(golden[f"{key}_out"])) < 5e-5
    assert repr(cfg) == str(golden[f"{key}_cfg_after"])
    assert smp.unet_calls == int(golden[f"{key}_unet_calls"])
    x0_last = smp.trace[-1][2]
    x0_last = torch.stack(x0_last) if isinstance(x0_last, (tuple, list)) else x0_last
    assert util.max_abs(x0_last, torch.from_numpy(golden[f"{key}_x0_last"])) < 5e-5


def test_unet_call_count_formula():
    # 2(T-s)+s forwards, checked with a stub denoiser
    for T, s in ((10, 2), (10, 0), (7, 6)):
        cfg = cases.base_config("mri", s)
        smp = lo.Sampler(cfg, {}, util.hp_of("mnist"), image_size=8, timesteps=T, model_fn=lambda x, c, t: 0.5 * x + c)
        smp.sample(cases.cond_uniform(1, 8), cases.mask_left_columns(1, 8, 2), (0.0, 2.0), list(cases.noise_tape(1, 8, T)))
        assert smp.unet_calls == 2 * (T - s) + s  # steps T-1..s run both branches (the one at t == s fuses), then s single steps


def test_branch_mode_rejects_other_objectives():
    cfg = cases.base_config("mri", 2)
    smp = lo.Sampler(cfg, {}, util.hp_of("mnist"), image_size=8, timesteps=4, objective="pred_noise", model_fn=lambda x, c, t: x)
    with pytest.raises(UnboundLocalError):
        smp.sample(cases.cond_uniform(1, 8), cases.mask_left_columns(1, 8, 2), (0.0, 2.0), list(cases.noise_tape(1, 8, 4)))


def test_non_binary_mask_asserts():
    cfg = cases.base_config("mri", 2)
    smp = lo.Sampler(cfg, {}, util.hp_of("mnist"), image_size=8, timesteps=4, model_fn=lambda x, c, t: x)
    with pytest.raises(AssertionError):
        smp.sample(cases.cond_uniform(1, 8), 0.3 * torch.rand(1, 1, 8, 8), (0.0, 2.0), list(cases.noise_tape(1, 8, 4)))


@pytest.mark.skipif(not rh.available(), reason="reference tree only exists in the build container")
def test_oracle_against_live_reference(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    (tmp_path / "fusion_test").mkdir()
    ddpm = rh.load_reference()
    torch.manual_seed(0)
    m = ddpm.Unet(**cases.MODEL_KW["mnist"]).eval()
    B, S, T = 2, 32, 6
    cond, mask, tape = cases.cond_uniform(B, S), cases.mask_left_columns(B, S), cases.noise_tape(B, S, T)
    cfg = cases.base_config("mri", 2)
    gd = ddpm.GaussianDiffusion(cfg, m, image_size=S, timesteps=T, objective="pred_x0", auto_normalize=False).eval()
    with rh.noise_tape(list(tape)):
        ref = gd.sample(cond, None, batch_size=B, mask=mask, min_max_val=(0.0, 2.0))
    cfg2 = cases.base_config("mri", 2)
    smp = lo.Sampler(cfg2, {k: v.detach() for k, v in m.state_dict().items()}, util.hp_of("mnist"), image_size=S, timesteps=T)
    with torch.no_grad():
        out = smp.sample(cond, mask, (0.0, 2.0), list(tape))
    assert util.max_abs(out, ref) < 5e-5 and cfg == cfg2
