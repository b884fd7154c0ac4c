"""GPU: engine robustness -- checkpoint ingestion end to end, engine invalidation, LinearAttention soft-max shift
recovery, deferred checks, plan re-use across branch-mode changes, the vectorised step pass at bench size."""
import ctypes as C

import pytest
import torch

from localdiffusion_hallucination_b200 import GaussianDiffusion, _lib, load_reference_checkpoint
from oracle import ld_oracle as lo
from tests import util
from tests.golden import cases

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MM = cases.MNIST_MIN_MAX


def _gd(seed, prec="fp32", T=8, s=2, **opts):
    m = util.make_model("mnist", prec, seed=seed, device=DEV, **opts)
    return GaussianDiffusion(cases.base_config("mri", s), m, image_size=32, timesteps=T, objective="pred_x0").to(DEV)


def _reference_style_ckpt(gd, step=7):
    """`Trainer.save` layout (ddpm.py:1495-1507); the EMA copy's Unet tensors differ from the online weights (schedule buffers do not)."""
    sd = {k: v.detach().cpu().clone() for k, v in gd.state_dict().items()}
    bump = lambda k, v: v * 0.9 if (k.startswith("model.") and v.is_floating_point()) else v.clone()
    ema = {"initted": torch.tensor(True), "step": torch.tensor(step)}
    ema.update({"ema_model." + k: bump(k, v) for k, v in sd.items()})
    ema.update({"online_model." + k: v.clone() for k, v in sd.items()})
    return {"step": step, "model": sd, "opt": {}, "ema": ema, "scaler": None}


def _inputs(B=2, T=8):
    return cases.cond_uniform(B, 32), cases.mask_left_columns(B, 32), cases.noise_tape(B, 32, T)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_checkpoint_ingest_then_sample(tmp_path, prec):
    """f3: write a checkpoint in the reference's `Trainer.save` layout, ingest the EMA weights into a differently initialised model
    (whose engine already exists), sample, and compare with a model that was constructed with those weights."""
    cond, mask, tape = _inputs()
    src = _gd(1, prec)
    ck = _reference_style_ckpt(src)
    path = tmp_path / "model-best1.pt"
    torch.save(ck, path)
    want_sd = {k[len("ema_model."):]: v for k, v in ck["ema"].items() if k.startswith("ema_model.")}
    ref = _gd(5, prec)
    ref.load_state_dict(want_sd)
    want = ref.sample(cond, None, batch_size=2, mask=mask, min_max_val=MM, noise=tape)
    dst = _gd(2, prec)
    before = dst.sample(cond, None, batch_size=2, mask=mask, min_max_val=MM, noise=tape)   # engine built with the OLD weights
    assert load_reference_checkpoint(dst, str(path)) == 7
    got = dst.sample(cond, None, batch_size=2, mask=mask, min_max_val=MM, noise=tape)
    if prec == "fp32":
        assert util.max_abs(got, want) < 1e-4                              # atomics order only
    else:
        assert util.psnr(got, want, MM[1]) > 45.0                          # ... amplified by bf16 rounding of the activations
    assert util.max_abs(got, before) > 1e-3                                # and it really changed
    # oracle on the ingested weights
    sd = {k[len("model."):]: v.cpu() for k, v in dst.state_dict().items() if k.startswith("model.")}
    smp = lo.Sampler(cases.base_config("mri", 2), sd, util.hp_of("mnist"), image_size=32, timesteps=8)
    with torch.no_grad():
        o = smp.sample(cond, mask, MM, list(tape))
    if prec == "fp32":
        assert util.max_abs(got, o) < 2e-3 * MM[1]
    else:
        assert util.psnr(got, o, MM[1]) > 40.0


def test_parent_load_and_inplace_updates_reach_the_engine():
    cond, mask, tape = _inputs()
    a, b = _gd(1), _gd(2)
    want = a.sample(cond, None, batch_size=2, mask=mask, min_max_val=MM, noise=tape)
    b.sample(cond, None, batch_size=2, mask=mask, min_max_val=MM, noise=tape)
    b.load_state_dict(a.state_dict())                                      # the reference's own flow (ddpm.py:1517): parent load
    got = b.sample(cond, None, batch_size=2, mask=mask, min_max_val=MM, noise=tape)
    assert util.max_abs(got, want) < 1e-4
    with torch.no_grad():                                                  # EMA-style in-place update
        for p in b.model.parameters():
            p.mul_(0.5)
    got2 = b.sample(cond, None, batch_size=2, mask=mask, min_max_val=MM, noise=tape)
    assert util.max_abs(got2, want) > 1e-3


def _get_option(h, name):
    v = C.c_int64(-1)
    _lib.check(_lib.lib().ld_get_option(h, name.encode(), C.byref(v)))
    return int(v.value)


def test_linattn_shift_underflow_recovers_by_itself():
    """VERDICT r1 weak 4 / ADVICE: with extreme to_qkv weights the analytic soft-max shift of the fused LinearAttention underflows
    (every weight of a (head, d) row flushes to zero).  The engine has to notice after the first timestep, switch to the exact-max
    kernels for good and repeat the call: no error, and from then on the same arithmetic as option la_exact=1.  (At such weight
    scales |k| is in the hundreds, far beyond what bf16 activations resolve inside an exponential, so agreement with the fp32 oracle is
    not the point here -- in the regime where bf16 is meaningful the bound cannot underflow: it needs bound - max(k) > 87.)"""
    cond, mask, tape = _inputs(T=12)
    outs = {}
    for name, opts in (("auto", {}), ("exact", dict(la_exact=1))):
        gd = _gd(0, "bf16", T=12, **opts)
        with torch.no_grad():
            for k, p in gd.model.named_parameters():
                if k.endswith("to_qkv.weight") and p.shape[0] == 384:      # LinearAttention blocks (4 heads x 32 x 3)
                    p[128:256].zero_()
                    p[128:256, 0].fill_(60.0)                              # k depends on ONE channel of xhat: |k| <= bound * |xhat_0| << bound
        h = gd.model.engine()
        if name == "auto":
            assert _get_option(h, "la_exact") == 0
        outs[name] = gd.sample(cond, None, batch_size=2, mask=mask, min_max_val=MM, noise=tape)
        assert bool(torch.isfinite(outs[name]).all())
        assert _get_option(gd.model.engine(), "la_exact") == 1             # "auto": switched by the engine itself
    # same kernels after the switch; with soft-max weights this peaked a last-bit difference (atomics order) can move the winning pixel,
    # so the two runs are compared statistically
    assert util.psnr(outs["auto"], outs["exact"], MM[1]) > 25.0


def test_async_option_defers_the_asserts():
    lib = _lib.lib()
    cond, mask, tape = _inputs()
    gd = _gd(0, "fp32")
    want = gd.sample(cond, None, batch_size=2, mask=mask, min_max_val=MM, noise=tape)
    h = gd.model.engine()
    _lib.check(lib.ld_set_option(h, b"async", 1))
    got = gd.sample(cond, None, batch_size=2, mask=mask, min_max_val=MM, noise=tape)
    _lib.check(lib.ld_sample_finish(h))
    assert util.max_abs(got, want) < 1e-4
    # a non-binary mask is reported by ld_sample_finish, not by the call
    gd.config.update(cases.base_config("mri", 2))
    bad = 0.3 * torch.rand(2, 1, 32, 32)
    gd.sample(cond, None, batch_size=2, mask=bad, min_max_val=MM, noise=tape)
    with pytest.raises(AssertionError):
        _lib.check(lib.ld_sample_finish(h))
    _lib.check(lib.ld_set_option(h, b"async", 0))


def test_plans_survive_branch_mode_changes():
    """The reference's test loop alternates anomalous masks (branched) with all-ones masks (vanilla DDPM fallback, ddpm.py:1110-1117)."""
    lib = _lib.lib()
    cond, mask, tape = _inputs()
    ones = torch.ones_like(mask)
    gd = _gd(0)
    a1 = gd.sample(cond, None, batch_size=2, mask=mask, min_max_val=MM, noise=tape)
    h = gd.model.engine()
    b1 = gd.sample(cond, None, batch_size=2, mask=ones, min_max_val=MM, noise=tape)
    ws = lib.ld_workspace_bytes(h)
    a2 = gd.sample(cond, None, batch_size=2, mask=mask, min_max_val=MM, noise=tape)
    b2 = gd.sample(cond, None, batch_size=2, mask=ones, min_max_val=MM, noise=tape)
    assert lib.ld_workspace_bytes(h) == ws and ws > 0                      # nothing was rebuilt
    assert util.max_abs(a1, a2) < 1e-4 and util.max_abs(b1, b2) < 1e-4 and util.max_abs(a1, b1) > 1e-3


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_vectorised_step_is_bit_exact_at_bench_size(kind):
    """The fused posterior / clamp / composite pass at BASELINE configs[1] size (B=16, 256x256): bit-exact against plain torch
    fp32 ops in the reference's evaluation order (ddpm.py:659-666, 697-708, 775-810, 852-858)."""
    lib = _lib.lib()
    T, t, B, S = 50, 7, 16, 256
    n = B * S * S
    gd = _gd(0, T=T)
    h = gd.model.engine()
    gd._push_schedule(h)
    g = torch.Generator().manual_seed(3)
    o_out, o_in, x_out, x_in, z = (3 * torch.randn(n, generator=g) for _ in range(5))
    cond = 2 * torch.rand(n, generator=g)
    mask = (torch.rand(n, generator=g) > 0.6).float()
    mask[::7] = 0.4
    lo_, hi_ = 0.0, 2.0
    c1, c2 = gd.posterior_mean_coef1[t].cpu(), gd.posterior_mean_coef2[t].cpu()
    sg = (0.5 * gd.posterior_log_variance_clipped[t].cpu()).exp()
    bm = (mask >= 1.0).float()
    if kind == 2:
        x0 = o_out.clamp(lo_, hi_)
        want = (c1 * x0 + c2 * x_out + sg * z,)
    else:
        x0o = torch.where(bm == 0.0, torch.tensor(lo_), o_out * bm).clamp(lo_, hi_)
        x0i = o_in.clamp(lo_, hi_)
        if kind == 0:
            want = (c1 * x0o + c2 * x_out + sg * z, c1 * x0i + c2 * x_in + sg * z)
        else:
            x0 = (x0i * (1.0 - bm) + x0o).clamp(lo_, hi_)
            xo, xi = x_out * bm, x_in * (1.0 - bm)
            xt = torch.where(xo == 0.0, xi, xo)
            want = (c1 * x0 + c2 * xt + sg * z,)
    sd = _lib.SampleDesc()
    sd.mask_x, sd.ood_uses_cond, sd.cond_in_floor, sd.min_val, sd.max_val = 1, 0, 0.95, lo_, hi_
    d = [v.clone().to(DEV).contiguous() for v in (x_out, x_in, o_out, o_in, cond, mask, z)]
    _lib.check(lib.ld_posterior_step(h, kind, t, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), d[4].data_ptr(),
                                     d[5].data_ptr(), d[6].data_ptr(), C.byref(sd), n, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert torch.equal(d[0].cpu(), want[0])
    if kind == 0:
        assert torch.equal(d[1].cpu(), want[1])
