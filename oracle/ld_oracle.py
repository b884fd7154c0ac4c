"""CPU oracle for the LocalDiffusion conditional reverse-diffusion sampler.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import this file; the product package
(`localdiffusion_hallucination_b200`) never does and has no CPU fallback.

This is a *restatement* (plain PyTorch fp32 CPU ops, functional style, driven by a flat
state_dict) of the algorithm on the reference's hot path.  It is not a copy of the reference
classes: there are no nn.Modules here, weights are looked up by their state_dict key.
Every function cites the reference lines it follows (paths relative to the reference root).

Parity pinning: the reference ships no tests / golden vectors for this path, so
the oracle is pinned against outputs of the reference itself, run in the build container
through `oracle/ref_harness.py`; the generating script is `tests/golden/make_golden.py` and the
resulting fixtures live in `tests/golden/*.npz`.  `tests/test_oracle_golden.py` re-checks the
oracle against those fixtures on every run (and against the live reference when it is present).

Floating point: everything is fp32 like the reference (`ddpm.py:567`, `amp=False`), schedules are
derived in fp64 and cast (`ddpm.py:547-593`).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# hyper-parameters of the denoiser (ddpm.py:287-307)
# ----------------------------------------------------------------------------------------------
@dataclass
class UnetHP:
    dim: int = 32
    init_dim: Optional[int] = None
    dim_mults: Sequence[int] = (1, 2, 4, 8)
    channels: int = 1
    groups: int = 8
    theta: float = 10000.0
    dim_head: int = 32
    heads: int = 4
    full_attn: Sequence[bool] = (False, False, False, True)
    mode: str = "mri"

    def dims(self):
        i = self.init_dim if self.init_dim is not None else self.dim
        return [i] + [self.dim * m for m in self.dim_mults]


# ----------------------------------------------------------------------------------------------
# small blocks
# ----------------------------------------------------------------------------------------------
def rms_norm(x: Tensor, g: Tensor) -> Tensor:
    """ddpm.py:131-132 -- x / max(||x||_2 over C, 1e-12) * g * sqrt(C)."""
    n = x.pow(2).sum(dim=1, keepdim=True).sqrt().clamp_min(1e-12)
    return x / n * g * (x.shape[1] ** 0.5)


def time_embedding(sd: Dict[str, Tensor], hp: UnetHP, t: Tensor) -> Tensor:
    """ddpm.py:142-149 + 339-344 -- sin/cos(dim) -> Linear -> GELU(erf) -> Linear."""
    half = hp.dim // 2
    step = math.log(hp.theta) / (half - 1)
    freq = torch.exp(torch.arange(half) * -step)
    arg = t[:, None] * freq[None, :]
    e = torch.cat((arg.sin(), arg.cos()), dim=-1)
    e = F.linear(e, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"])
    e = F.gelu(e)
    return F.linear(e, sd["time_mlp.3.weight"], sd["time_mlp.3.bias"])


def conv_gn_act(sd, pfx: str, x: Tensor, groups: int, film=None) -> Tensor:
    """ddpm.py:177-186 (`Block`): conv3x3 -> GroupNorm -> x*(scale+1)+shift -> SiLU."""
    y = F.conv2d(x, sd[pfx + ".proj.weight"], sd[pfx + ".proj.bias"], padding=1)
    y = F.group_norm(y, groups, sd[pfx + ".norm.weight"], sd[pfx + ".norm.bias"], eps=1e-5)
    if film is not None:
        sc, sh = film
        y = y * (sc + 1) + sh
    return F.silu(y)


def resnet_block(sd, pfx: str, x: Tensor, temb: Optional[Tensor], groups: int) -> Tensor:
    """ddpm.py:200-212 (`ResnetBlock.forward`)."""
    film = None
    if temb is not None:
        ss = F.linear(F.silu(temb), sd[pfx + ".mlp.1.weight"], sd[pfx + ".mlp.1.bias"])
        ss = ss[:, :, None, None]
        film = ss.chunk(2, dim=1)
    h = conv_gn_act(sd, pfx + ".block1", x, groups, film)
    h = conv_gn_act(sd, pfx + ".block2", h, groups)
    if (pfx + ".res_conv.weight") in sd:
        x = F.conv2d(x, sd[pfx + ".res_conv.weight"], sd[pfx + ".res_conv.bias"])
    return h + x


def linear_attention(sd, pfx: str, x: Tensor, heads: int, dim_head: int) -> Tensor:
    """ddpm.py:234-251 (`LinearAttention.forward`)."""
    b, c, hh, ww = x.shape
    xn = rms_norm(x, sd[pfx + ".norm.g"])
    qkv = F.conv2d(xn, sd[pfx + ".to_qkv.weight"])
    q, k, v = [z.reshape(b, heads, dim_head, hh * ww) for z in qkv.chunk(3, dim=1)]
    q = q.softmax(dim=-2) * dim_head ** -0.5
    k = k.softmax(dim=-1)
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(b, heads * dim_head, hh, ww)
    out = F.conv2d(out, sd[pfx + ".to_out.0.weight"], sd[pfx + ".to_out.0.bias"])
    return rms_norm(out, sd[pfx + ".to_out.1.g"])


def full_attention(sd, pfx: str, x: Tensor, heads: int, dim_head: int) -> Tensor:
    """ddpm.py:271-282 (`Attention.forward`) + attend.py:98-113 (math path, dropout p=0)."""
    b, c, hh, ww = x.shape
    xn = rms_norm(x, sd[pfx + ".norm.g"])
    qkv = F.conv2d(xn, sd[pfx + ".to_qkv.weight"])
    q, k, v = [z.reshape(b, heads, dim_head, hh * ww).transpose(-1, -2) for z in qkv.chunk(3, dim=1)]
    sim = torch.einsum("bhid,bhjd->bhij", q, k) * dim_head ** -0.5
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bhij,bhjd->bhid", attn, v)
    out = out.transpose(-1, -2).reshape(b, heads * dim_head, hh, ww)
    return F.conv2d(out, sd[pfx + ".to_out.weight"], sd[pfx + ".to_out.bias"])


def cond_basic_block(sd, pfx: str, x: Tensor) -> Tensor:
    """unet_model.py:37-51 (`BasicBlock.forward`), GroupNorm(16) (unet_model.py:6)."""
    a = F.conv2d(x, sd[pfx + ".convblock.0.weight"], sd[pfx + ".convblock.0.bias"], padding=1)
    a = F.relu(F.group_norm(a, 16, sd[pfx + ".convblock.1.weight"], sd[pfx + ".convblock.1.bias"], eps=1e-5))
    a = F.conv2d(a, sd[pfx + ".convblock.3.weight"], sd[pfx + ".convblock.3.bias"], padding=1)
    a = F.group_norm(a, 16, sd[pfx + ".convblock.4.weight"], sd[pfx + ".convblock.4.bias"], eps=1e-5)
    if (pfx + ".identity.0.weight") in sd:
        i = F.conv2d(x, sd[pfx + ".identity.0.weight"], sd[pfx + ".identity.0.bias"], padding=1)
        i = F.group_norm(i, 16, sd[pfx + ".identity.1.weight"], sd[pfx + ".identity.1.bias"], eps=1e-5)
    else:
        i = x
    return F.relu(a + i)


def cond_encoder(sd, hp: UnetHP, cond: Tensor) -> Tensor:
    """unet_model.py:122-137 (`ResUnet.forward`); mnist/mvtecSR stop after block 3."""
    p = "cond_model."
    y = cond_basic_block(sd, p + "residual_conv1.0", cond.float())
    y = F.max_pool2d(y, 2)
    y = cond_basic_block(sd, p + "residual_conv2.0", y)
    y = F.max_pool2d(y, 2)
    y = cond_basic_block(sd, p + "residual_conv3.0", y)
    if hp.mode in ("mnist", "mvtecSR"):
        return y
    y = F.max_pool2d(y, 2)
    return cond_basic_block(sd, p + "mid_conv.0", y)


def _attn(sd, hp: UnetHP, pfx: str, x: Tensor, full: bool) -> Tensor:
    if full:
        return full_attention(sd, pfx, x, hp.heads, hp.dim_head)
    return linear_attention(sd, pfx, x, hp.heads, hp.dim_head)


def unet_forward(sd: Dict[str, Tensor], hp: UnetHP, x: Tensor, cond: Tensor, t: Tensor, taps: Optional[dict] = None) -> Tensor:
    """ddpm.py:404-451 (`Unet.forward`), self_condition off.

    `sd` is `Unet.state_dict()` (keys without the `model.` prefix).  `t` is int64 `[B]`.
    `taps`, when given, receives named intermediate activations (for layer-wise parity tests).
    """
    def tap(name, v):
        if taps is not None:
            taps[name] = v
        return v

    L = len(hp.dim_mults)
    assert x.shape[-1] % (2 ** (L - 1)) == 0 and x.shape[-2] % (2 ** (L - 1)) == 0  # ddpm.py:405
    g = hp.groups
    x = tap("init_conv", F.conv2d(x, sd["init_conv.weight"], sd["init_conv.bias"], padding=3))
    r = x
    temb = time_embedding(sd, hp, t.float())
    skips: List[Tensor] = []
    for i in range(L):
        p = "downs.%d." % i
        x = tap(p + "0", resnet_block(sd, p + "0", x, temb, g))
        skips.append(x)
        x = tap(p + "1", resnet_block(sd, p + "1", x, temb, g))
        x = tap(p + "2", _attn(sd, hp, p + "2", x, hp.full_attn[i]) + x)
        skips.append(x)
        if i < L - 1:  # ddpm.py:120-124: pixel-unshuffle (c p1 p2) then 1x1
            x = F.conv2d(F.pixel_unshuffle(x, 2), sd[p + "3.1.weight"], sd[p + "3.1.bias"])
        else:  # ddpm.py:372
            x = F.conv2d(x, sd[p + "3.weight"], sd[p + "3.bias"], padding=1)
        tap(p + "3", x)
    x = tap("mid_block1", resnet_block(sd, "mid_block1", x, temb, g))
    x = tap("mid_attn", full_attention(sd, "mid_attn", x, hp.heads, hp.dim_head) + x)
    x = tap("mid_block2", resnet_block(sd, "mid_block2", x, temb, g))
    x = torch.cat((x, cond_encoder(sd, hp, cond)), dim=1)  # ddpm.py:434-435
    x = tap("conv_fusion", resnet_block(sd, "conv_fusion", x, None, g))  # ddpm.py:436 -- called WITHOUT t
    for i in range(L):
        p = "ups.%d." % i
        x = tap(p + "0", resnet_block(sd, p + "0", torch.cat((x, skips.pop()), dim=1), temb, g))
        x = tap(p + "1", resnet_block(sd, p + "1", torch.cat((x, skips.pop()), dim=1), temb, g))
        x = tap(p + "2", _attn(sd, hp, p + "2", x, hp.full_attn[L - 1 - i]) + x)
        if i < L - 1:  # ddpm.py:114-118: nearest x2 then 3x3
            x = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), sd[p + "3.1.weight"], sd[p + "3.1.bias"], padding=1)
        else:  # ddpm.py:391
            x = F.conv2d(x, sd[p + "3.weight"], sd[p + "3.bias"], padding=1)
        tap(p + "3", x)
    x = tap("final_res_block", resnet_block(sd, "final_res_block", torch.cat((x, r), dim=1), temb, g))
    return F.conv2d(x, sd["final_conv.weight"], sd["final_conv.bias"])


# ----------------------------------------------------------------------------------------------
# schedules (ddpm.py:460-494, 547-593)
# ----------------------------------------------------------------------------------------------
def beta_schedule(name: str, T: int) -> Tensor:
    if name == "linear":  # ddpm.py:460-467
        s = 1000 / T
        return torch.linspace(s * 0.0001, s * 0.02, T, dtype=torch.float64)
    u = torch.linspace(0, T, T + 1, dtype=torch.float64) / T
    if name == "cosine":  # ddpm.py:469-479
        ac = torch.cos((u + 0.008) / 1.008 * math.pi * 0.5) ** 2
    elif name == "sigmoid":  # ddpm.py:481-494 (start=-3, end=3, tau=1)
        v0, v1 = torch.tensor(-3.0).sigmoid(), torch.tensor(3.0).sigmoid()
        ac = (-((u * 6 - 3)).sigmoid() + v1) / (v1 - v0)
    else:
        raise ValueError(f"unknown beta schedule {name}")
    ac = ac / ac[0]
    return torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)


def diffusion_buffers(name: str, T: int, objective: str = "pred_x0") -> Dict[str, Tensor]:
    """The 13 fp32 `[T]` buffers of `GaussianDiffusion` (ddpm.py:547-615)."""
    b = beta_schedule(name, T)
    a = 1.0 - b
    ac = torch.cumprod(a, dim=0)
    acp = F.pad(ac[:-1], (1, 0), value=1.0)
    pv = b * (1.0 - acp) / (1.0 - ac)
    snr = ac / (1 - ac)
    lw = {"pred_noise": snr / snr, "pred_x0": snr, "pred_v": snr / (snr + 1)}[objective]
    d = dict(
        betas=b, alphas_cumprod=ac, alphas_cumprod_prev=acp,
        sqrt_alphas_cumprod=ac.sqrt(), sqrt_one_minus_alphas_cumprod=(1 - ac).sqrt(),
        log_one_minus_alphas_cumprod=(1 - ac).log(), sqrt_recip_alphas_cumprod=(1 / ac).sqrt(),
        sqrt_recipm1_alphas_cumprod=(1 / ac - 1).sqrt(), posterior_variance=pv,
        posterior_log_variance_clipped=pv.clamp(min=1e-20).log(),
        posterior_mean_coef1=b * acp.sqrt() / (1 - ac),
        posterior_mean_coef2=(1 - acp) * a.sqrt() / (1 - ac), loss_weight=lw,
    )
    return {k: v.to(torch.float32) for k, v in d.items()}


# ----------------------------------------------------------------------------------------------
# the sampler (ddpm.py:668-977, 1078-1125), DDPM branch only
# ----------------------------------------------------------------------------------------------
_NON_MRI = ("mnist", "mvtec", "oct", "imagenet")


@dataclass
class Sampler:
    """Restatement of `GaussianDiffusion` restricted to DDPM sampling with objective pred_x0.

    `config` is shared by reference and mutated exactly like the reference does
    (ddpm.py:780-781, 1093-1117) so call-to-call behaviour can be compared.
    """
    config: dict
    sd: Dict[str, Tensor]  # Unet state_dict (no prefix)
    hp: UnetHP
    image_size: int
    timesteps: int = 1000
    beta_schedule: str = "sigmoid"
    objective: str = "pred_x0"
    model_fn: Optional[Callable] = None  # override of the denoiser (tests)
    trace: Optional[list] = None  # per-step records when a list is supplied
    unet_calls: int = 0
    buf: Dict[str, Tensor] = field(default_factory=dict)

    def __post_init__(self):
        self.ctor_branch_out = self.config["branch_out"]  # ddpm.py:519
        self.ctor_start_intermediate = self.config["start_intermediate"]  # ddpm.py:520
        self.buf = diffusion_buffers(self.beta_schedule, self.timesteps, self.objective)
        self.num_timesteps = self.timesteps

    # -- denoiser ---------------------------------------------------------------------------
    def _model(self, x, cond, t):
        self.unet_calls += 1
        if self.model_fn is not None:
            return self.model_fn(x, cond, t)
        return unet_forward(self.sd, self.hp, x, cond, t)

    # -- ddpm.py:631-653, 731-761: x0 from the model output for the three objectives (single trajectory) ------------------
    def _x0_from_output(self, x, o, t: int):
        if self.objective == "pred_noise":
            return self.buf["sqrt_recip_alphas_cumprod"][t] * x - self.buf["sqrt_recipm1_alphas_cumprod"][t] * o
        if self.objective == "pred_v":
            return self.buf["sqrt_alphas_cumprod"][t] * x - self.buf["sqrt_one_minus_alphas_cumprod"][t] * o
        return o

    # -- ddpm.py:659-666 -----------------------------------------------------------------------
    def _posterior(self, x0, xt, t: int):
        mean = self.buf["posterior_mean_coef1"][t] * x0 + self.buf["posterior_mean_coef2"][t] * xt
        return mean, self.buf["posterior_log_variance_clipped"][t]

    # -- ddpm.py:668-766 + 768-838 ---------------------------------------------------------------
    def _mean_variance(self, x, mask, mm, cond, t: int):
        cfg = self.config
        tt = torch.full((cond.shape[0],), t, dtype=torch.long)
        lo, hi = float(mm[0]), float(mm[1])
        if cfg["branch_out"]:
            if self.objective != "pred_x0":
                # ddpm.py:731-733,757-761: other objectives reference an unbound `model_output`
                raise UnboundLocalError("branch sampling requires objective='pred_x0'")
            bm = (mask >= 1.0).float()  # ddpm.py:672
            cond_out = cond * bm  # ddpm.py:677
            floor = 0.5 if cfg["data"] == "mnist" else 0.95  # ddpm.py:683-686
            cond_in = (cond * torch.clip(1.0 - bm, floor, 1.0)).float()
            o_out = self._model(x[0], cond_out, tt)  # ddpm.py:694
            o_in = self._model(x[1], cond_in, tt)  # ddpm.py:695
            if cfg["mask_x"]:  # ddpm.py:697-708
                assert len(torch.unique(bm)) == 2, "mask should be binary"
                o_out = torch.where(bm == 0.0, torch.tensor(lo), o_out * bm)
                d = cfg["data"]
                if any(s in d for s in _NON_MRI) and "mri" not in d:
                    o_out = cond_out
            x0_out = o_out.clamp(lo, hi)  # ddpm.py:775-776
            x0_in = o_in.clamp(lo, hi)
            if t <= cfg["start_timestep"] and cfg["start_intermediate"]:  # ddpm.py:779-810
                cfg["branch_out"] = False
                cfg["mask_x"] = False
                m = (mask >= 1.0).float()
                x0 = x0_in * (1.0 - m) + x0_out  # ddpm.py:785-786 (an add, not a select)
                xo = x[0] * m
                xi = x[1] * (1.0 - m)
                assert bool((xo == 0).any()) and bool((xi == 0).any()), "x_out and x_in should be masked"
                xt = torch.where(xo == 0.0, xi, xo)  # ddpm.py:797
                x0 = x0.clamp(lo, hi)
                mean, lv = self._posterior(x0, xt, t)
                return mean, lv, x0, xt
            m_o, lv = self._posterior(x0_out, x[0], t)
            m_i, _ = self._posterior(x0_in, x[1], t)
            return (m_o, m_i), lv, (x0_out, x0_in), None
        o = self._model(x, cond, tt)  # ddpm.py:716 -- full, unmasked condition
        x0 = self._x0_from_output(x, o, t).clamp(lo, hi)  # ddpm.py:731-761, 821
        mean, lv = self._posterior(x0, x, t)
        return mean, lv, x0, None

    # -- ddpm.py:841-860 -------------------------------------------------------------------------
    def _p_sample(self, x, mask, mm, cond, t: int, draw):
        mean, lv, x0, _ = self._mean_variance(x, mask, mm, cond, t)
        if self.config["branch_out"]:  # re-read after a possible flip
            z = draw() if t > 0 else 0.0
            s = (0.5 * lv).exp()
            return [mean[0] + s * z, mean[1] + s * z], x0
        z = draw() if t > 0 else 0.0
        return mean + (0.5 * lv).exp() * z, x0

    # -- ddpm.py:1078-1125 + 930-977 -------------------------------------------------------------
    def sample(self, cond: Tensor, mask: Tensor, min_max_val, noise_tape: Sequence[Tensor], gt: Optional[Tensor] = None):
        """`noise_tape[0]` is x_T, `noise_tape[1:]` the per-step draws for t = T-1 .. 1."""
        cfg = self.config
        if cfg["branch_out"] is False:
            cfg["branch_out"] = self.ctor_branch_out
        if cfg["start_intermediate"] is False:
            cfg["start_intermediate"] = self.ctor_start_intermediate
        self.ctor_start_intermediate = bool(cfg["start_intermediate"])  # ddpm.py:1099-1104
        if cfg["ood_AD"] or cfg["ood_confidence"]:
            cfg["mask_cond"] = True
            cfg["mask_x"] = True
        if cfg["branch_out"]:
            u = torch.unique(mask)
            if len(u) == 1 and float(u[0]) == 1.0:  # ddpm.py:1110-1117: vanilla DDPM
                cfg["mask_cond"] = cfg["mask_x"] = cfg["branch_out"] = cfg["start_intermediate"] = False
        it = iter(noise_tape)
        draw = lambda: next(it).clone()
        img = draw()
        if self.ctor_start_intermediate and cfg.get("use_gt", False):  # ddpm.py:937-944
            tg = int(cfg["use_gt_timestep"])
            img = self.buf["sqrt_alphas_cumprod"][tg] * gt + self.buf["sqrt_one_minus_alphas_cumprod"][tg] * img
            self.num_timesteps = tg
        for t in reversed(range(self.num_timesteps)):
            if cfg["branch_out"] and t == self.num_timesteps - 1:
                img = [img, img]  # ddpm.py:955-957
            img, x0 = self._p_sample(img, mask, min_max_val, cond, t, draw)
            if self.trace is not None:
                self.trace.append((t, img, x0))
        if not self.ctor_start_intermediate and self.ctor_branch_out:  # ddpm.py:965-970
            img = torch.stack(img, dim=0) if isinstance(img, list) else torch.stack((img, img), dim=0)
        return img


def ddim_sample(smp: "Sampler", cond: Tensor, mask: Tensor, min_max_val, noise_tape: Sequence[Tensor], sampling_timesteps: int,
                eta: float = 0.0):
    """Restatement of `GaussianDiffusion.sample` -> `ddim_sample` (ddpm.py:1078-1124, 979-1075) for objective pred_x0.

    `noise_tape[0]` is x_T, then one draw per step that has a successor (the reference draws `randn_like` only after its
    `time_next < 0` early-out, ddpm.py:1009-1020).  Returns a tensor, or the list [x_out, x_in] when the branches were
    never composited (the reference returns `img` as it is, ddpm.py:1073-1076)."""
    cfg = smp.config
    if cfg["branch_out"] is False:
        cfg["branch_out"] = smp.ctor_branch_out
    if cfg["start_intermediate"] is False:
        cfg["start_intermediate"] = smp.ctor_start_intermediate
    smp.ctor_start_intermediate = bool(cfg["start_intermediate"])
    if cfg["ood_AD"] or cfg["ood_confidence"]:
        cfg["mask_cond"] = True
        cfg["mask_x"] = True
    if cfg["branch_out"]:
        u = torch.unique(mask)
        if len(u) == 1 and float(u[0]) == 1.0:
            cfg["mask_cond"] = cfg["mask_x"] = cfg["branch_out"] = cfg["start_intermediate"] = False
    lo, hi = float(min_max_val[0]), float(min_max_val[1])
    T = smp.num_timesteps
    times = torch.linspace(-1, T - 1, steps=sampling_timesteps + 1)  # ddpm.py:984
    times = list(reversed(times.int().tolist()))
    pairs = list(zip(times[:-1], times[1:]))
    start_ddim = times[-cfg["start_timestep"] - 2]  # ddpm.py:987
    ac, sr, srm1 = smp.buf["alphas_cumprod"], smp.buf["sqrt_recip_alphas_cumprod"], smp.buf["sqrt_recipm1_alphas_cumprod"]
    it = iter(noise_tape)
    draw = lambda: next(it).clone()
    img = draw()
    eps_of = lambda xt, t, x0: (sr[t] * xt - x0) / srm1[t]  # ddpm.py:637-641
    for time, time_next in pairs:
        tt = torch.full((cond.shape[0],), time, dtype=torch.long)
        if cfg["branch_out"]:
            if time == T - 1:
                img = [img, img]  # ddpm.py:1003-1004
            # model_predictions(clip_x_start=True, rederive_pred_noise=True), ddpm.py:668-766
            bm = (mask >= 1.0).float()
            cond_out = cond * bm
            floor = 0.5 if cfg["data"] == "mnist" else 0.95
            cond_in = (cond * torch.clip(1.0 - bm, floor, 1.0)).float()
            o_out = smp._model(img[0], cond_out, tt)
            o_in = smp._model(img[1], cond_in, tt)
            if cfg["mask_x"]:
                assert len(torch.unique(bm)) == 2, "mask should be binary"
                o_out = torch.where(bm == 0.0, torch.tensor(lo), o_out * bm)
                d = cfg["data"]
                if any(s in d for s in _NON_MRI) and "mri" not in d:
                    o_out = cond_out
            x0_out, x0_in = o_out.clamp(lo, hi), o_in.clamp(lo, hi)
            e_out, e_in = eps_of(img[0], time, x0_out), eps_of(img[1], time, x0_in)
            if time_next < 0:
                img = [x0_out, x0_in]
                continue
            alpha, alpha_next = ac[time], ac[time_next]
            sigma = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
            c = (1 - alpha_next - sigma ** 2).sqrt()
            noise = draw()
            if time <= start_ddim and cfg["start_intermediate"]:  # ddpm.py:1022-1043
                cfg["branch_out"] = False
                cfg["mask_x"] = False
                x0 = torch.where(x0_out == 0.0, x0_in, x0_out).clamp(lo, hi)
                xo, xi = e_out * bm, e_in * (1.0 - bm)
                assert bool((xo == 0).any()) and bool((xi == 0).any()), "x_out and x_in should be masked"
                eps = torch.where(xo == 0.0, xi, xo)
                img = x0 * alpha_next.sqrt() + c * eps + sigma * noise
            else:
                img = [x0_out * alpha_next.sqrt() + c * e_out + sigma * noise, x0_in * alpha_next.sqrt() + c * e_in + sigma * noise]
        else:
            # ddpm.py:716, 731-761 with clip_x_start (and the idempotent re-clamp of 1050-1051); eps is re-derived from the clamped x0
            x0 = smp._x0_from_output(img, smp._model(img, cond, tt), time).clamp(lo, hi)
            eps = eps_of(img, time, x0)
            if time_next < 0:
                img = x0
                continue
            alpha, alpha_next = ac[time], ac[time_next]
            sigma = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
            c = (1 - alpha_next - sigma ** 2).sqrt()
            noise = draw()
            img = x0 * alpha_next.sqrt() + c * eps + sigma * noise
    return img


def psnr(a: Tensor, b: Tensor, peak: float) -> float:
    mse = float(((a.double() - b.double()) ** 2).mean())
    return float("inf") if mse == 0 else 10.0 * math.log10(peak * peak / mse)
