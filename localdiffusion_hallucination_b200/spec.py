"""Parameter layout of the denoiser: every `Unet.state_dict()` key with its shape and initialiser,
generated procedurally from the constructor arguments.

`param_specs()` yields entries in the order the reference *constructs* its sub-modules
(`ddpm.py:308-398`, `unet_model.py:104-118`), which is the order PyTorch's global RNG is consumed
in, so `torch.manual_seed(s); Unet(...)` reproduces the reference's default initialisation bit for
bit (checked in tests/golden/make_golden.py).  `TOP_LEVEL_ORDER` is the order in which the
reference *registers* its top-level children (it creates the empty `downs`/`ups` lists before the
middle blocks, `ddpm.py:359-360`), which is the key order of `state_dict()`.
"""
from dataclasses import dataclass
from typing import List, Optional, Tuple

TOP_LEVEL_ORDER = ("cond_model", "init_conv", "time_mlp", "downs", "ups", "mid_block1", "mid_attn",
                   "mid_block2", "conv_fusion", "final_res_block", "final_conv")


@dataclass(frozen=True)
class ParamSpec:
    key: str
    shape: Tuple[int, ...]
    init: str  # 'kaiming' (conv/linear weight), 'bias' (uniform +-1/sqrt(fan_in)), 'ones', 'zeros'
    fan_in: int = 0


def _conv(key, co, ci, k, bias=True):
    out = [ParamSpec(key + ".weight", (co, ci, k, k), "kaiming", ci * k * k)]
    if bias:
        out.append(ParamSpec(key + ".bias", (co,), "bias", ci * k * k))
    return out


def _linear(key, co, ci):
    return [ParamSpec(key + ".weight", (co, ci), "kaiming", ci), ParamSpec(key + ".bias", (co,), "bias", ci)]


def _gn(key, c):
    return [ParamSpec(key + ".weight", (c,), "ones"), ParamSpec(key + ".bias", (c,), "zeros")]


def _resnet(key, ci, co, tdim):  # ddpm.py:189-198
    out = _linear(key + ".mlp.1", 2 * co, tdim)
    out += _conv(key + ".block1.proj", co, ci, 3) + _gn(key + ".block1.norm", co)
    out += _conv(key + ".block2.proj", co, co, 3) + _gn(key + ".block2.norm", co)
    if ci != co:
        out += _conv(key + ".res_conv", co, ci, 1)
    return out


def _attn(key, c, hid, full):  # ddpm.py:214-232, 253-269
    out = [ParamSpec(key + ".norm.g", (1, c, 1, 1), "ones")]
    out += _conv(key + ".to_qkv", 3 * hid, c, 1, bias=False)
    if full:
        out += _conv(key + ".to_out", c, hid, 1)
    else:
        out += _conv(key + ".to_out.0", c, hid, 1) + [ParamSpec(key + ".to_out.1.g", (1, c, 1, 1), "ones")]
    return out


def _cond_block(key, ci, cm, co):  # unet_model.py:18-34
    out = _conv(key + ".convblock.0", cm, ci, 3) + _gn(key + ".convblock.1", cm)
    out += _conv(key + ".convblock.3", co, cm, 3) + _gn(key + ".convblock.4", co)
    out += _conv(key + ".identity.0", co, ci, 3) + _gn(key + ".identity.1", co)
    return out


def cond_is_deep(mode: str) -> bool:
    """`ResUnet` builds `mid_conv` only for these modes (unet_model.py:115)."""
    return mode in ("mri", "mvtec", "mvtecGray")


def cond_returns_early(mode: str) -> bool:
    """`ResUnet.forward` returns after block 3 for these modes (unet_model.py:131)."""
    return mode in ("mnist", "mvtecSR")


def param_specs(dim, init_dim, dim_mults, channels, attn_heads, attn_dim_head, full_attn, mode, out_dim) -> List[ParamSpec]:
    dims = [init_dim] + [dim * m for m in dim_mults]
    L = len(dim_mults)
    tdim = dim * 4
    hid = attn_heads * attn_dim_head
    cin = 3 if ("mvtec" in mode and "mvtecGray" not in mode) else 1  # unet_model.py:94-99
    s: List[ParamSpec] = []
    s += _cond_block("cond_model.residual_conv1.0", cin, 32, 32)
    s += _cond_block("cond_model.residual_conv2.0", 32, 32, 64)
    s += _cond_block("cond_model.residual_conv3.0", 64, 64, 128)
    if cond_is_deep(mode):
        s += _cond_block("cond_model.mid_conv.0", 128, 128, 256)
    s += _conv("init_conv", init_dim, channels, 7)
    s += _linear("time_mlp.1", tdim, dim) + _linear("time_mlp.3", tdim, tdim)
    for i in range(L):
        di, dn = dims[i], dims[i + 1]
        p = f"downs.{i}"
        s += _resnet(p + ".0", di, di, tdim) + _resnet(p + ".1", di, di, tdim) + _attn(p + ".2", di, hid, full_attn[i])
        s += _conv(p + ".3.1", dn, 4 * di, 1) if i < L - 1 else _conv(p + ".3", dn, di, 3)
    mid = dims[-1]
    s += _resnet("mid_block1", mid, mid, tdim) + _attn("mid_attn", mid, hid, True) + _resnet("mid_block2", mid, mid, tdim)
    s += _resnet("conv_fusion", 2 * mid, mid, tdim)
    for i in range(L):
        di, dn = dims[L - 1 - i], dims[L - i]
        p = f"ups.{i}"
        s += _resnet(p + ".0", dn + di, dn, tdim) + _resnet(p + ".1", dn + di, dn, tdim)
        s += _attn(p + ".2", dn, hid, full_attn[L - 1 - i])
        s += _conv(p + ".3.1", di, dn, 3) if i < L - 1 else _conv(p + ".3", di, dn, 3)
    s += _resnet("final_res_block", 2 * dim, dim, tdim)
    s += _conv("final_conv", out_dim, dim, 1)
    return s
