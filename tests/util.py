"""Helpers shared by the test modules."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ld_oracle as lo  # noqa: E402
from tests.golden import cases  # noqa: E402


def hp_of(name):
    kw = dict(cases.MODEL_KW[name])
    return lo.UnetHP(dim=kw["dim"], init_dim=kw["init_dim"], dim_mults=kw.get("dim_mults", (1, 2, 4, 8)),
                     full_attn=kw.get("full_attn", (False, False, False, True)), heads=kw.get("attn_heads", 4), mode=kw["mode"])


def make_model(name, precision="fp32", seed=0, device=None, **opts):
    """Product `Unet` with the seed-0 default initialisation (bit-identical to the reference's)."""
    from localdiffusion_hallucination_b200.workload import make_model as mk

    return mk(name, precision, seed, device, **opts)


def cpu_state_dict(m):
    return {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_abs(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max())


def psnr(a, b, peak):
    mse = float(((a.double().cpu() - b.double().cpu()) ** 2).mean())
    return float("inf") if mse == 0 else 10.0 * math.log10(peak * peak / mse)
