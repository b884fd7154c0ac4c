"""Host logic: ingestion of the reference's `model-<milestone>.pt` layout (ddpm.py:1495-1527, test.py:139-147)."""
import pytest
import torch

from localdiffusion_hallucination_b200 import GaussianDiffusion, load_reference_checkpoint
from tests import util
from tests.golden import cases


def _diffusion(seed):
    m = util.make_model("mnist", "fp32", seed=seed)
    return GaussianDiffusion(cases.base_config("mnist", 2), m, image_size=32, timesteps=50, objective="pred_x0")


def _reference_style_ckpt(gd, step=7):
    sd = gd.state_dict()
    ema = {"initted": torch.tensor(True), "step": torch.tensor(step)}
    ema.update({"ema_model." + k: v.clone() + 0.25 for k, v in sd.items()})      # EMA copy differs from the online weights
    ema.update({"online_model." + k: v.clone() for k, v in sd.items()})
    return {"step": step, "model": {k: v.clone() for k, v in sd.items()}, "opt": {}, "ema": ema, "scaler": None}


def test_loads_ema_weights_like_test_py(tmp_path):
    src, dst = _diffusion(1), _diffusion(2)
    ck = _reference_style_ckpt(src)
    path = tmp_path / "model-best3.pt"
    torch.save(ck, path)
    assert load_reference_checkpoint(dst, str(path)) == 7
    for k, v in src.state_dict().items():
        assert torch.equal(dst.state_dict()[k], v + 0.25), k
    load_reference_checkpoint(dst, ck, use_ema=False)
    for k, v in src.state_dict().items():
        assert torch.equal(dst.state_dict()[k], v), k


def test_state_dict_layout_matches_reference_counts():
    gd = _diffusion(0)
    keys = list(gd.state_dict())
    assert sum(k.startswith("model.") for k in keys) + 13 == len(keys)          # Unet tensors + 13 schedule buffers
    assert "model.conv_fusion.mlp.1.weight" in keys                             # dead but present in the reference state_dict


def test_strict_mismatch_raises():
    gd = _diffusion(0)
    ck = _reference_style_ckpt(gd)
    del ck["ema"]["ema_model.model.init_conv.weight"]
    with pytest.raises(RuntimeError):
        load_reference_checkpoint(gd, ck)
    with pytest.raises(KeyError):
        load_reference_checkpoint(gd, {"step": 1, "model": {}})


def test_parent_load_state_dict_invalidates_the_engine():
    """`diffusion.load_state_dict(...)` (ddpm.py:1517) never calls the child Unet's load_state_dict; the engine must be
    invalidated from a hook that runs under the parent load, and in-place parameter updates must be noticed too."""
    gd = _diffusion(0)
    calls = []
    gd.model.release_engine = lambda: calls.append(1)
    gd.load_state_dict(gd.state_dict())
    assert calls, "the Unet post hook did not run under the parent's load_state_dict"
    v0 = sum(p._version for p in gd.model.parameters())
    with torch.no_grad():
        next(gd.model.parameters()).mul_(1.0)           # EMA / optimiser style in-place update
    assert sum(p._version for p in gd.model.parameters()) != v0   # what Unet.engine() compares


def test_weights_only_loading_is_the_default(tmp_path):
    src, dst = _diffusion(1), _diffusion(2)
    ck = _reference_style_ckpt(src)
    ck["opt"] = {"state": {}, "param_groups": [{"lr": 1e-4}]}
    path = tmp_path / "model-1.pt"
    torch.save(ck, path)
    assert load_reference_checkpoint(dst, str(path)) == 7   # plain containers + tensors load with weights_only=True
