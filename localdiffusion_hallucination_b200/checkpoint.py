"""Ingestion of the reference's training checkpoints.

`Trainer.save` (ddpm.py:1495-1507) writes `{'step', 'model': GaussianDiffusion.state_dict(), 'opt',
'ema': EMA.state_dict(), 'scaler'}`; inference samples from the EMA copy (`trainer.ema.ema_model`,
test.py:144-147,393), whose tensors sit under the `ema_model.` prefix of `data['ema']` (ema_pytorch layout:
`ema_model.*`, `online_model.*`, `initted`, `step`).  Neither `ema_pytorch` nor `accelerate` is needed here.
"""
import torch

_EMA_PREFIX = "ema_model."


def extract_state_dict(ckpt, use_ema=True):
    """Return the `GaussianDiffusion.state_dict()`-shaped dict (`model.*` Unet tensors + schedule buffers) held by
    a reference checkpoint dict.  `use_ema=True` mirrors test.py (EMA weights); False takes `data['model']`."""
    if not isinstance(ckpt, dict):
        raise TypeError("expected the dict written by Trainer.save (ddpm.py:1495-1507)")
    if use_ema:
        if "ema" not in ckpt:
            raise KeyError("checkpoint has no 'ema' entry (ddpm.py:1503)")
        sd = {k[len(_EMA_PREFIX):]: v for k, v in ckpt["ema"].items() if k.startswith(_EMA_PREFIX)}
        if not sd:
            raise KeyError("no 'ema_model.*' tensors under data['ema']")
        return sd
    if "model" not in ckpt:
        raise KeyError("checkpoint has no 'model' entry (ddpm.py:1501)")
    return dict(ckpt["model"])


def load_reference_checkpoint(diffusion, ckpt, use_ema=True, strict=True, map_location="cpu", allow_pickle=False):
    """`Trainer.load` + `trainer.ema.ema_model` (ddpm.py:1509-1527, test.py:139-147) for the B200 `GaussianDiffusion`.
    `ckpt` is a path to `model-<milestone>.pt` or the already loaded dict.  Returns the checkpoint's `step`.
    Files are read with `weights_only=True`; `allow_pickle=True` falls back to full unpickling (the reference's `Trainer.save`
    stores the optimiser and GradScaler state too, which older torch versions cannot load tensors-only) -- trusted files only."""
    if not isinstance(ckpt, dict):
        try:   # tensors-only unpickling first: a checkpoint file is untrusted input
            ckpt = torch.load(ckpt, map_location=map_location, weights_only=True)
        except Exception:
            if not allow_pickle:
                raise
            ckpt = torch.load(ckpt, map_location=map_location, weights_only=False)
    sd = extract_state_dict(ckpt, use_ema)
    own = diffusion.state_dict()
    # buffers that only the training loss reads may be absent from / extra in older checkpoints: never block on them
    missing = [k for k in own if k not in sd]
    unexpected = [k for k in sd if k not in own]
    if strict and (missing or unexpected):
        raise RuntimeError(f"checkpoint does not match the model: missing {missing[:5]}, unexpected {unexpected[:5]}")
    diffusion.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)
    diffusion.model.release_engine()  # weights changed: re-pack on next use
    return int(ckpt.get("step", 0))
