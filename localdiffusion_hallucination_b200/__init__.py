"""B200-native (sm_100a) drop-in for the LocalDiffusion conditional reverse-diffusion sampler.

Public surface mirrors the reference's `ddpm.py`: `Unet`, `GaussianDiffusion`.
"""
from .checkpoint import extract_state_dict, load_reference_checkpoint
from .diffusion import GaussianDiffusion
from .unet import Unet
from . import producers

__all__ = ["Unet", "GaussianDiffusion", "load_reference_checkpoint", "extract_state_dict", "producers"]
