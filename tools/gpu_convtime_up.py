"""Isolated times of the three up-sampling convolutions of the mri UNet (N = 32 images at 256x256 input): the replicate-on-load
kernel (2) against the folded low-resolution convolution with pixel-shuffle output (3).  Development aid."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from localdiffusion_hallucination_b200 import _lib  # noqa

lib = _lib.lib()
torch.zeros(1, device="cuda")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
for (c0, hw, co) in [(64, 256, 32), (128, 128, 64), (256, 64, 128)]:
    for kern in (2, 3):
        ms = C.c_float(0)
        rc = lib.ld_debug_conv_time(kern, c0, 0, N, hw, hw, 1, co, 3, 10, C.byref(ms), None)
        by = N * hw * hw * (c0 / 4 + co) * 2
        print(f"k{kern} up C{c0}->{co} @{hw} N={N}: {ms.value*1000:8.1f} us  {by/ms.value/1e6:8.1f} GB/s rc={rc}")
