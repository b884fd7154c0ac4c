"""Prints isolated kernel times for the conv shapes of the mri UNet at 256x256 (development aid)."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from localdiffusion_hallucination_b200 import _lib  # noqa

lib = _lib.lib()
torch.zeros(1, device="cuda")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
shapes = [  # C0, C1, H, Cout, ks, up
    (32, 0, 256, 32, 3, 0), (32, 32, 256, 32, 3, 0), (32, 0, 256, 384, 1, 0), (32, 32, 256, 32, 1, 0),
    (32, 0, 128, 32, 3, 0), (64, 0, 64, 64, 3, 0), (128, 0, 32, 128, 3, 0), (256, 0, 32, 256, 3, 0),
    (256, 256, 32, 256, 3, 0), (256, 128, 32, 256, 3, 0), (128, 0, 32, 384, 1, 0), (64, 0, 256, 32, 3, 1),
]
for (c0, c1, hw, co, ks, up) in shapes:
    for kern in (1, 2):
        ms = C.c_float(0)
        rc = lib.ld_debug_conv_time(kern, c0, c1, N, hw, hw, up, co, ks, 10, C.byref(ms), None)
        cin = c0 + c1
        fl = 2.0 * N * hw * hw * cin * co * ks * ks
        by = N * hw * hw * ((cin if not up else cin / 4) + co) * 2
        print(f"k{kern} C{c0}+{c1}->{co} {ks}x{ks} @{hw} up{up} N={N}: {ms.value*1000:8.1f} us  {fl/ms.value/1e9:8.1f} TF/s  {by/ms.value/1e6:8.1f} GB/s rc={rc}")
