"""One in-situ variant of the dominant 3x3 convolution (ld_debug_conv_variant_time) at bench size, for ncu captures and timings
(development aid).  usage: gpu_conv_variant_one.py VARIANT   (0 plain, 1 stats, 2 normalise-on-load + stats, 3 dual + stats)"""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
from localdiffusion_hallucination_b200 import _lib
lib = _lib.lib(); torch.zeros(1, device="cuda")
var = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ms = C.c_float(0)
c1 = 32 if var == 3 else 0
rc = lib.ld_debug_conv_variant_time(var, 32, c1, 32, 256, 256, 32, 10, C.byref(ms), None)
print(f"variant {var}: {ms.value*1000:.1f} us rc={rc}")
