"""In-situ conv variants at any shape (development aid): gpu_variant_shape.py VARIANT C0 C1 N S Cout   (variant as in gpu_conv_variant_one.py)"""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
from localdiffusion_hallucination_b200 import _lib
lib = _lib.lib(); torch.zeros(1, device="cuda")
var, c0, c1, N, S, co = (int(v) for v in sys.argv[1:7])
ms = C.c_float(0)
rc = lib.ld_debug_conv_variant_time(var, c0, c1, N, S, S, co, 10, C.byref(ms), None)
print(f"variant {var} C{c0}+{c1}->{co} @{S} N={N}: {ms.value*1000:.1f} us rc={rc}")
