import ctypes as C, os, sys
import torch
sys.path.insert(0, ".")
from localdiffusion_hallucination_b200 import _lib
lib = _lib.lib(); torch.zeros(1, device="cuda")
ms = C.c_float(0)
lib.ld_debug_conv_time(2, 32, 0, 32, 256, 256, 0, 32, 3, 3, C.byref(ms), None)
print(f"dbg={os.environ.get('LD_CONV_DBG','0')}: {ms.value*1000:.1f} us")
