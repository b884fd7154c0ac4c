"""One isolated conv timing (development aid).  env: LD_SAMPLER_LIB (alternative library), LD_CONV_DBG.
usage: gpu_conv_one.py [C0 C1 HW Cout ks up N [kernel]]  (kernel 2 = tcgen05, 3 = folded up-sampling conv)"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, ".")
from localdiffusion_hallucination_b200 import _lib
if os.environ.get("LD_SAMPLER_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["LD_SAMPLER_LIB"])
lib = _lib.lib(); torch.zeros(1, device="cuda")
a = [int(v) for v in sys.argv[1:]] or [32, 0, 256, 32, 3, 0, 32]
kern = a[7] if len(a) > 7 else 2
c0, c1, hw, co, ks, up, N = a[:7]
ms = C.c_float(0)
rc = lib.ld_debug_conv_time(kern, c0, c1, N, hw, hw, up, co, ks, 10, C.byref(ms), None)
print(f"lib={os.path.basename(_lib.LIB_PATH)} dbg={os.environ.get('LD_CONV_DBG','0')} C{c0}+{c1}->{co} k{ks} @{hw} up{up} N={N}: {ms.value*1000:.1f} us rc={rc}")
