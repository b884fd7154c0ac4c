"""Development aid: time the 32->32 3x3 conv with parts of the kernel disabled (LD_CONV_DBG bit mask)."""
import ctypes as C, os, subprocess, sys
if len(sys.argv) > 1:
    import torch
    sys.path.insert(0, ".")
    from localdiffusion_hallucination_b200 import _lib
    lib = _lib.lib(); torch.zeros(1, device="cuda")
    ms = C.c_float(0)
    for (c0, hw, co) in [(32, 256, 32), (64, 128, 64)]:
        lib.ld_debug_conv_time(2, c0, 0, 32, hw, hw, 0, co, 3, 5, C.byref(ms), None)
        print(f"dbg={os.environ.get('LD_CONV_DBG','0'):>2} C{c0}->{co} @{hw}: {ms.value*1000:8.1f} us")
else:
    for d in (0, 1, 2, 4, 7):
        subprocess.run([sys.executable, __file__, "x"], env=dict(os.environ, LD_CONV_DBG=str(d)))
