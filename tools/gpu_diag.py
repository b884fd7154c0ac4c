"""Diagnostic sweep for the GPU box: prints error metrics for every layer / kernel without stopping."""
import sys
import time
import traceback

import torch

sys.path.insert(0, ".")
from oracle import ld_oracle as lo  # noqa: E402
from tests import util  # noqa: E402
from tests.golden import cases  # noqa: E402
from tests import test_gpu_kernels as tk  # noqa: E402
from tests import test_gpu_unet as tu  # noqa: E402

DEV = "cuda:0"
print(torch.cuda.get_device_name(0))
for case in tk.CONV_CASES:
    for kernel in (0, 1, 2):
        try:
            tk.test_conv_kernels_match_torch(case, kernel)
            print("conv", case, kernel, "ok")
        except Exception as e:  # noqa
            print("conv", case, kernel, "FAIL", type(e).__name__, str(e)[:200])
for name, S, B, ts in tu.UNET_CASES:
    for prec, opts in (("fp32", {}), ("bf16", dict(use_tc=0)), ("bf16", dict(use_tc=1))):
        try:
            m = util.make_model(name, prec, device=DEV, debug_keep=1, **opts)
            x, cond, t = cases.noise_tape(B, S, 1)[0], cases.cond_uniform(B, S), torch.tensor(ts)
            t0 = time.time()
            y = m(x.to(DEV), cond.to(DEV), t.to(DEV))
            torch.cuda.synchronize()
            taps = {}
            with torch.no_grad():
                ref = lo.unet_forward(util.cpu_state_dict(m), util.hp_of(name), x, cond, t, taps=taps)
            got = tu.fetch_taps(m)
            print(f"unet {name} {prec} {opts}: out rel {util.rel_err(y, ref):.3e} ({time.time()-t0:.2f}s)")
            for k in taps:
                print(f"    {k:20s} rel {util.rel_err(got[k], taps[k]):.3e}")
        except Exception as e:  # noqa
            traceback.print_exc()
