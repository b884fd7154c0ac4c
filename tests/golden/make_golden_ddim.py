"""Golden fixtures for the DDIM branch sampler (`ddim_sample`, ddpm.py:979-1075): runs the UNMODIFIED reference on CPU
through oracle/ref_harness.py and checks the oracle restatement against it.  Build container only.

    python tests/golden/make_golden_ddim.py      ->  tests/golden/golden_ddim.npz
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ld_oracle as lo  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402
from tests.golden import cases  # noqa: E402
from tests.golden.make_golden import hp_of, ref_model  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def run_case(ddpm, key, name, model, S, B, cond, mask, mm):
    c = cases.DDIM_CASES[key]
    cfg_ref = cases.base_config(c["data"], c["s"], **c.get("cfg", {}))
    gd = ddpm.GaussianDiffusion(cfg_ref, model, image_size=S, timesteps=c["T"], sampling_timesteps=c["steps"], beta_schedule=c["sched"],
                                objective="pred_x0", ddim_sampling_eta=c["eta"], auto_normalize=False).eval()
    assert gd.is_ddim_sampling
    tape = cases.noise_tape(B, S, c["steps"])
    with rh.noise_tape(list(tape)) as nt:
        ret = gd.sample(cond, None, batch_size=B, mask=mask, min_max_val=mm)
        used = nt.i
    cfg_or = cases.base_config(c["data"], c["s"], **c.get("cfg", {}))
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    smp = lo.Sampler(cfg_or, sd, hp_of(name), image_size=S, timesteps=c["T"], beta_schedule=c["sched"])
    with torch.no_grad():
        o2 = lo.ddim_sample(smp, cond, mask, mm, list(tape), c["steps"], c["eta"])
    is_pair = isinstance(ret, list)
    assert is_pair == isinstance(o2, list), key
    r = torch.stack(ret) if is_pair else ret
    o = torch.stack(o2) if is_pair else o2
    err = float((r - o).abs().max())
    assert err < 5e-5, (key, err)
    assert cfg_ref == cfg_or, (cfg_ref, cfg_or)
    print(f"  {key}: T={c['T']} steps={c['steps']} s={c['s']} eta={c['eta']} pair={is_pair} draws={used} unet_calls={smp.unet_calls} "
          f"ref-vs-oracle {err:.2e} range [{float(r.min()):.3f}, {float(r.max()):.3f}]")
    return dict(out=r.numpy(), pair=np.array(int(is_pair)), draws=np.array(used), unet_calls=np.array(smp.unet_calls), cfg_after=np.array(repr(cfg_ref)))


def main():
    assert rh.available(), "reference tree not found"
    ddpm = rh.load_reference()
    torch.set_num_threads(os.cpu_count())
    os.chdir(tempfile.mkdtemp())
    G = {}
    m = ref_model(ddpm, "mnist")
    S = 32
    cond, mask = cases.cond_uniform(8, S), cases.mask_left_columns(8, S)
    for key in ("d1", "d1eta", "d1pair", "d1s0", "d1nobranch", "d1ones"):
        B = cases.DDIM_CASES[key]["B"]
        mk = torch.ones(B, 1, S, S) if key == "d1ones" else mask[:B]
        G.update({f"{key}_{k}": v for k, v in run_case(ddpm, key, "mnist", m, S, B, cond[:B], mk, cases.MNIST_MIN_MAX).items()})
    m = ref_model(ddpm, "mri")
    S, B = 64, 2
    cond, mask = cases.mri_like(B, S)
    G.update({f"d2_{k}": v for k, v in run_case(ddpm, "d2", "mri", m, S, B, cond, mask, cases.MRI_MIN_MAX).items()})
    # schedule of one case, to pin the host-side coefficient construction
    from localdiffusion_hallucination_b200 import GaussianDiffusion, Unet
    c = cases.DDIM_CASES["d1eta"]
    gd = GaussianDiffusion(cases.base_config(c["data"], c["s"]), Unet(**cases.MODEL_KW["mnist"]), image_size=32, timesteps=c["T"],
                           sampling_timesteps=c["steps"], beta_schedule=c["sched"], objective="pred_x0", ddim_sampling_eta=c["eta"])
    times, coefs, fuse = gd.ddim_schedule()
    G["d1eta_times"], G["d1eta_coefs"], G["d1eta_fuse"] = np.array(times), coefs.numpy(), np.array(fuse)
    path = os.path.join(OUT, "golden_ddim.npz")
    np.savez_compressed(path, **G)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
