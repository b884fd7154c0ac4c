"""Golden fixtures that pin the BENCHMARKED configurations to the live reference.

    python tests/golden/make_golden_big.py        # build container only; ~10 min of CPU

Runs the UNMODIFIED reference classes (`/root/reference/ddpm.py`, through oracle/ref_harness.py) on CPU and writes
`tests/golden/golden_big.npz`:

  u256      mri Unet.forward at 256x256, B=2                                   (the bench resolution; other kernel paths than 64x64)
  s256      mri sampler at 256x256, B=2, T=6, s=2                               (BASELINE configs[1] shape, short chain)
  long      mri sampler at 64x64, B=2, **T=1000**, s=2, x0 snapshot every 100 steps  (the bench chain length; bf16 accumulation)
  c4        mri_attn8 sampler (full attention at two levels, 8 heads) at 128x128, B=1, T=12, s=3   (BASELINE configs[3])
  c5        mri sampler at 512x512, B=1, T=3, s=1: stride-2 sub-sample + fp64 checksums of the full frame (BASELINE configs[4])
  sw{s}     start_timestep sweep, mri 64x64, B=1, T=8, s in {0,2,4,6}: outputs + UNet call counts `2(T-s)+s` (configs[4])
  gt        use_gt start (ddpm.py:937-944): mnist model, T=24, use_gt_timestep=10

Every case is also run through oracle/ld_oracle.py and asserted equal, so the oracle is pinned at these sizes too.
"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ld_oracle as lo  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402
from tests.golden import cases  # noqa: E402
from tests.golden.make_golden import hp_of, ref_model  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def checksums(t):
    d = t.double()
    return np.array([float(d.sum()), float((d * d).sum()), float(d.abs().max())], dtype=np.float64)


def run_case(ddpm, name, model, data, S, T, s, B, cond, mask, mm, gt=None, tol=5e-5, **cfgkw):
    """reference + oracle on the same inputs; returns (ret, x0 list of the reference, cfg repr, unet calls, oracle trace)"""
    cfg_ref = cases.base_config(data, s, **cfgkw)
    gd = ddpm.GaussianDiffusion(cfg_ref, model, image_size=S, timesteps=T, beta_schedule="sigmoid", objective="pred_x0",
                                auto_normalize=False).eval()
    steps = int(cfg_ref["use_gt_timestep"]) if cfg_ref.get("use_gt") else T
    tape = cases.noise_tape(B, S, steps)
    t0 = time.time()
    with rh.noise_tape(list(tape)):
        ret, x0_lst, _ = gd.sample(cond, gt, batch_size=B, mask=mask, min_max_val=mm, return_all_outputs=True)
    t_ref = time.time() - t0
    cfg_or = cases.base_config(data, s, **cfgkw)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    smp = lo.Sampler(cfg_or, sd, hp_of(name), image_size=S, timesteps=T, trace=[])
    with torch.no_grad():
        o2 = smp.sample(cond, mask, mm, list(tape), gt=gt)
    err = float((ret - o2).abs().max())
    assert err < tol, (name, S, T, err)
    assert cfg_ref == cfg_or, (cfg_ref, cfg_or)
    print(f"  {name} S={S} T={T} s={s} B={B}: reference {t_ref:.1f}s, ref-vs-oracle max abs {err:.2e}, range [{float(ret.min()):.3f}, {float(ret.max()):.3f}]",
          flush=True)
    return ret, x0_lst, repr(cfg_ref), smp.unet_calls


def x0_tensor(e):
    return torch.stack(e) if isinstance(e, list) else e


def main():
    assert rh.available(), "reference tree not found"
    ddpm = rh.load_reference()
    torch.set_num_threads(os.cpu_count())
    os.chdir(tempfile.mkdtemp())
    os.makedirs("fusion_test", exist_ok=True)  # ddpm.py:793-794 np.save target
    G = {}
    mm = cases.MRI_MIN_MAX

    # ---- u256: UNet forward at the bench resolution ------------------------------------------------------
    m = ref_model(ddpm, "mri")
    B, S = 2, 256
    x, cond, t = cases.noise_tape(B, S, 1)[0], cases.cond_uniform(B, S), torch.tensor([999, 17])
    with torch.no_grad():
        y = m(x, cond, t)
        y2 = lo.unet_forward({k: v for k, v in m.state_dict().items()}, hp_of("mri"), x, cond, t)
    assert float((y - y2).abs().max()) < 5e-5
    G["u256_out"] = y.numpy()
    print(f"u256: out std {float(y.std()):.4f}", flush=True)

    # ---- s256: short sampler chain at the bench resolution ------------------------------------------------
    cond, mask = cases.mri_like(B, S)
    ret, x0s, cfg, calls = run_case(ddpm, "mri", m, "mri", S, 6, 2, B, cond, mask, mm)
    G["s256_out"], G["s256_x0_first"], G["s256_x0_last"] = ret.numpy(), x0_tensor(x0s[0]).numpy(), x0_tensor(x0s[-1]).numpy()
    G["s256_cfg_after"], G["s256_unet_calls"] = np.array(cfg), np.array(calls)

    # ---- long: the bench chain length (T = 1000) on 64x64 ---------------------------------------------------
    S, T, s = 64, 1000, 2
    cond, mask = cases.mri_like(B, S)
    ret, x0s, cfg, calls = run_case(ddpm, "mri", m, "mri", S, T, s, B, cond, mask, mm, tol=1e-3)
    G["long_out"] = ret.numpy()
    snaps = list(range(99, T, 100))  # loop iterations 99, 199, .. 999 (t = T-1-i)
    G["long_snap_iters"] = np.array(snaps)
    for i in snaps:
        e = x0s[i]
        G[f"long_x0_{i}"] = (torch.stack(e) if isinstance(e, list) else torch.stack((e, e))).numpy()
    G["long_cfg_after"], G["long_unet_calls"] = np.array(cfg), np.array(calls)
    assert calls == 2 * (T - s) + s

    # ---- c4: attention-heavy model (configs[3]) ---------------------------------------------------------------
    m4 = ref_model(ddpm, "mri_attn8")
    S = 128
    cond, mask = cases.mri_like(1, S)
    ret, x0s, cfg, calls = run_case(ddpm, "mri_attn8", m4, "mri", S, 12, 3, 1, cond, mask, mm)
    G["c4_out"], G["c4_x0_last"] = ret.numpy(), x0_tensor(x0s[-1]).numpy()
    G["c4_cfg_after"], G["c4_unet_calls"] = np.array(cfg), np.array(calls)
    del m4

    # ---- c5: 512x512 (configs[4]); the full frame is checked against the live oracle on the GPU box, which is pinned here
    S = 512
    cond, mask = cases.mri_like(1, S)
    ret, x0s, cfg, calls = run_case(ddpm, "mri", m, "mri", S, 3, 1, 1, cond, mask, mm)
    G["c5_out_sub2"], G["c5_out_sums"] = ret[:, :, ::2, ::2].contiguous().numpy(), checksums(ret)
    G["c5_x0_last_sums"] = checksums(x0_tensor(x0s[-1]))
    G["c5_cfg_after"], G["c5_unet_calls"] = np.array(cfg), np.array(calls)

    # ---- start_timestep sweep (configs[4]): forwards per image = 2(T-s)+s ---------------------------------------
    S, T = 64, 8
    cond, mask = cases.mri_like(1, S)
    for s in (0, 2, 4, 6):
        ret, x0s, cfg, calls = run_case(ddpm, "mri", m, "mri", S, T, s, 1, cond, mask, mm)
        assert calls == 2 * (T - s) + s
        G[f"sw{s}_out"], G[f"sw{s}_cfg_after"], G[f"sw{s}_unet_calls"] = ret.numpy(), np.array(cfg), np.array(calls)

    # ---- use_gt start (ddpm.py:937-944) ----------------------------------------------------------------------------
    mn = ref_model(ddpm, "mnist")
    S, T, B = 32, 24, 2
    cond, mask = cases.cond_uniform(B, S), cases.mask_left_columns(B, S)
    gt = cases.cond_uniform(B, S, seed=7)
    ret, x0s, cfg, calls = run_case(ddpm, "mnist", mn, "mri", S, T, 2, B, cond, mask, cases.MNIST_MIN_MAX, gt=gt, use_gt=True,
                                    use_gt_timestep=10)
    G["gt_out"], G["gt_cfg_after"], G["gt_unet_calls"] = ret.numpy(), np.array(cfg), np.array(calls)
    assert len(x0s) == 10

    path = os.path.join(OUT, "golden_big.npz")
    np.savez_compressed(path, **G)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
