"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference classes
(`/root/reference/ddpm.py`) on CPU through oracle/ref_harness.py.  Only runs in the build container.

    python tests/golden/make_golden.py

For every case the reference output is also compared with the CPU restatement in
oracle/ld_oracle.py, and the reference's seed-0 default initialisation with the product's
`Unet` parameter construction, so a drift in either shows up here first.
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

OUT = os.path.dirname(os.path.abspath(__file__))


def ref_model(ddpm, name, seed=0):
    torch.manual_seed(seed)
    return ddpm.Unet(**cases.MODEL_KW[name]).eval()


def hp_of(name):
    kw = dict(cases.MODEL_KW[name])
    return lo.UnetHP(dim=kw["dim"], init_dim=kw["init_dim"], dim_mults=kw.get("dim_mults", (1, 2, 4, 8)),
                     full_attn=kw.get("full_attn", (False, False, False, True)), heads=kw.get("attn_heads", 4), mode=kw["mode"])


def run_sampler_case(ddpm, name, model, data, S, T, s, B, cond, mask, mm, schedule="sigmoid", **cfgkw):
    cfg_ref = cases.base_config(data, s, **cfgkw)
    gd = ddpm.GaussianDiffusion(cfg_ref, model, image_size=S, timesteps=T, beta_schedule=schedule, objective="pred_x0",
                                auto_normalize=False).eval()
    tape = cases.noise_tape(B, S, T)
    with rh.noise_tape(list(tape)):
        out = gd.sample(cond, None, batch_size=B, mask=mask, min_max_val=mm, return_all_outputs=True)
    ret, x0_lst, _ = out
    # oracle restatement on the same inputs
    cfg_or = cases.base_config(data, s, **cfgkw)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    smp = lo.Sampler(cfg_or, sd, hp_of(name), image_size=S, timesteps=T, beta_schedule=schedule, trace=[])
    with torch.no_grad():
        o2 = smp.sample(cond, mask, mm, list(tape))
    err = float((ret - o2).abs().max())
    assert err < 5e-5, (name, data, err)
    assert cfg_ref == cfg_or, (cfg_ref, cfg_or)
    # x0 of the last step (single tensor after fusion, pair otherwise)
    last = x0_lst[-1]
    last = torch.stack(last) if isinstance(last, list) else last
    first = x0_lst[0]
    first = torch.stack(first) if isinstance(first, list) else first
    print(f"  {name}/{data} S={S} T={T} s={s} B={B}: ref-vs-oracle max abs {err:.2e}, out range [{float(ret.min()):.3f}, {float(ret.max()):.3f}]")
    return dict(out=ret.numpy(), x0_first=first.numpy(), x0_last=last.numpy(), cfg_after=repr(cfg_ref), unet_calls=smp.unet_calls)


def main():
    assert rh.available(), "reference tree not found"
    ddpm = rh.load_reference()
    torch.set_num_threads(os.cpu_count())
    wd = tempfile.mkdtemp()
    os.chdir(wd)
    os.makedirs("fusion_test", exist_ok=True)  # ddpm.py:793-794 np.save target
    from localdiffusion_hallucination_b200 import GaussianDiffusion, Unet  # product shims (CPU construction only)

    G = {}
    # ---- default-initialisation parity + weight fingerprints ---------------------------------------
    for name in cases.MODEL_KW:
        m = ref_model(ddpm, name)
        torch.manual_seed(0)
        mine = Unet(**cases.MODEL_KW[name])
        a, b = m.state_dict(), mine.state_dict()
        assert list(a.keys()) == list(b.keys()), name
        for k in a:
            assert torch.equal(a[k], b[k]), (name, k)
        G[f"wsum_{name}"] = np.array(cases.weight_checksum(a), dtype=np.float64)
        print(f"init parity ok: {name} ({len(a)} tensors)")

    # ---- schedules: the reference's registered buffers ------------------------------------------------
    m = ref_model(ddpm, "mnist")
    for sched, T in (("sigmoid", 1000), ("sigmoid", 50), ("linear", 100), ("cosine", 200)):
        gd = ddpm.GaussianDiffusion(cases.base_config(), m, image_size=32, timesteps=T, beta_schedule=sched, objective="pred_x0")
        mine = GaussianDiffusion(cases.base_config(), Unet(**cases.MODEL_KW["mnist"]), image_size=32, timesteps=T, beta_schedule=sched, objective="pred_x0")
        for k in ("betas", "posterior_mean_coef1", "posterior_mean_coef2", "posterior_log_variance_clipped", "sqrt_recip_alphas_cumprod", "loss_weight"):
            assert torch.equal(getattr(gd, k), getattr(mine, k)), (sched, T, k)
            G[f"sched_{sched}_{T}_{k}"] = getattr(gd, k).numpy()
        assert list(gd.state_dict().keys()) == list(mine.state_dict().keys())
    print("schedule parity ok")

    # ---- UNet forwards ---------------------------------------------------------------------------------
    for name, S, B, ts in (("mnist", 32, 2, [99, 3]), ("mri", 64, 2, [999, 17]), ("mri_attn8", 64, 1, [500])):
        m = ref_model(ddpm, name)
        x = cases.noise_tape(B, S, 1)[0]
        cond = cases.cond_uniform(B, S)
        t = torch.tensor(ts)
        with torch.no_grad():
            y = m(x, cond, t)
            feat = m.cond_model(cond)
            y2 = lo.unet_forward({k: v for k, v in m.state_dict().items()}, hp_of(name), x, cond, t)
        assert float((y - y2).abs().max()) < 2e-5
        G[f"unet_{name}_out"] = y.numpy()
        G[f"unet_{name}_feat_mean"] = feat.mean(dim=(2, 3)).numpy()
        print(f"unet golden: {name} out std {float(y.std()):.4f}")

    # ---- sampler cases -----------------------------------------------------------------------------------
    m = ref_model(ddpm, "mnist")
    S, B = 32, 8
    cond, mask = cases.cond_uniform(B, S), cases.mask_left_columns(B, S)
    r = run_sampler_case(ddpm, "mnist", m, "mnist", S, 100, 2, B, cond, mask, cases.MNIST_MIN_MAX)  # BASELINE config 1
    G.update({f"c1_{k}": v for k, v in r.items()})
    r = run_sampler_case(ddpm, "mnist", m, "mri", S, 24, 5, 4, cond[:4], mask[:4], cases.MNIST_MIN_MAX)  # masked-OOD path
    G.update({f"c1mri_{k}": v for k, v in r.items()})
    r = run_sampler_case(ddpm, "mnist", m, "mri", S, 40, 0, 2, cond[:2], mask[:2], cases.MNIST_MIN_MAX, schedule="linear")
    G.update({f"c1s0_{k}": v for k, v in r.items()})
    # never fuse: stacked pair output (ddpm.py:965-970)
    r = run_sampler_case(ddpm, "mnist", m, "mri", S, 8, 2, 2, cond[:2], mask[:2], cases.MNIST_MIN_MAX, start_intermediate=False)
    G.update({f"c1pair_{k}": v for k, v in r.items()})
    # all-ones mask: vanilla DDPM fallback (ddpm.py:1110-1117)
    r = run_sampler_case(ddpm, "mnist", m, "mri", S, 8, 2, 2, cond[:2], torch.ones(2, 1, S, S), cases.MNIST_MIN_MAX)
    G.update({f"c1ones_{k}": v for k, v in r.items()})
    # branch_out disabled from the start
    r = run_sampler_case(ddpm, "mnist", m, "mri", S, 8, 2, 2, cond[:2], mask[:2], cases.MNIST_MIN_MAX, branch_out=False)
    G.update({f"c1nobranch_{k}": v for k, v in r.items()})
    # mri model, synthetic T1-like input with OOD blob, nonzero soft mask
    m = ref_model(ddpm, "mri")
    S, B = 64, 2
    cond, mask = cases.mri_like(B, S)
    r = run_sampler_case(ddpm, "mri", m, "mri", S, 12, 3, B, cond, mask, cases.MRI_MIN_MAX)
    G.update({f"c2s_{k}": v for k, v in r.items()})

    str_keys = {k: v for k, v in G.items() if isinstance(v, str)}
    arr = {k: v for k, v in G.items() if not isinstance(v, str)}
    for k, v in str_keys.items():
        arr[k] = np.array(v)
    np.savez_compressed(os.path.join(OUT, "golden.npz"), **arr)
    print("wrote", os.path.join(OUT, "golden.npz"), os.path.getsize(os.path.join(OUT, "golden.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
