"""`GaussianDiffusion` -- the reference's sampler API (ddpm.py:496-1125) hosted on the sm_100a engine.

Host side (this file): constructor contract, the 13 schedule buffers, the per-call `config`
fix-ups and their mutation semantics (ddpm.py:1080-1119), noise-tape generation in the reference's
draw order.  Device side (C ABI `ld_sample`): the whole T-step loop -- IND/OOD UNet launch, masked
x0, clamp, posterior update, one-off fusion composite -- with no host round trip per step.
"""
import ctypes as C

import torch
from torch import nn

from . import _lib
from .schedule import make_buffers

_NON_MRI = ("mnist", "mvtec", "oct", "imagenet")


class GaussianDiffusion(nn.Module):
    def __init__(self, config, model, *, image_size, timesteps=1000, sampling_timesteps=None, objective="pred_v",
                 beta_schedule="sigmoid", schedule_fn_kwargs=dict(), ddim_sampling_eta=0.0, auto_normalize=False,
                 offset_noise_strength=0.0, min_snr_loss_weight=False, min_snr_gamma=5):
        super().__init__()
        assert not (type(self) == GaussianDiffusion and model.channels != model.out_dim)  # ddpm.py:515
        assert not model.random_or_learned_sinusoidal_cond  # ddpm.py:516
        self.config = config  # shared by reference and mutated, like the reference (ddpm.py:518)
        self.branch_out = self.config["branch_out"]
        self.start_intermediate = self.config["start_intermediate"]
        self.model = model
        self.cnt = -1
        self.channels = model.channels
        self.self_condition = model.self_condition
        self.image_size = image_size
        self.objective = objective
        assert objective in {"pred_noise", "pred_x0", "pred_v"}, \
            "objective must be either pred_noise (predict noise) or pred_x0 (predict image start) or pred_v"
        bufs = make_buffers(beta_schedule, timesteps, objective, min_snr_loss_weight, min_snr_gamma, **schedule_fn_kwargs)
        for k, v in bufs.items():
            self.register_buffer(k, v)
        self.num_timesteps = int(timesteps)
        self.num_timesteps_ori = int(timesteps)
        self.sampling_timesteps = sampling_timesteps if sampling_timesteps is not None else timesteps
        assert self.sampling_timesteps <= timesteps  # ddpm.py:561
        self.is_ddim_sampling = self.sampling_timesteps < timesteps
        self.ddim_sampling_eta = ddim_sampling_eta
        if auto_normalize:
            raise NotImplementedError("auto_normalize=True is not used by the reference's drivers (test.py:138)")
        self._schedule_on = None
        self.last_x_start = None

    def call_classifier(self):  # ddpm.py:622-625
        if self.config.get("classifier", False):
            raise NotImplementedError("the PatchCore classifier gate is outside the sampler hot path (SURVEY.md §2)")

    @property
    def device(self):
        return self.betas.device

    # ---------------------------------------------------------------------------------------
    def _push_schedule(self, h):
        # keyed on the engine GENERATION (a re-created handle may re-use the freed address) and on the buffers' version counters
        # (a parent load_state_dict / in-place edit of the 13 buffers must reach the device)
        bufs = (self.posterior_mean_coef1, self.posterior_mean_coef2, self.posterior_log_variance_clipped)
        key = (self.model._engine_gen, str(self.device), tuple(b._version for b in bufs), tuple(b.data_ptr() for b in bufs))
        if self._schedule_on == key:
            return
        c1 = self.posterior_mean_coef1.detach().cpu().contiguous()
        c2 = self.posterior_mean_coef2.detach().cpu().contiguous()
        lv = self.posterior_log_variance_clipped.detach().cpu().contiguous()
        sg = (0.5 * lv).exp().contiguous()  # ddpm.py:853
        _lib.check(_lib.lib().ld_set_schedule(h, c1.numel(), c1.data_ptr(), c2.data_ptr(), lv.data_ptr(), sg.data_ptr()))
        # objective (ddpm.py:534-536): pred_x0 needs nothing; pred_noise / pred_v hand over the pair of per-timestep coefficients of
        # predict_start_from_noise / predict_start_from_v (ddpm.py:631-653)
        if self.objective == "pred_x0":
            _lib.check(_lib.lib().ld_set_objective(h, 0, None, None))
        else:
            names = {"pred_noise": ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod"),
                     "pred_v": ("sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod")}[self.objective]
            a, b = (getattr(self, n).detach().cpu().contiguous() for n in names)
            _lib.check(_lib.lib().ld_set_objective(h, a.numel(), a.data_ptr(), b.data_ptr()))
        self._schedule_on = key

    def make_noise_tape(self, shape, steps, device):
        """x_T then one draw per step for t = steps-1 .. 1, in the reference's order (ddpm.py:934-935, 852/857)."""
        torch.manual_seed(10)
        tape = torch.empty((steps,) + tuple(shape), device=device)
        for i in range(steps):
            tape[i] = torch.randn(shape, device=device)
        return tape

    # ---------------------------------------------------------------------------------------
    def ddim_schedule(self):
        """(times, coefs, fuse_step) of `ddim_sample` (ddpm.py:982-987, 1013-1018), built with the reference's own fp32
        tensor expressions so the coefficients are bit-identical.  coefs[i] = (sqrt_recip_alphas_cumprod[t],
        sqrt_recipm1_alphas_cumprod[t], sqrt(alpha_next), c, sigma); the last step has no successor."""
        T, Ssteps, eta = self.num_timesteps, self.sampling_timesteps, self.ddim_sampling_eta
        times = torch.linspace(-1, T - 1, steps=Ssteps + 1)
        times = list(reversed(times.int().tolist()))
        pairs = list(zip(times[:-1], times[1:]))
        self.start_timestep_ddim = times[-self.config["start_timestep"] - 2]
        ac = self.alphas_cumprod.detach().cpu()
        sr, srm1 = self.sqrt_recip_alphas_cumprod.detach().cpu(), self.sqrt_recipm1_alphas_cumprod.detach().cpu()
        coefs = torch.zeros(len(pairs), 5, dtype=torch.float32)
        for i, (t, tn) in enumerate(pairs):
            coefs[i, 0], coefs[i, 1] = sr[t], srm1[t]
            if tn < 0:
                continue
            alpha, alpha_next = ac[t], ac[tn]
            sigma = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
            c = (1 - alpha_next - sigma ** 2).sqrt()
            coefs[i, 2], coefs[i, 3], coefs[i, 4] = alpha_next.sqrt(), c, sigma
        fuse = -1
        if self.config["start_intermediate"]:
            for i, (t, _) in enumerate(pairs):
                if t <= self.start_timestep_ddim:  # ddpm.py:1022
                    fuse = i
                    break
        return [t for t, _ in pairs], coefs, fuse

    def _ddim_sample(self, h, dev, cond_img, mask, min_max_val, shape, noise):
        """ddpm.py:979-1075 on the device loop `ld_sample_ddim`.  The reference draws x_T and one `randn_like` per step
        (none for the last) from the global generator WITHOUT re-seeding; `noise=` replaces those draws."""
        cfg = self.config
        B, Cc, S, _ = shape
        times, coefs, fuse = self.ddim_schedule()
        n = len(times)
        if noise is None:
            noise = torch.stack([torch.randn(shape, device=dev) for _ in range(n)])
        noise = noise.to(dev, torch.float32).contiguous()
        assert noise.shape[0] >= n and tuple(noise.shape[1:]) == tuple(shape)
        cond = cond_img.to(dev, torch.float32).contiguous()
        assert tuple(cond.shape) == tuple(shape)
        mk = mask.to(dev, torch.float32).contiguous() if mask is not None else None
        sd = _lib.SampleDesc()
        sd.batch, sd.height, sd.width, sd.num_timesteps = B, S, S, n
        sd.branch_out = int(bool(cfg["branch_out"]))
        sd.start_intermediate = int(bool(cfg["start_intermediate"]))
        sd.start_timestep = int(cfg["start_timestep"])
        sd.mask_x = int(bool(cfg["mask_x"]))
        data = cfg["data"]
        sd.ood_uses_cond = int(any(s in data for s in _NON_MRI) and "mri" not in data)
        sd.cond_in_floor = 0.5 if data == "mnist" else 0.95
        sd.min_val, sd.max_val = float(min_max_val[0]), float(min_max_val[1])
        will_fuse = bool(sd.branch_out) and 0 <= fuse < n - 1
        pair = bool(sd.branch_out) and not will_fuse  # the reference returns the list [x_out, x_in] (ddpm.py:1010, 1073)
        sd.return_pair = int(pair)
        out = torch.empty((2,) + tuple(shape) if pair else tuple(shape), device=dev)
        tt = torch.tensor(times, dtype=torch.int32)
        rc = _lib.lib().ld_sample_ddim(h, C.byref(sd), cond.data_ptr(), mk.data_ptr() if mk is not None else None,
                                       noise.data_ptr(), out.data_ptr(), tt.data_ptr(), coefs.data_ptr(), n,
                                       fuse if cfg["start_intermediate"] else -1,
                                       C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        if will_fuse:  # ddpm.py:1023-1024
            cfg["branch_out"] = False
            cfg["mask_x"] = False
        _lib.check(rc)
        return [out[0], out[1]] if pair else out

    @torch.inference_mode()
    def sample(self, cond_img, gt, batch_size=16, return_all_timesteps=False, return_all_outputs=False, mask=None,
               ood_confidence_ad=False, min_max_val=None, instance=0, noise=None):
        """Same signature as the reference (ddpm.py:1078) plus `noise=`: an optional host- or
        device-resident tape `[T, B, C, S, S]` (x_T first) replacing the seed-10 device draws."""
        cfg = self.config
        self.instance = instance
        self.cnt += 1
        self.min_max_val = min_max_val
        # --- per-call flag fix-ups, ddpm.py:1093-1117 ---
        if cfg["branch_out"] == False:  # noqa: E712 (mirrors the reference's comparison)
            cfg["branch_out"] = self.branch_out
        if cfg["start_intermediate"] == False:  # noqa: E712
            cfg["start_intermediate"] = self.start_intermediate
        self.start_intermediate = bool(cfg["start_intermediate"])
        if cfg["ood_AD"] or cfg["ood_confidence"]:
            cfg["mask_cond"] = True
            cfg["mask_x"] = True
        if cfg["branch_out"]:
            u = torch.unique(mask)
            if len(u) == 1 and u == 1:
                cfg["mask_cond"] = cfg["mask_x"] = cfg["branch_out"] = cfg["start_intermediate"] = False
        if cfg["branch_out"] and self.objective != "pred_x0":
            # the reference's branch path only defines `model_output_*` (ddpm.py:694-695): the other objectives hit an unbound name
            raise UnboundLocalError("cannot access local variable 'model_output' where it is not associated with a value "
                                    "(branch sampling needs objective='pred_x0', ddpm.py:731-761)")
        if return_all_timesteps and cfg["branch_out"]:
            # `torch.stack(imgs, dim=1)` over a list that holds [out, in] pairs (ddpm.py:865, 964)
            raise TypeError("expected Tensor as element 1 in argument 0, but got list (return_all_timesteps with branch_out, ddpm.py:964)")
        if return_all_timesteps and self.is_ddim_sampling:
            raise NotImplementedError("return_all_timesteps is supported on the DDPM path only")
        if cfg.get("classifier", False) and cfg["start_intermediate"]:
            raise NotImplementedError("classifier gate is out of scope (ddpm.py:883-916)")

        h = self.model.engine()
        dev = self.model._handle_device
        self._push_schedule(h)
        B, Cc, S = batch_size, self.channels, self.image_size
        if self.is_ddim_sampling:  # ddpm.py:1122-1124: return_all_outputs is not forwarded to ddim_sample
            return self._ddim_sample(h, dev, cond_img, mask, min_max_val, (B, Cc, S, S), noise)
        steps = self.num_timesteps
        use_gt = bool(self.start_intermediate and cfg.get("use_gt", False))
        if use_gt:
            steps = int(cfg["use_gt_timestep"])
        if noise is None:
            noise = self.make_noise_tape((B, Cc, S, S), steps, dev)
        noise = noise.to(dev, torch.float32)
        assert noise.shape[0] >= steps and tuple(noise.shape[1:]) == (B, Cc, S, S)
        if use_gt:  # ddpm.py:937-944: start from q_sample(gt, t)
            tg = int(cfg["use_gt_timestep"])
            noise = noise.clone()
            noise[0] = self.sqrt_alphas_cumprod[tg] * gt.to(dev) + self.sqrt_one_minus_alphas_cumprod[tg] * noise[0]
            self.num_timesteps = tg
        noise = noise.contiguous()
        cond = cond_img.to(dev, torch.float32).contiguous()
        assert tuple(cond.shape) == (B, Cc, S, S)
        mk = mask.to(dev, torch.float32).contiguous() if mask is not None else None

        sd = _lib.SampleDesc()
        sd.batch, sd.height, sd.width, sd.num_timesteps = B, S, S, steps
        sd.branch_out = int(bool(cfg["branch_out"]))
        sd.start_intermediate = int(bool(cfg["start_intermediate"]))
        sd.start_timestep = int(cfg["start_timestep"])
        sd.mask_x = int(bool(cfg["mask_x"]))
        data = cfg["data"]
        sd.ood_uses_cond = int(any(s in data for s in _NON_MRI) and "mri" not in data)  # ddpm.py:704-708
        sd.cond_in_floor = 0.5 if data == "mnist" else 0.95  # ddpm.py:683-686
        sd.min_val, sd.max_val = float(min_max_val[0]), float(min_max_val[1])
        pair = (not self.start_intermediate) and bool(self.branch_out)  # ddpm.py:965-970
        sd.return_pair = int(pair)
        sd.record_x0 = int(return_all_outputs or return_all_timesteps)
        out = torch.empty((2, B, Cc, S, S) if pair else (B, Cc, S, S), device=dev)
        trace = torch.zeros((steps, 2, B, Cc, S, S), device=dev) if sd.record_x0 else None
        will_fuse = sd.branch_out and sd.start_intermediate
        rc = _lib.lib().ld_sample(h, C.byref(sd), cond.data_ptr(), mk.data_ptr() if mk is not None else None,
                                  noise.data_ptr(), out.data_ptr(), trace.data_ptr() if trace is not None else None,
                                  C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        if will_fuse:  # the fusion step flips these for the rest of the call (ddpm.py:780-781)
            cfg["branch_out"] = False
            cfg["mask_x"] = False
        _lib.check(rc)
        if return_all_timesteps:  # ddpm.py:946, 964: imgs = [x_T, x_{T-1}, ..., x_0] stacked along dim 1 (single trajectory only)
            out = torch.cat((noise[0].unsqueeze(1), trace[:, 1].permute(1, 0, 2, 3, 4)), dim=1)
            if pair:  # ddpm.py:965-970 stacks whatever `ret` is
                out = torch.stack((out, out), dim=0)
        if return_all_outputs:  # ddpm.py:973-974: (ret, x_start_lst, confidence_map)
            lst = []
            branched = bool(sd.branch_out)
            for i in range(steps):
                t = steps - 1 - i
                if branched and not (sd.start_intermediate and t <= sd.start_timestep):
                    lst.append([trace[i, 0].cpu(), trace[i, 1].cpu()])
                else:
                    branched = False
                    lst.append(trace[i, 0].cpu())
            return out, lst, []
        return out
