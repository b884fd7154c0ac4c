"""`Unet` -- the reference's denoiser API (ddpm.py:286-451) hosted on the sm_100a engine.

Same constructor arguments, attributes (`channels`, `out_dim`, `self_condition`,
`random_or_learned_sinusoidal_cond`, `downsample_factor`) and `state_dict()` layout as the
reference class, so `load_state_dict(reference_sd, strict=True)` works.  `forward` runs entirely in
hand-written CUDA through the C ABI (`ld_unet_forward`); there is no PyTorch compute path.
"""
import ctypes as C
import math
import os

import torch
from torch import nn

from . import _lib
from .spec import TOP_LEVEL_ORDER, cond_is_deep, cond_returns_early, param_specs

_PRECISIONS = {"fp32": 0, "bf16": 1}


class _Node(nn.Module):
    """Bare parameter container; children are created from dotted state_dict keys."""

    def child(self, name):
        if name not in self._modules:
            self.add_module(name, _Node())
        return self._modules[name]


def _tuple(v, n):
    return tuple(v) if isinstance(v, (tuple, list)) else (v,) * n


class Unet(nn.Module):
    def __init__(self, dim, init_dim=None, out_dim=None, dim_mults=(1, 2, 4, 8), channels=1, self_condition=False,
                 cond_img=True, resnet_block_groups=8, learned_variance=False, learned_sinusoidal_cond=False,
                 random_fourier_features=False, learned_sinusoidal_dim=16, sinusoidal_pos_emb_theta=10000,
                 attn_dim_head=32, attn_heads=4, full_attn=(False, False, False, True), flash_attn=False, mode="mri",
                 precision=None):
        super().__init__()
        if self_condition:
            raise NotImplementedError("self_condition is unreachable in the reference's sampling configs (ddpm.py:294)")
        if learned_sinusoidal_cond or random_fourier_features:
            raise NotImplementedError("GaussianDiffusion asserts these off (ddpm.py:516)")
        if learned_variance:
            raise NotImplementedError("learned_variance is rejected by GaussianDiffusion (ddpm.py:515)")
        L = len(dim_mults)
        heads, dh, fa = _tuple(attn_heads, L), _tuple(attn_dim_head, L), _tuple(full_attn, L)
        assert len(fa) == L  # ddpm.py:353
        if len(set(heads)) != 1 or len(set(dh)) != 1:
            raise NotImplementedError("per-level attn_heads / attn_dim_head are not supported")
        self.mode = mode
        self.channels = channels
        self.self_condition = False
        self.cond_img = cond_img
        self.random_or_learned_sinusoidal_cond = False
        self.dim = dim
        self.init_dim = init_dim if init_dim is not None else dim
        self.dim_mults = tuple(dim_mults)
        self.full_attn = tuple(bool(f) for f in fa)
        self.attn_heads, self.attn_dim_head = heads[0], dh[0]
        self.resnet_block_groups = resnet_block_groups
        self.theta = sinusoidal_pos_emb_theta
        self.out_dim = out_dim if out_dim is not None else channels
        self.flash_attn = flash_attn  # accepted for API parity; the engine has its own attention kernels
        self.precision = precision or os.environ.get("LD_PRECISION", "bf16")
        if self.precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
        if cond_is_deep(mode) == cond_returns_early(mode):
            raise NotImplementedError(f"mode {mode!r}: the reference's ResUnet has no usable forward for it")

        for name in TOP_LEVEL_ORDER:  # registration order == state_dict key order of the reference
            self.add_module(name, _Node())
        self._specs = param_specs(dim, self.init_dim, self.dim_mults, channels, self.attn_heads, self.attn_dim_head,
                                  self.full_attn, mode, self.out_dim)
        for sp in self._specs:  # construction order == RNG consumption order of the reference
            *path, leaf = sp.key.split(".")
            node = self
            for part in path:
                node = node.child(part) if isinstance(node, _Node) else node._modules[part]
            p = nn.Parameter(torch.empty(sp.shape))
            with torch.no_grad():
                if sp.init == "kaiming":
                    nn.init.kaiming_uniform_(p, a=math.sqrt(5))
                elif sp.init == "bias":
                    b = 1 / math.sqrt(sp.fan_in) if sp.fan_in > 0 else 0
                    nn.init.uniform_(p, -b, b)
                elif sp.init == "ones":
                    p.fill_(1.0)
                else:
                    p.zero_()
            node.register_parameter(leaf, p)
        self._handle = None
        self._handle_device = None
        self._engine_gen = 0        # bumped whenever a native engine is created (consumers re-push per-engine state)
        self._param_version = None  # sum of the parameters' autograd version counters when the weights were last packed
        # a parent's load_state_dict (the reference's `diffusion.load_state_dict(data['model'])`, ddpm.py:1517, or EMA copies)
        # never calls a child's load_state_dict override, but it does run the child's post hooks: invalidate the engine there
        self.register_load_state_dict_post_hook(lambda module, incompatible_keys: module.release_engine())

    # -- reference API -----------------------------------------------------------------------
    @property
    def downsample_factor(self):
        return 2 ** (len(self.dim_mults) - 1)

    def forward(self, x, cond_img, time, x_self_cond=None):
        d = self.downsample_factor
        assert all(s % d == 0 for s in x.shape[-2:]), \
            f"your input dimensions {tuple(x.shape[-2:])} need to be divisible by {d}, given the unet"  # ddpm.py:405
        h = self.engine()
        dev = self._handle_device
        x = x.to(dev, torch.float32).contiguous()
        cond = cond_img.to(dev, torch.float32).contiguous()
        t = time.to(dev, torch.int64).contiguous()
        n, c, hh, ww = x.shape
        assert c == 1 and cond.shape == x.shape and t.shape == (n,)
        out = torch.empty_like(x)
        _lib.check(_lib.lib().ld_unet_forward(h, x.data_ptr(), cond.data_ptr(), t.data_ptr(), out.data_ptr(), n, hh, ww,
                                              C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        return out

    def encode_condition(self, cond_img):
        """`self.cond_model(cond_img)` (ddpm.py:434) -> NCHW fp32 features, for parity tests."""
        h = self.engine()
        dev = self._handle_device
        cond = cond_img.to(dev, torch.float32).contiguous()
        n, _, hh, ww = cond.shape
        f = 8 if cond_is_deep(self.mode) else 4
        cf = 256 if cond_is_deep(self.mode) else 128
        out = torch.empty(n, cf, hh // f, ww // f, device=dev)
        _lib.check(_lib.lib().ld_cond_encode(h, cond.data_ptr(), out.data_ptr(), n, hh, ww,
                                             C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        return out

    # -- engine management ----------------------------------------------------------------------
    def release_engine(self):
        if self._handle is not None:
            _lib.lib().ld_destroy(self._handle)
        self._handle = None
        self._handle_device = None

    def __del__(self):
        try:
            self.release_engine()
        except Exception:
            pass

    def model_desc(self):
        d = _lib.ModelDesc()
        d.dim, d.init_dim, d.n_levels = self.dim, self.init_dim, len(self.dim_mults)
        for i, (m, f) in enumerate(zip(self.dim_mults, self.full_attn)):
            d.dim_mults[i], d.full_attn[i] = m, int(f)
        d.channels, d.resnet_groups = self.channels, self.resnet_block_groups
        d.attn_heads, d.attn_dim_head = self.attn_heads, self.attn_dim_head
        d.sinusoidal_theta = float(self.theta)
        d.cond_mode = 0 if cond_is_deep(self.mode) else 1
        d.precision = _PRECISIONS[self.precision]
        return d

    def engine(self, options=None):
        """Create (once) the native handle on the device the parameters live on and push the weights."""
        lib = _lib.lib()
        dev = next(self.parameters()).device
        # in-place updates through autograd-visible ops (`p.copy_`, `p.mul_`, optimiser / EMA steps) bump the version counters;
        # writes through `p.data` do not -- call release_engine() after those
        pv = sum(p._version for p in self.parameters())
        if self._handle is not None and self._handle_device == dev and self._param_version == pv:
            return self._handle
        self.release_engine()
        if dev.type != "cuda":
            # no CPU path: let the library produce its own loud error (LD_ERR_NO_DEVICE) or refuse here
            if lib.ld_device_count() == 0:
                raise _lib.LdError(_lib.LD_ERR_NO_DEVICE, "no sm_100 CUDA device is available (there is no CPU fallback)")
            raise RuntimeError("move the model to a CUDA device first (`.cuda()`); there is no CPU fallback")
        if self.out_dim != self.channels or self.channels != 1:
            raise NotImplementedError("the engine supports channels == out_dim == 1 (the reference's configs)")
        h = C.c_void_p()
        desc = self.model_desc()
        _lib.check(lib.ld_create(C.byref(desc), dev.index if dev.index is not None else torch.cuda.current_device(), C.byref(h)))
        try:
            for k, v in (options or getattr(self, "_engine_options", {}) or {}).items():
                _lib.check(lib.ld_set_option(h, k.encode(), int(v)))
            sd = self.state_dict()
            for sp in self._specs:
                w = sd[sp.key].detach().to("cpu", torch.float32).contiguous()
                shape = (C.c_int64 * w.dim())(*w.shape)
                _lib.check(lib.ld_load_weight(h, sp.key.encode(), w.data_ptr(), shape, w.dim()))
            _lib.check(lib.ld_finalize_weights(h))
        except Exception:
            lib.ld_destroy(h)
            raise
        self._handle, self._handle_device = h, dev
        self._param_version = pv
        self._engine_gen += 1
        return h

    def set_engine_options(self, **opts):
        """Tunables forwarded to `ld_set_option` (takes effect when the engine is next created)."""
        self._engine_options = dict(opts)
        self.release_engine()

    def launch_count(self):
        return int(_lib.lib().ld_launch_count(self._handle)) if self._handle is not None else 0
