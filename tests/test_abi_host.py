"""CPU: C-ABI library loads and exports everything the header declares; host logic; loud no-GPU errors."""
import ctypes as C
import os
import re

import pytest
import torch

from localdiffusion_hallucination_b200 import GaussianDiffusion, Unet, _lib
from localdiffusion_hallucination_b200.schedule import make_buffers
from oracle import ld_oracle as lo
from tests import util
from tests.golden import cases

HEADER = os.path.join(util.ROOT, "include", "ld_sampler.h")


def header_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"LD_API\s+[\w\s\*]+?\b(ld_\w+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    lib = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(_lib.SIGNATURES) == syms
    assert lib.ld_version().decode().endswith("sm_100a")


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.ModelDesc) == 4 * (3 + 8 + 8 + 7)
    assert C.sizeof(_lib.SampleDesc) == 4 * 14


@pytest.mark.parametrize("name", list(cases.MODEL_KW))
def test_native_weight_registry_matches_python_state_dict(name):
    lib = _lib.lib()
    m = Unet(**cases.MODEL_KW[name])
    h = C.c_void_p()
    d = m.model_desc()
    assert lib.ld_create(C.byref(d), 0, C.byref(h)) == 0
    try:
        n = lib.ld_num_weights(h)
        native = {}
        for i in range(n):
            key, shape, nd = C.c_char_p(), (C.c_int64 * 4)(), C.c_int()
            assert lib.ld_weight_info(h, i, C.byref(key), shape, C.byref(nd)) == 0
            native[key.value.decode()] = tuple(shape[: nd.value])
        sd = m.state_dict()
        assert set(native) == set(sd)
        for k, v in sd.items():
            assert native[k] == tuple(v.shape), k
        # unknown / mis-shaped keys are rejected
        buf = torch.zeros(4)
        sh = (C.c_int64 * 1)(4)
        assert lib.ld_load_weight(h, b"not.a.key", buf.data_ptr(), sh, 1) == _lib.LD_ERR_KEY
        assert lib.ld_load_weight(h, b"init_conv.bias", buf.data_ptr(), sh, 1) == _lib.LD_ERR_KEY
        assert b"size mismatch" in lib.ld_last_error()
    finally:
        lib.ld_destroy(h)


def test_state_dict_layout_counts():
    # 334 Unet tensors / 12,140,481 params (mri); 347 keys with the 13 buffers
    m = Unet(**cases.MODEL_KW["mri"])
    sd = m.state_dict()
    assert len(sd) == 334 and sum(v.numel() for v in sd.values()) == 12140481
    assert "conv_fusion.mlp.1.weight" in sd  # dead but present (ddpm.py:436)
    gd = GaussianDiffusion(cases.base_config(), m, image_size=64, timesteps=10, objective="pred_x0")
    full = gd.state_dict()
    assert len(full) == 347 and all(k.startswith("model.") or k in make_buffers("sigmoid", 10, "pred_x0") for k in full)
    m2 = Unet(**cases.MODEL_KW["mri"])
    m2.load_state_dict(sd, strict=True)
    m3 = Unet(**cases.MODEL_KW["mnist"])
    assert len(m3.state_dict()) == 264 and sum(v.numel() for v in m3.state_dict().values()) == 3353473


def test_product_schedule_equals_oracle_schedule():
    for sched, T in (("sigmoid", 1000), ("linear", 100), ("cosine", 64)):
        a, b = make_buffers(sched, T, "pred_x0"), lo.diffusion_buffers(sched, T)
        for k in a:
            assert torch.equal(a[k], b[k]), (sched, T, k)
    with pytest.raises(ValueError):
        make_buffers("nope", 10, "pred_x0")


def test_invalid_model_descriptions_are_rejected():
    lib = _lib.lib()
    m = Unet(**cases.MODEL_KW["mri"])
    d = m.model_desc()
    d.attn_dim_head = 64
    h = C.c_void_p()
    assert lib.ld_create(C.byref(d), 0, C.byref(h)) == _lib.LD_ERR_INVALID
    d = m.model_desc()
    d.dim_mults[3] = 4  # bottleneck would be 128 channels, cond encoder emits 256 (ddpm.py:380)
    assert lib.ld_create(C.byref(d), 0, C.byref(h)) == _lib.LD_ERR_INVALID


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_gpu_is_a_loud_error_not_a_fallback():
    m = Unet(**cases.MODEL_KW["mnist"])
    x = torch.zeros(1, 1, 32, 32)
    with pytest.raises(_lib.LdError) as e:
        m(x, x, torch.zeros(1, dtype=torch.long))
    assert e.value.code == _lib.LD_ERR_NO_DEVICE
    gd = GaussianDiffusion(cases.base_config("mri"), m, image_size=32, timesteps=4, objective="pred_x0")
    with pytest.raises(_lib.LdError):
        gd.sample(x, None, batch_size=1, mask=cases.mask_left_columns(1, 32), min_max_val=(0.0, 2.0))
    assert _lib.lib().ld_device_count() == 0


def test_constructor_contract():
    m = Unet(**cases.MODEL_KW["mnist"])
    assert m.channels == 1 and m.out_dim == 1 and m.self_condition is False and m.downsample_factor == 4
    with pytest.raises(AssertionError):
        GaussianDiffusion(cases.base_config(), m, image_size=32, timesteps=10, objective="nope")
    with pytest.raises(AssertionError):
        GaussianDiffusion(cases.base_config(), m, image_size=32, timesteps=10, sampling_timesteps=11, objective="pred_x0")
    with pytest.raises(AssertionError):  # ddpm.py:405
        m(torch.zeros(1, 1, 30, 30), torch.zeros(1, 1, 30, 30), torch.zeros(1, dtype=torch.long))
