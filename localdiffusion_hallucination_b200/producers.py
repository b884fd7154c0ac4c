"""The two stages in front of the sampler, on the device: conditional-image producers
(`data.py:814-836`, `data.py:380-414`, `test.py:17-37`) and the anomaly-map -> mask stage of the inference script
(`test.py:237-381`).  Thin ctypes wrappers over `ld_prep_mnist`, `ld_prep_mri`, `ld_mask_from_anomaly`; there is no CPU path."""
import ctypes as C

import torch

from . import _lib

MASK_RULES = {"mnist_8to3": 0, "mnist_8to5": 1, "mri_t12flair": 2, "mri_flair2t1": 3,
              "mvtec_transistor": 4, "mvtec_toothbrush": 5, "mvtec_grid": 6}


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _need_cuda(t):
    if t.device.type != "cuda":
        if _lib.lib().ld_device_count() == 0:
            raise _lib.LdError(_lib.LD_ERR_NO_DEVICE, "no sm_100 CUDA device is available (there is no CPU fallback)")
        raise RuntimeError("move the input to a CUDA device first; there is no CPU fallback")


def mnist_pair(raw):
    """`MNIST.__getitem__` (data.py:814-836): raw [N,S,S] (0..255) on the device -> (hr, cond), each [N,1,S,S] in [0,2]."""
    _need_cuda(raw)
    x = raw.to(torch.float32).contiguous()
    n, s, s2 = x.shape
    assert s == s2, "the reference up-samples to (W, W) (data.py:826): square images only"
    hr, cond = torch.empty(n, 1, s, s, device=x.device), torch.empty(n, 1, s, s, device=x.device)
    _lib.check(_lib.lib().ld_prep_mnist(x.data_ptr(), hr.data_ptr(), cond.data_ptr(), n, s, _stream(x.device)))
    return hr, cond


def mri_normalize(raw, mean, std, translate_zero=True, crop=224):
    """`MedDataset_png.transform` + `.normalize` (data.py:380-414): raw [N,Hs,Ws] -> [N,1,crop,crop]."""
    _need_cuda(raw)
    x = raw.to(torch.float32).contiguous()
    n, hs, ws = x.shape
    out = torch.empty(n, 1, crop, crop, device=x.device)
    scratch = torch.empty(max(n, 1), dtype=torch.int32, device=x.device)
    _lib.check(_lib.lib().ld_prep_mri(x.data_ptr(), out.data_ptr(), scratch.data_ptr(), n, hs, ws, crop, float(mean), float(std),
                                      int(bool(translate_zero)), _stream(x.device)))
    return out


def set_min_max_val(config, mode="mri"):
    """test.py:17-37: the clamp range `sample()` receives, in the reference's order (max, min[, min_t1]).  With translate_zero the
    reference adds an fp32 0-dim tensor, i.e. the maximum is rounded to fp32; reproduced with the same torch expression."""
    if mode == "mri":
        if not config["translate_zero"]:
            max_val = (4096 - config["mean_flair"]) / config["std_flair"]
            min_val = (0 - config["mean_flair"]) / config["std_flair"]
            min_val_t1 = (0 - config["mean_t1"]) / config["std_t1"]
        else:
            min_val2 = (0 - config["mean_flair"]) / config["std_flair"]
            min_val = 0.0
            max_val = (4096 - config["mean_flair"]) / config["std_flair"]
            max_val = float(max_val + torch.abs(torch.tensor(min_val2)))
            min_val_t1 = 0.0
        return max_val, min_val, min_val_t1
    if mode in ("mnist", "mvtec"):
        return 2.0, 0.0
    raise ValueError(mode)


def masks_from_anomaly(anomaly_map, rule, img_size=None, manual_cols=0, want_binary=True):
    """test.py:237-381: anomaly map [B,1,h,w] on the device -> (mask_pred, binary_mask) [B,1,S,S].  `rule` names the dataset's
    threshold table; `img_size` triggers the bilinear resize of the mnist / mvtec branch (test.py:246-247); `manual_cols = 7`
    reproduces the manual mask the shipped script substitutes (test.py:379-381)."""
    _need_cuda(anomaly_map)
    a = anomaly_map.to(torch.float32).contiguous()
    b, c, h, w = a.shape
    assert c == 1
    s = img_size if img_size is not None else h
    if img_size is None:
        assert h == w, "without a resize the map must already be square"
    lib = _lib.lib()
    mp = torch.empty(b, 1, s, s, device=a.device)
    bm = torch.empty(b, 1, s, s, device=a.device) if want_binary else None
    scratch = torch.empty(int(lib.ld_mask_scratch_bytes(b, s)), dtype=torch.uint8, device=a.device)
    _lib.check(lib.ld_mask_from_anomaly(a.data_ptr(), b, h, w, s, MASK_RULES[rule], int(manual_cols), mp.data_ptr(),
                                        bm.data_ptr() if bm is not None else None, scratch.data_ptr(), _stream(a.device)))
    return mp, bm


def knn_min(embedding, memory_bank):
    """PatchCore `nearest_neighbors(embedding, n_neighbors=1)` (models.py:179-217) on the tensor cores: embedding [M,D] and
    memory_bank [Nb,D] on the device -> (patch_scores [M], locations [M] int64).  The [M x Nb] distance matrix is never formed."""
    _need_cuda(embedding)
    x = embedding.to(torch.float32).contiguous()
    y = memory_bank.to(x.device, torch.float32).contiguous()
    m, d = x.shape
    nb, d2 = y.shape
    assert d == d2
    lib = _lib.lib()
    score = torch.empty(m, device=x.device)
    loc = torch.empty(m, dtype=torch.int64, device=x.device)
    scratch = torch.empty(int(lib.ld_knn_scratch_bytes(m, nb, d)), dtype=torch.uint8, device=x.device)
    _lib.check(lib.ld_knn_min(x.data_ptr(), y.data_ptr(), m, nb, d, score.data_ptr(), loc.data_ptr(), scratch.data_ptr(), _stream(x.device)))
    return score, loc
