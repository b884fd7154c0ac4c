"""ctypes binding of the C ABI declared in `include/ld_sampler.h` (libld_sampler.so, in-tree).

There is deliberately no fallback: if the shared library is missing the import of any compute
entry point raises, and every compute call fails with LD_ERR_NO_DEVICE on a machine without an
sm_100 GPU.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# LD_SAMPLER_LIB: development override (A/B builds of the same ABI); the product path is the in-tree library
LIB_PATH = os.path.abspath(os.environ["LD_SAMPLER_LIB"]) if os.environ.get("LD_SAMPLER_LIB") else os.path.join(_HERE, "libld_sampler.so")
LD_MAX_LEVELS = 8

LD_OK, LD_ERR_INVALID, LD_ERR_NO_DEVICE, LD_ERR_CUDA, LD_ERR_STATE, LD_ERR_KEY, LD_ERR_MASK = 0, -1, -2, -3, -4, -5, -6


class ModelDesc(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("init_dim", C.c_int32), ("n_levels", C.c_int32),
        ("dim_mults", C.c_int32 * LD_MAX_LEVELS), ("full_attn", C.c_int32 * LD_MAX_LEVELS),
        ("channels", C.c_int32), ("resnet_groups", C.c_int32), ("attn_heads", C.c_int32),
        ("attn_dim_head", C.c_int32), ("sinusoidal_theta", C.c_float), ("cond_mode", C.c_int32),
        ("precision", C.c_int32),
    ]


class SampleDesc(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("num_timesteps", C.c_int32),
        ("branch_out", C.c_int32), ("start_intermediate", C.c_int32), ("start_timestep", C.c_int32),
        ("mask_x", C.c_int32), ("ood_uses_cond", C.c_int32), ("cond_in_floor", C.c_float),
        ("min_val", C.c_float), ("max_val", C.c_float), ("return_pair", C.c_int32), ("record_x0", C.c_int32),
    ]


# name -> (restype, argtypes); must list every symbol of include/ld_sampler.h
SIGNATURES = {
    "ld_last_error": (C.c_char_p, []),
    "ld_version": (C.c_char_p, []),
    "ld_device_count": (C.c_int, []),
    "ld_create": (C.c_int, [C.POINTER(ModelDesc), C.c_int, C.POINTER(C.c_void_p)]),
    "ld_destroy": (C.c_int, [C.c_void_p]),
    "ld_num_weights": (C.c_int, [C.c_void_p]),
    "ld_weight_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    "ld_load_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int]),
    "ld_finalize_weights": (C.c_int, [C.c_void_p]),
    "ld_set_schedule": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ld_set_objective": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ld_unet_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ld_cond_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ld_sample": (C.c_int, [C.c_void_p, C.POINTER(SampleDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ld_sample_finish": (C.c_int, [C.c_void_p]),
    "ld_posterior_step": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.POINTER(SampleDesc), C.c_int64, C.c_void_p]),
    "ld_prep_mnist": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "ld_prep_mri": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p]),
    "ld_mask_scratch_bytes": (C.c_int64, [C.c_int, C.c_int]),
    "ld_mask_from_anomaly": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "ld_knn_scratch_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "ld_knn_min": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ld_launch_count": (C.c_int64, [C.c_void_p]),
    "ld_workspace_bytes": (C.c_int64, [C.c_void_p]),
    "ld_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "ld_get_option": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64)]),
    "ld_debug_num_taps": (C.c_int, [C.c_void_p]),
    "ld_debug_tap_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int32)]),
    "ld_debug_tap_fetch": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ld_debug_conv_time": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(C.c_float), C.c_void_p]),
    "ld_debug_conv_variant_time": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                                             C.c_void_p]),
    "ld_debug_conv_fused": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_void_p]),
    "ld_sample_ddim": (C.c_int, [C.c_void_p, C.POINTER(SampleDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_int, C.c_int, C.c_void_p]),
    "ld_debug_conv_dual": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ld_debug_linattn": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    "ld_debug_linattn_h": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "ld_debug_attention": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ld_debug_conv7": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ld_debug_conv": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None


class LdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ld_sampler error {code}: {msg}")
        self.code = code


def lib():
    """Load libld_sampler.so (once).  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / PyTorch fallback for the sampler)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(l, name)
            f.restype, f.argtypes = res, args
        _lib = l
    return _lib


def check(code):
    """Map a negative ld_status to the exception type the reference would have raised."""
    if code >= 0:
        return code
    msg = lib().ld_last_error().decode()
    if code == LD_ERR_MASK or (code == LD_ERR_INVALID and "divisible" in msg):
        raise AssertionError(msg)  # ddpm.py:405, 698, 790 are asserts
    raise LdError(code, msg)
