"""TEST INFRASTRUCTURE ONLY -- imports the *real* reference (`/root/reference`) on CPU.

This module only works in the build container (the reference tree does not travel to the
GPU box).  It is used by `tests/golden/make_golden.py` to (1) validate the CPU restatement in
`oracle/ld_oracle.py` against the reference's own classes and (2) generate the golden
fixtures committed under `tests/golden/`.  Nothing under the product package imports it.

The import recipe follows SURVEY.md Appendix A: the reference's `ddpm.py` imports several
packages that are not installed here (`ddpm.py:25-46`) and three in-repo symbols that do not
exist (`ddpm.py:30,46`), so stub modules are registered before the import.  No reference file
is modified or copied.
"""
import importlib.machinery
import os
import sys
import types

REF_ROOT = os.environ.get("LD_REFERENCE_ROOT", "/root/reference")


class _Any:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Any()

    def __getattr__(self, n):
        return _Any()


def _stub(name):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []

    def _ga(n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Any

    m.__getattr__ = _ga
    sys.modules[name] = m
    return m


_STUBS = [
    "ema_pytorch", "accelerate", "idx2numpy", "timm", "nibabel", "medpy", "medpy.io",
    "anomalib", "anomalib.models", "anomalib.models.components", "anomalib.models.patchcore",
    "anomalib.models.patchcore.anomaly_map", "anomalib.pre_processing", "train_fusion",
    "datasets", "datasets.utils", "datasets.utils.file_utils",
]

_ddpm = None


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "ddpm.py"))


def load_reference():
    """Return the reference's `ddpm` module (cached)."""
    global _ddpm
    if _ddpm is not None:
        return _ddpm
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    sys.dont_write_bytecode = True
    for n in _STUBS:
        if n not in sys.modules:
            _stub(n)
    import torch

    class _DBM(torch.nn.Module):
        pass

    sys.modules["anomalib.models.components"].DynamicBufferModule = _DBM
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import data as _d  # noqa

    _d.OCTID = _Any
    _d.ImageNetDatasetSR = _Any
    import ddpm  # noqa

    if not torch.cuda.is_available():
        # ddpm.py:743,748,754 hard-code .cuda() on the pred_x0 path
        torch.Tensor.cuda = lambda self, *a, **k: self
    _ddpm = ddpm
    return ddpm


class noise_tape:
    """Context manager: make the reference consume a pre-generated host noise tape.

    The reference draws `torch.randn(shape)` once (`ddpm.py:935`) and `torch.randn_like`
    once per step for t = T-1 .. 1 (`ddpm.py:852,857`).  `tape[0]` is x_T, `tape[1+i]` is the
    i-th per-step draw.
    """

    def __init__(self, tape):
        self.tape = tape
        self.i = 0

    def __enter__(self):
        import torch

        self._randn, self._randn_like, self._seed = torch.randn, torch.randn_like, torch.manual_seed

        def _next(*a, **k):
            t = self.tape[self.i]
            self.i += 1
            return t.clone()

        torch.randn = _next
        torch.randn_like = _next
        return self

    def __exit__(self, *exc):
        import torch

        torch.randn, torch.randn_like = self._randn, self._randn_like
        return False
