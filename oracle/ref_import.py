"""Import the UNMODIFIED reference modules (Module.py, Loss.py, ssim.py) from /root/reference.

TEST INFRASTRUCTURE ONLY.  Works only where /root/reference is mounted (the build container); it is used by
oracle/make_golden.py to generate tests/golden/*.pt and by tests that pin the oracle port against the real
reference.  Nothing on the product path, in `-m gpu` tests, smoke() or bench.py may import this.

Two shims:
  1. `osgeo` stub package (GDAL is not installed; CommonFunc.py:17-19 imports it at module scope);
  2. `Loss.vgg16` rebound to a random-init torchvision VGG16 (Loss.py:25 downloads ImageNet weights, and there
     is no network).  The perception term is out of scope (weight 0 in every parity run).
"""
import os
import sys

REFERENCE_DIR = os.environ.get("FCD_REFERENCE_DIR", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_shim")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "Module.py"))


def load():
    """Returns (Module, Loss, ssim) reference python modules."""
    if not available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_DIR}")
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    if REFERENCE_DIR not in sys.path:
        sys.path.append(REFERENCE_DIR)
    import torch
    import torchvision

    import ssim as ref_ssim  # noqa
    import Module as ref_module  # noqa
    import Loss as ref_loss  # noqa

    def _vgg16(pretrained=True):
        g = torch.random.get_rng_state()
        torch.manual_seed(1234)
        m = torchvision.models.vgg16(weights=None)
        torch.random.set_rng_state(g)
        return m

    ref_loss.vgg16 = _vgg16
    return ref_module, ref_loss, ref_ssim
