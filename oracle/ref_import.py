"""Import the UNMODIFIED reference modules (Module.py, Loss.py, ssim.py).

TEST / BASELINE INFRASTRUCTURE ONLY.  Source of the files: /root/reference where it is mounted (the build container —
oracle/make_golden.py generates tests/golden/*.pt from it and tests pin the oracle port against it), else the byte-identical
staged copy `oracle/_ref/` that oracle/build_ref.py makes (git-ignored, travels to the GPU box like a built .so), which is what
`bench.py --impl reference` / `cpu_baseline` time on the GPU box's host cores.  Nothing on the product path imports this,
and nothing that runs on the GPU box reads /root/reference.

Two shims:
  1. `osgeo` stub package (GDAL is not installed; CommonFunc.py:17-19 imports it at module scope);
  2. `Loss.vgg16` rebound to a random-init torchvision VGG16 (Loss.py:25 downloads ImageNet weights, and there
     is no network).  The perception term is out of scope (weight 0 in every parity run).
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_DIR = os.environ.get("FCD_REFERENCE_DIR", "/root/reference")
STAGED_DIR = os.path.join(_HERE, "_ref")
_SHIM = os.path.join(_HERE, "_shim")


def source_dir(prefer_staged: bool = False):
    """Directory the reference files are imported from: the mounted reference, else the staged copy, else None
    (`prefer_staged`: the staged copy first — bench.py, which must not read /root/reference at run time)."""
    for d in ((STAGED_DIR, REFERENCE_DIR) if prefer_staged else (REFERENCE_DIR, STAGED_DIR)):
        if all(os.path.isfile(os.path.join(d, f)) for f in ("Module.py", "Loss.py", "ssim.py", "CommonFunc.py")):
            return d
    return None


def available() -> bool:
    """The mounted reference (build container) — the live-pinning tests and golden generators need this one."""
    return os.path.isfile(os.path.join(REFERENCE_DIR, "Module.py"))


def importable() -> bool:
    return source_dir() is not None


def load(prefer_staged: bool = False):
    """Returns (Module, Loss, ssim) reference python modules."""
    src = source_dir(prefer_staged)
    if src is None:
        raise RuntimeError(f"reference neither mounted at {REFERENCE_DIR} nor staged at {STAGED_DIR} (oracle/build_ref.py)")
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    if src not in sys.path:
        sys.path.append(src)
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
