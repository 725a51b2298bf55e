"""Recipe for `oracle/_ref/`: the UNMODIFIED reference files of the hot path, staged so that they can travel to the GPU box.

TEST / BASELINE INFRASTRUCTURE ONLY.  `/root/reference` exists in the build container but not on the GPU box, and the
reference is a set of flat Python scripts that cannot be pip-installed (no setup.py / pyproject.toml).  This recipe copies the
four files the hot path lives in — byte for byte, SHA-256 recorded in MANIFEST.json — from where they lie under
`/root/reference` into `oracle/_ref/`, which is git-ignored (never part of the repository's history) but not
gpurun-ignored, exactly like a compiled `oracle/_ref/*.so` of a C reference would be.  `bench.py --impl reference` and the
`cpu_baseline` leg import the reference classes from there (via oracle/ref_import.py) so that the reference arm times the
reference's own code, not the oracle port.  `__graft_entry__.build()` runs this whenever `/root/reference` is present.

    python oracle/build_ref.py            # (re)stage;  prints the manifest
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("FCD_REFERENCE_DIR", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
# Module.py (networks), Loss.py (loss stack), ssim.py (MS-SSIM) and CommonFunc.py (wildcard-imported by the first two:
# Module.py:12, Loss.py:14)
FILES = ("Module.py", "Loss.py", "ssim.py", "CommonFunc.py")


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def build(verbose=False):
    """-> manifest dict, or None when /root/reference is not mounted (then whatever is already staged is kept)."""
    if not os.path.isfile(os.path.join(REF_SRC, "Module.py")):
        return None
    os.makedirs(REF_DST, exist_ok=True)
    manifest = {"source": REF_SRC, "files": {}}
    for f in FILES:
        src, dst = os.path.join(REF_SRC, f), os.path.join(REF_DST, f)
        if not (os.path.exists(dst) and _sha(dst) == _sha(src)):
            shutil.copyfile(src, dst)
        manifest["files"][f] = _sha(dst)
    with open(os.path.join(REF_DST, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    if verbose:
        print(json.dumps(manifest, indent=1))
    return manifest


def staged() -> bool:
    return all(os.path.isfile(os.path.join(REF_DST, f)) for f in FILES)


if __name__ == "__main__":
    m = build(verbose=True)
    if m is None:
        print(f"{REF_SRC} not mounted; staged copy present: {staged()}")
        sys.exit(0 if staged() else 1)
