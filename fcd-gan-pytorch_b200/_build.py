"""In-tree build of libfcd_b200.so (hand-written sm_100a CUDA behind a C ABI).

`python -m fcdgan_b200._build` or `__graft_entry__.build()` compiles every `csrc/*.cu` with
`nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo` and links one shared library next to this file.
Objects are cached under `build/` keyed by source mtime, so incremental rebuilds take seconds.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
BUILD = os.path.join(ROOT, "build", "fcd_b200")
LIB = os.path.join(HERE, "libfcd_b200.so")
PROBES_LIB = os.path.join(HERE, "libfcd_b200_probes.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(ROOT, "include", "fcd_b200.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src: str, hdr_mtime: float, verbose: bool) -> str:
    obj = os.path.join(BUILD, os.path.basename(src)[:-3] + ".o")
    if os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_mtime):
        return obj
    cmd = [NVCC, *NVCC_FLAGS, "-c", src, "-o", obj]
    if verbose:
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    srcs = _sources()
    hdr_mtime = _headers_mtime()
    if force:
        for f in os.listdir(BUILD):
            os.remove(os.path.join(BUILD, f))
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, hdr_mtime, verbose), srcs))
    if (not os.path.exists(LIB)) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
        if verbose:
            print(" ".join(cmd), flush=True)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


def build_probes(verbose: bool = False) -> str:
    """libfcd_b200_probes.so: the tcgen05 bring-up probes / UMMA issue-rate micro-benchmarks (csrc/probes/, declared in
    include/fcd_b200_probes.h).  Not part of the product library and not built by build(); scripts/gpu_probe.py,
    probe_halo.py and umma_bench*.py ask for it."""
    build(verbose=verbose)
    src = os.path.join(CSRC, "probes", "probe_tc.cu")
    if os.path.exists(PROBES_LIB) and os.path.getmtime(PROBES_LIB) >= max(os.path.getmtime(src), _headers_mtime()):
        return PROBES_LIB
    cmd = [NVCC, *NVCC_FLAGS, "-shared", src, "-o", PROBES_LIB, "-L" + HERE, "-l:libfcd_b200.so",
           "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"]
    if verbose:
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
    return PROBES_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    if "--probes" in sys.argv:
        print(build_probes(verbose=True))
