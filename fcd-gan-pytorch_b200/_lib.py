"""ctypes binding of libfcd_b200.so (the C ABI declared in include/fcd_b200.h).

The argument/return types of every entry point are PARSED FROM THE HEADER, so the header is the single
source of truth for the ABI (tests/test_abi.py checks that the library exports every declared symbol).

The product path has no CPU or library fallback: if the shared library is missing or a call returns a
non-zero status, a RuntimeError is raised (`FcdError`).
"""
from __future__ import annotations

import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "libfcd_b200.so")
HEADER = os.path.join(ROOT, "include", "fcd_b200.h")


class FcdError(RuntimeError):
    pass


def _ctype(decl: str):
    decl = decl.strip()
    if "*" in decl:
        return ctypes.c_char_p if decl.replace(" ", "") == "constchar*" else ctypes.c_void_p
    base = re.sub(r"\bconst\b", "", decl).split()
    # drop the parameter name (last identifier) when a type is present
    words = base[:-1] if len(base) > 1 and base[-1] not in ("int", "float", "double", "size_t", "long") else base
    t = " ".join(words)
    return {"int": ctypes.c_int, "float": ctypes.c_float, "double": ctypes.c_double, "size_t": ctypes.c_size_t,
            "long long": ctypes.c_longlong, "void": None}[t]


def parse_header(path: str = HEADER):
    """-> {name: (restype, [argtypes])} for every `fcd_*` prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    src = "\n".join(l for l in src.splitlines() if not l.lstrip().startswith("#"))
    out = {}
    for m in re.finditer(r"(const\s+char\s*\*|size_t|int)\s+(fcd_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = ctypes.c_char_p if "char" in ret else (ctypes.c_size_t if ret == "size_t" else ctypes.c_int)
        argtypes = [] if args in ("", "void") else [_ctype(a) for a in args.split(",")]
        out[name] = (restype, argtypes)
    return out


_lib = None
_sigs = None


def signatures():
    global _sigs
    if _sigs is None:
        _sigs = parse_header()
    return _sigs


def load():
    """Load (once) and return the ctypes handle; raises FcdError if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FcdError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the fcdgan_b200 hot path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in signatures().items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    # A/B switches without code changes: FCD_OPTIONS="conv_halo=0,conv_tma_out=0" (see fcd_set_option in the header)
    for item in filter(None, os.environ.get("FCD_OPTIONS", "").split(",")):
        name, _, value = item.partition("=")
        if lib.fcd_set_option(name.strip().encode(), int(value or 1)) != 0:
            raise FcdError(f"FCD_OPTIONS: {lib.fcd_last_error().decode(errors='replace')}")
    return lib


_fn_cache = {}


def call(name: str, *args):
    """Call an int-returning entry point and raise FcdError with the library's message on failure."""
    fn = _fn_cache.get(name)
    if fn is None:
        fn = _fn_cache[name] = getattr(load(), name)
    rc = fn(*args)
    if rc != 0:
        raise FcdError(f"{name} -> {rc}: {load().fcd_last_error().decode(errors='replace')}")
    return rc


_probes = None


def call_probe(name: str, *args):
    """Entry points of the separate bring-up library libfcd_b200_probes.so (include/fcd_b200_probes.h); built on demand."""
    global _probes
    if _probes is None:
        from . import _build

        load()
        lib = ctypes.CDLL(_build.build_probes())
        for fname, (restype, argtypes) in parse_header(os.path.join(ROOT, "include", "fcd_b200_probes.h")).items():
            fn = getattr(lib, fname)
            fn.argtypes, fn.restype = argtypes, restype
        _probes = lib
    rc = getattr(_probes, name)(*args)
    if rc != 0:
        raise FcdError(f"{name} -> {rc}: {load().fcd_last_error().decode(errors='replace')}")
    return rc


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch

    return torch.cuda.current_stream().cuda_stream
