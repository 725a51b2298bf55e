"""ctypes binding of libfcd_b200.so (the C ABI declared in include/fcd_b200.h).

The product path has no CPU or library fallback: if the shared library is missing or a call returns a
non-zero status, a RuntimeError is raised (`FcdError`).
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfcd_b200.so")


class FcdError(RuntimeError):
    pass


_T = {"p": ctypes.c_void_p, "i": ctypes.c_int, "f": ctypes.c_float, "d": ctypes.c_double,
      "z": ctypes.c_size_t, "l": ctypes.c_longlong}

# name -> (argument codes, return code); 'p' pointer, 'i' int, 'f' float, 'd' double, 'z' size_t, 'l' long long.
SIGNATURES = {
    "fcd_version": ("", "i"),
    "fcd_conv2d_tc_supported": ("iiiii", "i"),
    "fcd_pack_conv_weight": ("piiiiiiippp", "i"),
    "fcd_conv2d_fwd": ("ppipppp" + "i" * 10 + "ppip", "i"),
    "fcd_conv2d_dgrad_strided": ("ppipppi" + "i" * 9 + "p", "i"),
    "fcd_conv2d_wgrad_workspace": ("i" * 10, "z"),
    "fcd_conv2d_wgrad": ("ppippipp" + "i" * 12 + "pzip", "i"),
    "fcd_debug_umma_probe": ("ppp" + "i" * 13 + "p", "i"),
}

_lib = None


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
    lib.fcd_last_error.restype = ctypes.c_char_p
    lib.fcd_last_error.argtypes = []
    for name, (args, ret) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = [_T[a] for a in args]
        fn.restype = _T[ret]
    _lib = lib
    return lib


def register(name: str, args: str, ret: str = "i"):
    SIGNATURES[name] = (args, ret)
    if _lib is not None:
        fn = getattr(_lib, name)
        fn.argtypes = [_T[a] for a in args]
        fn.restype = _T[ret]


def call(name: str, *args):
    """Call an int-returning entry point and raise FcdError with the library's message on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise FcdError(f"{name} -> {rc}: {lib.fcd_last_error().decode(errors='replace')}")
    return rc


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch

    return torch.cuda.current_stream().cuda_stream
