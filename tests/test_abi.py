"""CPU: the C-ABI library loads without a GPU and exports every symbol include/fcd_b200.h declares (no compute
calls here).  The ctypes signatures are generated from the header itself (fcdgan_b200/_lib.py)."""
import ctypes
import os
import re
import subprocess

from fcdgan_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_parses_to_expected_entry_points():
    sigs = _lib.signatures()
    assert len(sigs) >= 45
    for must in ("fcd_conv2d_fwd", "fcd_conv2d_wgrad", "fcd_conv2d_dgrad_strided", "fcd_bn_finalize", "fcd_bn_act_fwd",
                 "fcd_maxpool2_fwd", "fcd_upsample2x_bilinear_fwd", "fcd_outconv_sigmoid_fwd", "fcd_masked_recon_fwd",
                 "fcd_ssim_level_fwd", "fcd_ssim_level_bwd", "fcd_msssim_combine_fwd", "fcd_region_loss_fwd",
                 "fcd_stage_nchw_to_split"):
        assert must in sigs
    # spot-check one prototype: 23 arguments, pointers / ints / stream in the declared order
    ret, args = sigs["fcd_conv2d_fwd"]
    assert ret is ctypes.c_int and len(args) == 23
    assert args[0] is ctypes.c_void_p and args[2] is ctypes.c_int and args[-1] is ctypes.c_void_p
    ret, args = sigs["fcd_conv2d_wgrad_workspace"]
    assert ret is ctypes.c_size_t and all(a is ctypes.c_int for a in args)
    assert sigs["fcd_bn_finalize"][1][2] is ctypes.c_double and sigs["fcd_bn_finalize"][1][9] is ctypes.c_float


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = _lib.load()
    for name in _lib.signatures():
        assert hasattr(lib, name), name
    assert lib.fcd_version() >= 100
    assert isinstance(lib.fcd_last_error(), bytes)


def test_no_undeclared_public_symbols():
    """every exported fcd_* symbol is declared in the header (the header is the whole public surface)."""
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\sT\s+(fcd_\w+)", out))
    assert exported == set(_lib.signatures()), exported ^ set(_lib.signatures())


def test_sass_uses_blackwell_tensor_and_tma_paths():
    """the conv engine really is tcgen05 + TMA: UTCHMMA / LDTM / UTMALDG appear in the sm_100a SASS."""
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        return
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "LDTM" in sass and "UTMALDG" in sass
    assert "sm_100a" in sass or "SM100" in sass.upper() or "sm_100" in sass


def test_argument_errors_are_reported_not_raised_in_c():
    """bad arguments return FCD_ERR_ARG with a message (no abort, no CUDA needed for the check)."""
    lib = _lib.load()
    rc = lib.fcd_pack_conv_weight(None, 1, 1, 3, 3, 16, 16, 0, None, None, None)
    assert rc == 1 and b"null pointer" in lib.fcd_last_error()
    rc = lib.fcd_ssim_level_fwd(1, 1, 1, 32, 32, 1, 13, 0.0, 0.0, 1, None, 0, None)
    assert rc == 3 and b"win_size" in lib.fcd_last_error()
