"""ctypes binding of libpcodec_b200.so (the C-ABI declared in include/pcodec_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails, we raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PCODEC_LIB") or os.path.join(HERE, "libpcodec_b200.so")  # (PCODEC_LIB: bisecting builds)

MAX_SEGMENTS = 4
MAX_TAPS = 25

OK, ERR_BAD_ARG, ERR_UNSUPPORTED, ERR_OVERFLOW = 0, 1, 2, 3

MASK_ONES, MASK_ZEROS, MASK_THRESHOLD = 0, 1, 2

EPI_LINEAR, EPI_GELU, EPI_ADD, EPI_ADD_GELU, EPI_GATE, EPI_GDN, EPI_IGDN, EPI_LRP, EPI_CLAMP01, EPI_LEAKY, EPI_LEAKY_ADD = range(11)
FLAG_SQUARE_INPUT, FLAG_PIXEL_SHUFFLE2, FLAG_SUBPIXEL_NCHW, FLAG_NO_F32_OUT, FLAG_SQUARE_OUT_PLANES = 1, 2, 4, 8, 16


class Segment(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("channels", C.c_int), ("pixel_stride", C.c_int)]


class Planes(C.Structure):
    _fields_ = [("hi", C.c_void_p), ("lo", C.c_void_p), ("pixel_stride", C.c_int)]


CONV_PLAN_BYTES = 2048


class ConvDesc(C.Structure):
    _fields_ = [
        ("seg", Segment * MAX_SEGMENTS),
        ("n_segments", C.c_int), ("batch", C.c_int), ("in_h", C.c_int), ("in_w", C.c_int),
        ("n_taps", C.c_int), ("dy", C.c_int8 * MAX_TAPS), ("dx", C.c_int8 * MAX_TAPS), ("in_step", C.c_int),
        ("weight", C.c_void_p), ("bias", C.c_void_p),
        ("cin_total", C.c_int), ("cout", C.c_int),
        ("grid_h", C.c_int), ("grid_w", C.c_int), ("out_step", C.c_int), ("out_off_y", C.c_int),
        ("out_off_x", C.c_int), ("out_h", C.c_int), ("out_w", C.c_int),
        ("out", C.c_void_p), ("out_pixel_stride", C.c_int),
        ("epilogue", C.c_int), ("flags", C.c_int),
        ("r1", C.c_void_p), ("r1_pixel_stride", C.c_int),
        ("r2", C.c_void_p), ("r2_pixel_stride", C.c_int),
        ("tc_weights", C.c_void_p), ("tc_split", C.c_int),
        ("seg16", Planes * MAX_SEGMENTS), ("out_hi", C.c_void_p), ("out_lo", C.c_void_p), ("out_plane_stride", C.c_int),
        ("plan", C.c_void_p), ("r1_16", Planes), ("r2_16", Planes),
    ]


class PcodecError(RuntimeError):
    pass


_ERR = {1: "bad argument", 2: "unsupported configuration", 3: "capacity overflow"}

_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); every symbol include/pcodec_b200.h declares
PROTOTYPES = {
    "pcodec_version": (_i, []),
    "pcodec_device_info": (_i, [_vp, _vp, _vp]),
    "pcodec_launch_count": (_i64, []),
    "pcodec_reset_launch_count": (None, []),
    "pcodec_set_sync_launches": (None, [_i]),
    "pcodec_error_string": (C.c_char_p, [_i]),
    "pcodec_recent_launches": (_i, [C.c_char_p, _i]),
    "pcodec_pmf_to_quantized_cdf": (_i, [_vp, _i, _i, _vp]),
    "pcodec_rans_encode_batch": (_i, [_vp, _vp, _i, _i64, _vp, _i, _vp, _vp, _i, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp]),
    "pcodec_rans_decode_batch": (_i, [_vp, _vp, _i, _i64, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp]),
    "pcodec_rans_decode_ranges": (_i, [_vp, _vp, _vp, _i, _i64, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp]),
    "pcodec_rans_encode_segments": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _i, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp]),
    "pcodec_rans_decode_segments": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp]),
    "pcodec_masked_residual": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _i64, _i, _i, _vp, _i, _vp, _vp, _i, _vp]),
    "pcodec_layer_partition": (_i, [_vp, _i, _i, _i64, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "pcodec_selftest_rans_core_encode": (_i64, [_vp, _vp, _i64, _vp, _i, _vp, _vp, _vp, _i64]),
    "pcodec_quantile_threshold": (_i, [_vp, _i, _i64, _i, _i, _f, _vp, _vp, _vp]),
    "pcodec_slice_quantize": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i64, _i, _i, _vp, _vp, _i, _f,
                                   _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "pcodec_slice_quantize_cust": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i64, _i, _i, _vp, _vp, _i, _f,
                                        _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp]),
    "pcodec_slice_indexes": (_i, [_vp, _i, _i, _i64, _i, _i, _vp, _vp, _i, _f, _vp, _vp]),
    "pcodec_slice_dequantize": (_i, [_vp, _vp, _i, _i, _i64, _i, _vp, _i, _vp]),
    "pcodec_bottleneck_quantize": (_i, [_vp, _i, _vp, _i, _i64, _i, _vp, _vp, _vp, _i, _vp]),
    "pcodec_bottleneck_dequantize": (_i, [_vp, _vp, _i, _i64, _i, _vp, _i, _vp]),
    "pcodec_bottleneck_indexes": (_i, [_i, _i64, _i, _vp, _vp]),
    "pcodec_bottleneck_likelihood": (_i, [_vp, _i, _vp, _i, _i64, _i, _vp, _vp]),
    "pcodec_conv_taps": (_i, [C.POINTER(ConvDesc), _i, _vp]),
    "pcodec_conv_plan": (_i, [C.POINTER(ConvDesc)]),
    "pcodec_split_planes": (_i, [_vp, _i, _i64, _i, _vp, _vp, _i, _i, _vp]),
    "pcodec_conv_tc_prepare": (_i, [_vp, _i, _i, _i, C.POINTER(C.c_void_p), _vp]),
    "pcodec_conv_tc_release": (None, [_vp]),
    "pcodec_window_attention": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "pcodec_im2col_nchw": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "pcodec_nchw_to_nhwc": (_i, [_vp, _vp, _i, _i, _i64, _i, _i, _vp]),
    "pcodec_nhwc_to_nchw": (_i, [_vp, _i, _vp, _i, _i, _i64, _vp]),
}

_lib = None


def build_library(verbose: bool = False) -> str:
    """Compile csrc/ for sm_100a (nvcc cross-compiles without a GPU)."""
    script = os.path.join(HERE, "csrc", "build.sh")
    out = subprocess.run(["bash", script], capture_output=True, text=True)
    if out.returncode != 0:
        raise PcodecError("building libpcodec_b200.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)
    return LIB_PATH


def lib():
    """Load the shared library (raises PcodecError when it is absent — there is no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise PcodecError(
                f"{LIB_PATH} not found: the CUDA extension is required (run `python -c 'import __graft_entry__ as g; "
                f"g.build()'` or progressivecodec_b200/csrc/build.sh). There is no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc == 0:
        return
    if rc > 0:
        raise PcodecError(f"pcodec {what}: {_ERR.get(rc, 'error ' + str(rc))}")
    raise PcodecError(f"pcodec {what}: CUDA error {-rc} ({lib().pcodec_error_string(rc).decode()})\n"
                      f"last launches of this process (oldest first):\n{recent_launches()}")


def recent_launches() -> str:
    """Flight recorder of the library: entry points of the last launches (for attributing an asynchronous device fault)."""
    buf = C.create_string_buffer(8192)
    lib().pcodec_recent_launches(buf, len(buf))
    return buf.value.decode(errors="replace")
