"""ctypes loader for the C-ABI library (include/mvs_b200.h -> mvs_b200/libmvs_b200.so).

There is NO fallback: if the library is missing or a call fails, the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MVS_B200_LIB points at an alternative build of the same C-ABI (kernel tuning experiments); default: the in-tree library
LIB_PATH = os.environ.get("MVS_B200_LIB") or os.path.join(_HERE, "libmvs_b200.so")

# flags / enums (mirror include/mvs_b200.h)
ALIGN_CORNERS = 1
PL_ORDER = 2
REF_SUM_SQUARED = 4
RELU = 8
CLAMP_INDEX = 16
INPUT_IS_PROB = 32
BLEND_BF16 = 64
FAST_COORDS = 128
FEAT_F16 = 256
WARP_NO_TMA = 512
WARP_TMA = 1024
ACT_F16 = 2048
X_DW, Y_DW, SKIP_DW = 4096, 8192, 16384      # conv3d_c8: W-de-interleaved input / output / skip tensor
SKIP_PS = 131072                             # conv3d_c8 + FLAT2D: pixel-shuffled half-resolution skip operand
FLAT2D = 65536                               # conv3d_c8 / pack: plain 2D convolution (D = 1), image rows tiled by the kernel
KD1 = 32768                                  # conv3d_c8: weights zero outside the centre depth tap (stacked 2D images)
DEPTH_PLANE = 0
DEPTH_PIXEL = 1
F32 = 0
BF16 = 1
F16 = 2
U8 = 3
MAX_SRC = 8

_vp, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double

# name -> (restype, argtypes); the single source of truth the symbol test checks against the header
SIGNATURES = {
    "mvs_version": (_i, []),
    "mvs_sm": (_i, []),
    "mvs_last_error": (C.c_char_p, []),
    "mvs_launch_count": (_i64, []),
    "mvs_warp_fwd": (_i, [_vp] * 4 + [_i, _vp] + [_i] * 6 + [_vp]),
    "mvs_warp_bwd": (_i, [_vp] * 4 + [_i, _vp] + [_i] * 6 + [_vp]),
    "mvs_warp_taps": (_i, [_vp] * 3 + [_i] + [_vp] * 4 + [_i] * 5 + [_vp]),
    "mvs_warp_variance_fwd": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _vp] + [_i] * 6 + [_vp]),
    "mvs_warp_variance_bwd": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp] + [_i] * 6 + [_vp]),
    "mvs_warp_variance_c8_fwd": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _vp] + [_i] * 6 + [_vp]),
    "mvs_pack_c8": (_i, [_vp, _i, _vp, _i, _i, _i64, _vp]),
    "mvs_unpack_c8": (_i, [_vp, _vp, _i, _i, _i, _i64, _vp]),
    "mvs_pack_c8h": (_i, [_vp, _i, _vp, _i, _i, _i64, _vp]),
    "mvs_img_to_c8h": (_i, [_vp, _i, _vp, _i, _i, _i, _vp]),
    "mvs_s2d_c8": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "mvs_fpn_merge_c8h": (_i, [_vp] * 5 + [_i] * 7 + [_vp]),
    "mvs_border_add_c8h": (_i, [_vp, _vp] + [_i] * 4 + [_vp]),
    "mvs_conv3d_fwd": (_i, [_vp] * 6 + [_i] * 9 + [_vp]),
    "mvs_conv3d_wgrad": (_i, [_vp] * 3 + [_i] * 8 + [_vp]),
    "mvs_bn_stats": (_i, [_vp, _vp, _i, _i, _i64, _vp]),
    "mvs_bn_apply": (_i, [_vp] * 4 + [_i, _i, _i64, _i, _vp]),
    "mvs_bn_bwd_stats": (_i, [_vp] * 7 + [_i, _i, _i64, _i, _vp]),
    "mvs_bn_bwd_apply": (_i, [_vp] * 10 + [_i, _i, _i64, _i, _vp]),
    "mvs_conv3d_c8_packed_weight_bytes": (_i64, [_i] * 4),
    "mvs_conv3d_c8_pack_weights": (_i, [_vp, _vp] + [_i] * 4 + [_vp]),
    "mvs_conv3d_c8_pack_weights_ex": (_i, [_vp, _vp] + [_i] * 5 + [_vp]),
    "mvs_conv3d_c8_fwd": (_i, [_vp] * 6 + [_i] * 9 + [_vp]),
    "mvs_conv3d_c8_set_trace": (_i, [_vp, _i]),
    "mvs_softargmin_conf_fwd": (_i, [_vp, _vp, _i] + [_vp] * 4 + [_i] * 5 + [_vp]),
    "mvs_depth_range_samples": (_i, [_vp, _d, _i, _vp, _i, _i, _i, _vp]),
    "mvs_geo_consistency": (_i, [_vp] * 9 + [_i, _i, _d, C.c_float, _i, _vp]),
    "mvs_geo_backproject": (_i, [_vp] * 4 + [_i, _i, _vp]),
    "mvs_geo_fuse": (_i, [_vp, _vp, _vp, _i] + [_vp] * 6 + [_i, _i, _d, C.c_float, C.c_float, _i, _vp]),
    "mvs_cas_poses": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "mvs_cvp_depth_interval": (_i, [_vp, _vp, _vp, _i, _i, _i, _d, _vp]),
    "mvs_cas_hypotheses": (_i, [_vp] + [_i] * 7 + [_d, _vp, _i, _vp]),
}

_lib = None


class MvsError(RuntimeError):
    pass


def lib():
    """Load (once) and return the CDLL; raises MvsError when the extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MvsError(
                f"{LIB_PATH} is missing: build it with `python -m mvs_b200.csrc.build` "
                "(__graft_entry__.build()).  mvs_b200 has no CPU or PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(code: int, what: str):
    if code != 0:
        msg = lib().mvs_last_error().decode("utf-8", "replace")
        raise MvsError(f"{what} failed ({code}): {msg}")


def launch_count() -> int:
    return int(lib().mvs_launch_count())
