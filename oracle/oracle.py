"""ctypes binding + layer composition for the CPU oracle (oracle/mvs_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of mvs_oracle.c.  Only tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__.smoke() import this module; the product
package mvs_b200 never does.

All arrays are NumPy float32, C-contiguous, in the reference's layouts (NCHW / NCDHW).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmvs_oracle.so")

ALIGN_CORNERS = 1
PL_ORDER = 2
REF_SUM_SQUARED = 4
DEPTH_PLANE = 0
DEPTH_PIXEL = 1

_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mvs_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        try:
            build()          # no-op when the .so is newer than the source
        except Exception:
            if not os.path.exists(_SO):
                raise
        _lib = C.CDLL(_SO)
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def _depth_mode(depth, B, D):
    depth = np.ascontiguousarray(depth, np.float32)
    if depth.ndim == 2:
        assert depth.shape == (B, D) or depth.shape[0] == B
        return depth, DEPTH_PLANE
    assert depth.ndim == 4
    return depth, DEPTH_PIXEL


def homo_warp(src_fea, rot, trans, depth, flags=0):
    """src_fea [B,C,H,W], rot [B,3,3], trans [B,3], depth [B,D]|[B,D,H,W] -> [B,C,D,H,W]."""
    src_fea, p_src = _f(src_fea)
    B, Cc, H, W = src_fea.shape
    D = depth.shape[1]
    depth, mode = _depth_mode(depth, B, D)
    rot, p_rot = _f(np.asarray(rot).reshape(B, 9))
    trans, p_tr = _f(np.asarray(trans).reshape(B, 3))
    out = np.empty((B, Cc, D, H, W), np.float32)
    lib().mvso_homo_warp(p_src, p_rot, p_tr, depth.ctypes.data_as(C.POINTER(C.c_float)), mode,
                         out.ctypes.data_as(C.POINTER(C.c_float)), B, Cc, D, H, W, flags)
    return out


def warp_taps(rot, trans, depth, H, W, flags=0):
    """-> x0,y0 int32 [B,D,H,W], mask uint8 [B,D,H,W], ixy float32 [B,D,H,W,2]."""
    B, D = depth.shape[0], depth.shape[1]
    depth, mode = _depth_mode(depth, B, D)
    rot, p_rot = _f(np.asarray(rot).reshape(B, 9))
    trans, p_tr = _f(np.asarray(trans).reshape(B, 3))
    x0 = np.empty((B, D, H, W), np.int32)
    y0 = np.empty((B, D, H, W), np.int32)
    mask = np.empty((B, D, H, W), np.uint8)
    ixy = np.empty((B, D, H, W, 2), np.float32)
    lib().mvso_warp_taps(p_rot, p_tr, depth.ctypes.data_as(C.POINTER(C.c_float)), mode,
                         x0.ctypes.data_as(C.POINTER(C.c_int32)), y0.ctypes.data_as(C.POINTER(C.c_int32)),
                         mask.ctypes.data_as(C.POINTER(C.c_uint8)), ixy.ctypes.data_as(C.POINTER(C.c_float)),
                         B, D, H, W, flags)
    return x0, y0, mask, ixy


def cost_volume(ref, srcs, rot, trans, depth, flags=0, rows=None):
    """ref [B,C,H,W], srcs [nsrc,B,C,H,W], rot [B,nsrc,3,3], trans [B,nsrc,3] -> var [B,C,D,H,W]
    (rows=(y0, y1): only reference rows y0..y1-1 -> [B,C,D,y1-y0,W]; full-size checks in seconds)."""
    ref, p_ref = _f(ref)
    srcs, p_srcs = _f(srcs)
    nsrc = srcs.shape[0]
    B, Cc, H, W = ref.shape
    D = depth.shape[1]
    depth, mode = _depth_mode(depth, B, D)
    rot, p_rot = _f(np.asarray(rot).reshape(B, nsrc, 9))
    trans, p_tr = _f(np.asarray(trans).reshape(B, nsrc, 3))
    y0, y1 = rows if rows is not None else (0, H)
    assert 0 <= y0 <= y1 <= H
    out = np.empty((B, Cc, D, y1 - y0, W), np.float32)
    lib().mvso_cost_volume_rows(p_ref, p_srcs, nsrc, p_rot, p_tr, depth.ctypes.data_as(C.POINTER(C.c_float)),
                                mode, out.ctypes.data_as(C.POINTER(C.c_float)), B, Cc, D, H, W, flags, y0, y1)
    return out


def conv3d(x, w, bias=None, stride=1, transposed=False):
    x, p_x = _f(x)
    w, p_w = _f(w)
    B, Cin, D, H, W = x.shape
    Cout = w.shape[1] if transposed else w.shape[0]
    if transposed:
        oshape = (B, Cout, D * stride, H * stride, W * stride)
    else:
        oshape = (B, Cout, (D - 1) // stride + 1, (H - 1) // stride + 1, (W - 1) // stride + 1)
    y = np.empty(oshape, np.float32)
    if bias is not None:
        bias, p_b = _f(bias)
    else:
        p_b = None
    lib().mvso_conv3d(p_x, p_w, p_b, y.ctypes.data_as(C.POINTER(C.c_float)), B, Cin, Cout, D, H, W,
                      stride, int(bool(transposed)))
    return y


def bn_relu_skip(y, gamma, beta, mean, var, eps=1e-5, relu=True, skip=None):
    """In place on (a copy-free view of) y; returns y."""
    assert y.dtype == np.float32 and y.flags["C_CONTIGUOUS"]
    B, Cc = y.shape[:2]
    vol = int(np.prod(y.shape[2:]))
    g, pg = _f(gamma)
    b, pb = _f(beta)
    m, pm = _f(mean)
    v, pv = _f(var)
    if skip is not None:
        skip, ps = _f(skip)
        assert skip.shape == y.shape
    else:
        ps = None
    fn = lib().mvso_bn_relu_skip
    fn.argtypes = [C.POINTER(C.c_float)] * 5 + [C.c_float, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_long]
    fn(y.ctypes.data_as(C.POINTER(C.c_float)), pg, pb, pm, pv, eps, int(relu), ps, B, Cc, vol)
    return y


def softargmin_conf(logits, depth, clamp_index=False, want_prob=False):
    """logits [B,D,H,W] -> dict(depth, conf, index, expect_idx[, prob])."""
    logits, p_l = _f(logits)
    B, D, H, W = logits.shape
    depth, mode = _depth_mode(depth, B, D)
    od = np.empty((B, H, W), np.float32)
    oc = np.empty((B, H, W), np.float32)
    oi = np.empty((B, H, W), np.int64)
    oe = np.empty((B, H, W), np.float32)
    op = np.empty((B, D, H, W), np.float32) if want_prob else None
    lib().mvso_softargmin_conf(p_l, depth.ctypes.data_as(C.POINTER(C.c_float)), mode,
                               od.ctypes.data_as(C.POINTER(C.c_float)), oc.ctypes.data_as(C.POINTER(C.c_float)),
                               oi.ctypes.data_as(C.POINTER(C.c_int64)),
                               op.ctypes.data_as(C.POINTER(C.c_float)) if want_prob else None,
                               oe.ctypes.data_as(C.POINTER(C.c_float)), int(bool(clamp_index)), B, D, H, W)
    out = dict(depth=od, conf=oc, index=oi, expect_idx=oe)
    if want_prob:
        out["prob"] = op
    return out


def depth_range_samples(cur, interval, ndepth):
    cur, p = _f(cur)
    B, H, W = cur.shape
    out = np.empty((B, ndepth, H, W), np.float32)
    fn = lib().mvso_depth_range_samples
    fn.argtypes = [C.POINTER(C.c_float), C.c_double, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int]
    fn(p, float(interval), ndepth, out.ctypes.data_as(C.POINTER(C.c_float)), B, H, W)
    return out


# ------------------------------------------------------------------------------------------------
# CostRegNet topologies composed from the C layer functions.  `sd` is a {key: ndarray} state dict
# with the reference's own keys (SURVEY.md §5 "checkpoint / resume").
# ------------------------------------------------------------------------------------------------
def geo_pair(depth_ref, depth_src, cam, dist_thresh=1.0, rel_thresh=0.01):
    """C restatement of reproject_with_depth + check_geometric_consistency (MVSNet/eval.py:138-208) with the float64
    matmuls as explicit k-ordered FMA chains.  cam = the 60 float64 of mvs_b200.fusion.camera_block.
    Returns dict(mask uint8, depth_reprojected, x_src, y_src, x_reprojected, y_reprojected) [H,W]."""
    dr = np.ascontiguousarray(depth_ref, np.float32)
    ds = np.ascontiguousarray(depth_src, np.float32)
    cam = np.ascontiguousarray(cam, np.float64)
    assert cam.shape == (60,) and dr.shape == ds.shape and dr.ndim == 2
    H, W = dr.shape
    names = ("depth_reprojected", "x_src", "y_src", "x_reprojected", "y_reprojected")
    outs = {n: np.empty((H, W), np.float32) for n in names}
    mask = np.empty((H, W), np.uint8)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    fn = lib().mvso_geo_pair
    fn.restype = None
    fn(vp(dr), vp(ds), vp(cam), vp(mask), *[vp(outs[n]) for n in names], C.c_int(H), C.c_int(W), C.c_double(dist_thresh),
       C.c_float(rel_thresh))
    outs["mask"] = mask
    return outs


def _cbr(x, sd, conv_key, bn_key, stride=1, transposed=False, skip=None, eps=1e-5):
    y = conv3d(x, sd[conv_key + ".weight"], None, stride, transposed)
    return bn_relu_skip(y, sd[bn_key + ".weight"], sd[bn_key + ".bias"], sd[bn_key + ".running_mean"],
                        sd[bn_key + ".running_var"], eps, True, skip)


def costreg_mvsnet(x, sd):
    """MVSNet/models/mvsnet.py:48-93 (keys conv0.conv / conv0.bn, conv7.0 / conv7.1, prob w/ bias)."""
    c0 = _cbr(x, sd, "conv0.conv", "conv0.bn")
    c2 = _cbr(_cbr(c0, sd, "conv1.conv", "conv1.bn", 2), sd, "conv2.conv", "conv2.bn")
    c4 = _cbr(_cbr(c2, sd, "conv3.conv", "conv3.bn", 2), sd, "conv4.conv", "conv4.bn")
    y = _cbr(_cbr(c4, sd, "conv5.conv", "conv5.bn", 2), sd, "conv6.conv", "conv6.bn")
    y = _cbr(y, sd, "conv7.0", "conv7.1", 2, True, skip=c4)
    y = _cbr(y, sd, "conv9.0", "conv9.1", 2, True, skip=c2)
    y = _cbr(y, sd, "conv11.0", "conv11.1", 2, True, skip=c0)
    return conv3d(y, sd["prob.weight"], sd.get("prob.bias"))


def costreg_cas(x, sd):
    """CasMVSNet/models/module.py:407-438 (keys convN.conv / convN.bn for all blocks, prob no bias)."""
    c0 = _cbr(x, sd, "conv0.conv", "conv0.bn")
    c2 = _cbr(_cbr(c0, sd, "conv1.conv", "conv1.bn", 2), sd, "conv2.conv", "conv2.bn")
    c4 = _cbr(_cbr(c2, sd, "conv3.conv", "conv3.bn", 2), sd, "conv4.conv", "conv4.bn")
    y = _cbr(_cbr(c4, sd, "conv5.conv", "conv5.bn", 2), sd, "conv6.conv", "conv6.bn")
    y = _cbr(y, sd, "conv7.conv", "conv7.bn", 2, True, skip=c4)
    y = _cbr(y, sd, "conv9.conv", "conv9.bn", 2, True, skip=c2)
    y = _cbr(y, sd, "conv11.conv", "conv11.bn", 2, True, skip=c0)
    return conv3d(y, sd["prob.weight"], sd.get("prob.bias"))


def costreg_cvp(x, sd):
    """CVP-MVSNet/models/net.py:52-89; returns the squeezed [B,D,H,W] like the reference."""
    c0 = _cbr(_cbr(x, sd, "conv0.conv", "conv0.bn"), sd, "conv0a.conv", "conv0a.bn")
    c2 = _cbr(_cbr(_cbr(c0, sd, "conv1.conv", "conv1.bn", 2), sd, "conv2.conv", "conv2.bn"), sd,
              "conv2a.conv", "conv2a.bn")
    c4 = _cbr(_cbr(_cbr(c2, sd, "conv3.conv", "conv3.bn"), sd, "conv4.conv", "conv4.bn"), sd,
              "conv4a.conv", "conv4a.bn")
    c5 = _cbr(c4, sd, "conv5.0", "conv5.1", 1, True, skip=c2)
    c6 = _cbr(c5, sd, "conv6.0", "conv6.1", 2, True, skip=c0)
    return conv3d(c6, sd["prob0.weight"], sd.get("prob0.bias"))[:, 0]
