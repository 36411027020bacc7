"""Geometric-consistency filter + per-view depth fusion on the GPU (the step after the hot path).

Same names, arguments and return values as the reference's NumPy / cv2 functions
(MVSNet/eval.py:138-208 == CasMVSNet/test.py:237-294):

    reproject_with_depth(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src)
        -> depth_reprojected, x_reprojected, y_reprojected, x_src, y_src
    check_geometric_consistency(...same...) -> mask, depth_reprojected, x2d_src, y2d_src

NumPy in -> NumPy out (drop-in for eval.py / test.py, which call them with arrays read from .pfm files);
CUDA tensors in -> CUDA tensors out (no host round trip: feed the depth / confidence maps of `cascade_hot_path`
straight in).  `fuse_ref_view` is the loop body of `filter_depth` (MVSNet/eval.py:240-263) for ALL source views of one
reference view in ONE kernel launch.  The 3x3 / 4x4 camera algebra stays on the host in NumPy float64, exactly where
and how the reference computes it (np.linalg.inv, np.matmul): plumbing, 60 doubles per view pair.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np
import torch

from . import _lib as L
from .ops import _p, _stream, _ptr_array
from ._lib import lib, check


def _np64(a):
    return a.detach().cpu().numpy().astype(np.float64) if isinstance(a, torch.Tensor) else np.asarray(a, dtype=np.float64)


def camera_block(intrinsics_ref, extrinsics_ref, intrinsics_src, extrinsics_src) -> np.ndarray:
    """The 60 float64 the kernels consume, derived as the reference derives them (eval.py:151-176)."""
    Kr, Er, Ks, Es = _np64(intrinsics_ref), _np64(extrinsics_ref), _np64(intrinsics_src), _np64(extrinsics_src)
    if Kr.shape != (3, 3) or Ks.shape != (3, 3) or Er.shape != (4, 4) or Es.shape != (4, 4):
        raise ValueError("intrinsics must be 3x3 and extrinsics 4x4")
    return np.concatenate([np.linalg.inv(Kr).ravel(), np.matmul(Es, np.linalg.inv(Er))[:3].ravel(), Ks.ravel(),
                           np.linalg.inv(Ks).ravel(), np.matmul(Er, np.linalg.inv(Es))[:3].ravel(), Kr.ravel()])


def _device():
    if not torch.cuda.is_available():
        raise L.MvsError("mvs_b200.fusion needs a CUDA device (no CPU fallback)")
    lib()
    return torch.device("cuda", torch.cuda.current_device())


def _depth_in(a, dev):
    """float32 [H,W] on the GPU + whether the caller passed NumPy."""
    if isinstance(a, torch.Tensor):
        if not a.is_cuda:
            raise L.MvsError("mvs_b200.fusion takes NumPy arrays or CUDA tensors")
        return a.to(torch.float32).contiguous(), False
    a = np.ascontiguousarray(a, dtype=np.float32)
    return torch.from_numpy(a).to(dev), True


def _out(t, as_numpy, dtype=None):
    if not as_numpy:
        return t
    a = t.cpu().numpy()
    return a.astype(dtype) if dtype is not None else a


def _pair(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src, apply_mask, dist_thresh,
          rel_thresh):
    dev = depth_ref.device if isinstance(depth_ref, torch.Tensor) else _device()
    dr, as_np = _depth_in(depth_ref, dev)
    ds, _ = _depth_in(depth_src, dev)
    if dr.dim() != 2 or ds.shape != dr.shape:
        raise ValueError("depth maps must be [H, W] and of equal shape")
    H, W = dr.shape
    cam = torch.from_numpy(camera_block(intrinsics_ref, extrinsics_ref, intrinsics_src, extrinsics_src)).to(dev)
    mask = torch.empty((H, W), dtype=torch.uint8, device=dev)
    outs = [torch.empty((H, W), dtype=torch.float32, device=dev) for _ in range(5)]      # depth, x_src, y_src, x_rep, y_rep
    with torch.cuda.device(dev):
        check(lib().mvs_geo_consistency(_p(dr), _p(ds), _p(cam), _p(mask), _p(outs[0]), _p(outs[1]), _p(outs[2]), _p(outs[3]),
                                        _p(outs[4]), H, W, float(dist_thresh), float(rel_thresh), int(apply_mask), _stream()),
              "mvs_geo_consistency")
    return mask, outs, as_np


def reproject_with_depth(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src):
    """MVSNet/eval.py:138-185.  Returns depth_reprojected, x_reprojected, y_reprojected, x_src, y_src (float32 [H,W])."""
    _, o, as_np = _pair(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src, False, 1.0, 0.01)
    return tuple(_out(t, as_np) for t in (o[0], o[3], o[4], o[1], o[2]))


def check_geometric_consistency(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src,
                                dist_thresh=1.0, rel_thresh=0.01):
    """MVSNet/eval.py:188-208.  Returns mask (bool), depth_reprojected (0 where the check fails), x2d_src, y2d_src."""
    mask, o, as_np = _pair(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src, True,
                           dist_thresh, rel_thresh)
    return _out(mask.bool(), as_np), _out(o[0], as_np), _out(o[1], as_np), _out(o[2], as_np)


def fuse_ref_view(ref_depth_est, confidence, ref_intrinsics, ref_extrinsics, src_depth_ests: Sequence, src_intrinsics: Sequence,
                  src_extrinsics: Sequence, conf_thresh=0.8, min_views=3, dist_thresh=1.0, rel_thresh=0.01, per_source=False):
    """Loop body of filter_depth for one reference view (MVSNet/eval.py:240-263), all source views in one launch.

    Returns a dict: geo_mask_sum (int32), depth_est_averaged (float64, as NumPy's float32 / int32 division yields),
    photo_mask, geo_mask, final_mask (bool) [+ all_srcview_geomask, all_srcview_depth_ests when per_source]."""
    nsrc = len(src_depth_ests)
    if not (1 <= nsrc <= 16) or len(src_intrinsics) != nsrc or len(src_extrinsics) != nsrc:
        raise ValueError("1..16 source views with one intrinsics / extrinsics pair each")
    dev = ref_depth_est.device if isinstance(ref_depth_est, torch.Tensor) else _device()
    dr, as_np = _depth_in(ref_depth_est, dev)
    cf, _ = _depth_in(confidence, dev)
    srcs = [_depth_in(d, dev)[0] for d in src_depth_ests]
    H, W = dr.shape
    if cf.shape != dr.shape or any(s.shape != dr.shape for s in srcs):
        raise ValueError("all maps must be [H, W] of equal shape")
    cams = torch.from_numpy(np.concatenate([camera_block(ref_intrinsics, ref_extrinsics, K, E)
                                            for K, E in zip(src_intrinsics, src_extrinsics)])).to(dev)
    geo_sum = torch.empty((H, W), dtype=torch.int32, device=dev)
    avg = torch.empty((H, W), dtype=torch.float64, device=dev)
    final = torch.empty((H, W), dtype=torch.uint8, device=dev)
    masks = torch.empty((nsrc, H, W), dtype=torch.uint8, device=dev) if per_source else None
    reproj = torch.empty((nsrc, H, W), dtype=torch.float32, device=dev) if per_source else None
    with torch.cuda.device(dev):
        check(lib().mvs_geo_fuse(_p(dr), _p(cf), _ptr_array(srcs), nsrc, _p(cams), _p(geo_sum), _p(avg), _p(final), _p(masks),
                                 _p(reproj), H, W, float(dist_thresh), float(rel_thresh), float(conf_thresh), int(min_views),
                                 _stream()), "mvs_geo_fuse")
    out = {"geo_mask_sum": _out(geo_sum, as_np), "depth_est_averaged": _out(avg, as_np),
           "photo_mask": _out(cf > conf_thresh, as_np), "geo_mask": _out(geo_sum >= min_views, as_np),
           "final_mask": _out(final.bool(), as_np)}
    if per_source:
        out["all_srcview_geomask"] = _out(masks.bool(), as_np)
        out["all_srcview_depth_ests"] = _out(reproj, as_np)
    return out


def backproject(depth_est_averaged, valid_points, ref_intrinsics, ref_extrinsics):
    """World points of the valid pixels -- filter_depth, MVSNet/eval.py:293-301: returns `xyz_world.transpose((1, 0))`
    as float32 [n_valid, 3] in the reference's order (row-major over the image).  NumPy in -> NumPy out."""
    dev = depth_est_averaged.device if isinstance(depth_est_averaged, torch.Tensor) else _device()
    as_np = not isinstance(depth_est_averaged, torch.Tensor)
    d = (torch.from_numpy(np.ascontiguousarray(depth_est_averaged, dtype=np.float64)).to(dev) if as_np
         else depth_est_averaged.to(torch.float64).contiguous())
    m = valid_points if isinstance(valid_points, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(valid_points))
    m = m.to(dev).to(torch.uint8).contiguous()
    if d.dim() != 2 or m.shape != d.shape:
        raise ValueError("depth and mask must be [H, W] of equal shape")
    H, W = d.shape
    Kr, Er = _np64(ref_intrinsics), _np64(ref_extrinsics)
    cam = torch.from_numpy(np.concatenate([np.linalg.inv(Kr).ravel(), np.linalg.inv(Er)[:3].ravel()])).to(dev)
    xyz = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib().mvs_geo_backproject(_p(d), _p(m), _p(cam), _p(xyz), H, W, _stream()), "mvs_geo_backproject")
    pts = xyz[m.bool()]
    return _out(pts, as_np)


def filter_depth(depths, confidences, intrinsics, extrinsics, pairs, images=None, plyfilename=None, conf_thresh=0.8,
                 min_views=3, dist_thresh=1.0, rel_thresh=0.01):
    """The body of `filter_depth` (MVSNet/eval.py:212-326, CasMVSNet/test.py:297-410) WITHOUT the trip through the file
    system: the reference re-reads every depth / confidence map it has just written as .pfm (eval.py:232-236, 246) and
    loops over (ref, src) pairs in NumPy; here the maps stay on the device (feed `cascade_hot_path` / `CascadeMVSNet`
    outputs straight in), every reference view is ONE fused launch + one back-projection launch.

    depths / confidences: per-view [H,W] maps (NumPy or CUDA tensors); intrinsics / extrinsics: per-view 3x3 / 4x4;
    pairs: [(ref_view, [src_views...])] as `read_pair_file` returns; images: optional per-view [H,W,3] float in [0,1]
    ALREADY at the depth maps' resolution (the reference hard-codes a DTU crop `ref_img[1:-16:4, 1::4]`, eval.py:301).
    Returns (vertexs float32 [N,3], vertex_colors uint8 [N,3], per-view dicts); writes the PLY when `plyfilename` is given."""
    from . import io as mio
    vertexs, vertex_colors, per_view = [], [], []
    for ref_view, src_views in pairs:
        out = fuse_ref_view(depths[ref_view], confidences[ref_view], intrinsics[ref_view], extrinsics[ref_view],
                            [depths[s] for s in src_views], [intrinsics[s] for s in src_views],
                            [extrinsics[s] for s in src_views], conf_thresh=conf_thresh, min_views=min_views,
                            dist_thresh=dist_thresh, rel_thresh=rel_thresh)
        pts = backproject(out["depth_est_averaged"], out["final_mask"], intrinsics[ref_view], extrinsics[ref_view])
        pts = pts.cpu().numpy() if isinstance(pts, torch.Tensor) else pts
        fm = out["final_mask"].cpu().numpy() if isinstance(out["final_mask"], torch.Tensor) else out["final_mask"]
        vertexs.append(np.asarray(pts, np.float32))
        if images is not None:
            img = images[ref_view]
            img = img.cpu().numpy() if isinstance(img, torch.Tensor) else np.asarray(img)
            vertex_colors.append((img[fm] * 255).astype(np.uint8))           # eval.py:306
        else:
            vertex_colors.append(np.full((len(pts), 3), 255, np.uint8))
        per_view.append(out)
    vertexs = np.concatenate(vertexs, axis=0) if vertexs else np.zeros((0, 3), np.float32)
    vertex_colors = np.concatenate(vertex_colors, axis=0) if vertex_colors else np.zeros((0, 3), np.uint8)
    if plyfilename is not None:
        mio.write_ply(plyfilename, vertexs, vertex_colors)
    return vertexs, vertex_colors, per_view
