"""Deterministic synthetic DTU-shaped inputs for the cost-volume hot path.

Everything here is NumPy (legacy ``RandomState`` => bit-stable across NumPy versions and
machines), so the golden fixtures under ``tests/golden`` can be regenerated from seeds alone and
the GPU box never needs ``/root/reference``.

Rig (SURVEY.md §8(d) "cameras"): DTU pinhole ``fx=2892.33 fy=2883.18 cx=823.2 cy=619.07`` at
1600x1200, scaled to the feature-map extent; reference pose = identity looking down +z at a
scene centred ~680 mm; source cameras on a ring with 90-120 mm baselines, verging on the scene
centre.  ``proj = K @ [R|t]`` with last row ``[0,0,0,1]`` exactly as the reference dataset builds
it (MVSNet/datasets/dtu_yao.py:102-105).
"""
from __future__ import annotations

import math

import numpy as np

DTU_FX, DTU_FY, DTU_CX, DTU_CY = 2892.33, 2883.18, 823.2, 619.07
DTU_W = 1600.0
DTU_DEPTH_MIN, DTU_DEPTH_MAX = 425.0, 935.0
SCENE_Z = 680.0


def intrinsics(w: int) -> np.ndarray:
    """3x3 intrinsics (float64) for a feature map ``w`` pixels wide (same aspect as DTU)."""
    s = w / DTU_W
    return np.array([[DTU_FX * s, 0.0, DTU_CX * s], [0.0, DTU_FY * s, DTU_CY * s], [0.0, 0.0, 1.0]])


def _look_rotation(cx: float, cy: float, gain: float) -> np.ndarray:
    ay = math.atan2(cx, SCENE_Z) * gain   # yaw towards the optical axis of the reference
    ax = -math.atan2(cy, SCENE_Z) * gain
    rx = np.array([[1, 0, 0], [0, math.cos(ax), -math.sin(ax)], [0, math.sin(ax), math.cos(ax)]])
    ry = np.array([[math.cos(ay), 0, math.sin(ay)], [0, 1, 0], [-math.sin(ay), 0, math.cos(ay)]])
    return rx @ ry


def extrinsics(n_views: int, seed: int = 0) -> np.ndarray:
    """[n_views,4,4] float64 world->camera matrices; view 0 is the reference (identity)."""
    rng = np.random.RandomState(1000 + seed)
    ex = np.tile(np.eye(4), (n_views, 1, 1))
    for i in range(1, n_views):
        phi = 2.0 * math.pi * (i - 1) / max(n_views - 1, 1) + 0.3 + 0.1 * rng.uniform(-1, 1)
        base = rng.uniform(90.0, 120.0)
        c = np.array([base * math.cos(phi), base * math.sin(phi), rng.uniform(-8.0, 8.0)])
        r = _look_rotation(c[0], c[1], gain=0.8)
        ex[i, :3, :3] = r
        ex[i, :3, 3] = -r @ c
    return ex


def proj_matrices(n_views: int, w: int, seed: int = 0, batch: int = 1) -> np.ndarray:
    """MVSNet-style fused projection matrices [B,n_views,4,4] float32."""
    k = intrinsics(w)
    out = np.zeros((batch, n_views, 4, 4), np.float32)
    for b in range(batch):
        ex = extrinsics(n_views, seed + 17 * b)
        for v in range(n_views):
            p = np.eye(4)
            p[:3, :4] = k @ ex[v, :3, :4]
            out[b, v] = p.astype(np.float32)
    return out


def cas_proj_matrices(n_views: int, w: int, seed: int = 0, batch: int = 1) -> np.ndarray:
    """CasMVSNet-style [B,n_views,2,4,4]: [...,0]=extrinsic, [...,1,:3,:3]=intrinsic
    (CasMVSNet/datasets/general_eval.py builds it this way; DepthNet composes K@E itself,
    CasMVSNet/models/cas_mvsnet.py:30-33)."""
    k = intrinsics(w)
    out = np.zeros((batch, n_views, 2, 4, 4), np.float32)
    for b in range(batch):
        ex = extrinsics(n_views, seed + 17 * b)
        for v in range(n_views):
            out[b, v, 0] = ex[v].astype(np.float32)
            out[b, v, 1, :3, :3] = k.astype(np.float32)
    return out


def cvp_cameras(n_src: int, w: int, seed: int = 0, batch: int = 1):
    """CVP-MVSNet-style separate K / E: ref_in [B,3,3], src_in [B,nsrc,3,3], ref_ex [B,4,4],
    src_ex [B,nsrc,4,4] (CVP-MVSNet/models/net.py:99-118)."""
    k = intrinsics(w).astype(np.float32)
    ref_in = np.tile(k, (batch, 1, 1))
    src_in = np.tile(k, (batch, n_src, 1, 1))
    ref_ex = np.zeros((batch, 4, 4), np.float32)
    src_ex = np.zeros((batch, n_src, 4, 4), np.float32)
    for b in range(batch):
        ex = extrinsics(n_src + 1, seed + 17 * b)
        ref_ex[b] = ex[0]
        src_ex[b] = ex[1:]
    return ref_in, src_in, ref_ex, src_ex


def depth_planes(d: int, batch: int = 1, lo: float = DTU_DEPTH_MIN, hi: float = DTU_DEPTH_MAX) -> np.ndarray:
    """[B,D] float32 fronto-parallel hypotheses ``lo + k*(hi-lo)/d`` (cfg1: 425 + 10.6k)."""
    step = (hi - lo) / d
    return np.tile((lo + step * np.arange(d)).astype(np.float32), (batch, 1))


def depth_surface(h: int, w: int, batch: int = 1) -> np.ndarray:
    """[B,h,w] float32 smooth synthetic depth map (SURVEY §8(d) cfg3 stages 2/3)."""
    y, x = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    z = SCENE_Z + 80.0 * np.sin(2 * math.pi * x / w) * np.cos(2 * math.pi * y / h)
    return np.tile(z.astype(np.float32), (batch, 1, 1))


def depth_per_pixel(d: int, h: int, w: int, interval: float, batch: int = 1) -> np.ndarray:
    """[B,D,h,w] float32 per-pixel hypotheses centred on ``depth_surface`` with spacing
    ``interval`` (the shape CasMVSNet stages 2/3 and CVP refine levels feed the warp)."""
    centre = depth_surface(h, w, batch)[:, None]
    k = (np.arange(d, dtype=np.float32) - d / 2.0).reshape(1, d, 1, 1)
    return (centre + k * np.float32(interval)).astype(np.float32)


def features(n_views: int, c: int, h: int, w: int, seed: int = 0, batch: int = 1) -> np.ndarray:
    """[n_views,B,C,h,w] float32 ~N(0,1) feature maps."""
    rng = np.random.RandomState(2000 + seed)
    return rng.standard_normal((n_views, batch, c, h, w)).astype(np.float32)


def images_u8(n_views: int, h: int, w: int, seed: int = 0, batch: int = 1) -> np.ndarray:
    """[B,n_views,3,h,w] uint8 images: smooth structure + noise, as a decoded JPEG would hand them to the loader
    (CasMVSNet/datasets/general_eval.py:81-86 then scales by 1/255)."""
    rng = np.random.RandomState(4000 + seed)
    y, x = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    out = np.empty((batch, n_views, 3, h, w), np.uint8)
    for b in range(batch):
        for v in range(n_views):
            for c in range(3):
                ph = rng.uniform(0, 2 * math.pi, 2)
                f = rng.uniform(0.02, 0.2, 2)
                img = 127.5 + 70.0 * np.sin(f[0] * x + ph[0]) * np.cos(f[1] * y + ph[1]) + 25.0 * rng.standard_normal((h, w))
                out[b, v, c] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return out


def fill_state_dict(shapes: dict, seed: int = 0) -> dict:
    """Deterministic weights for a CostRegNet ``state_dict`` given ``{key: shape}``.

    Conv weights ~ U(-a,a) with a = 1/sqrt(fan_in) (PyTorch-default-like scale), BN weight
    ~U(0.5,1.5), BN bias ~N(0,0.1), running_mean ~N(0,0.1), running_var ~U(0.5,1.5) so that
    folding is exercised (SURVEY §8(d) "value distributions").  Keys are visited in sorted
    order so the result does not depend on dict ordering.
    """
    rng = np.random.RandomState(3000 + seed)
    out = {}
    for key in sorted(shapes):
        shp = tuple(shapes[key])
        if key.endswith("num_batches_tracked"):
            out[key] = np.zeros(shp, np.int64)
        elif key.endswith("running_var"):
            out[key] = rng.uniform(0.5, 1.5, shp).astype(np.float32)
        elif key.endswith("running_mean"):
            out[key] = (0.1 * rng.standard_normal(shp)).astype(np.float32)
        elif len(shp) == 5:
            fan = shp[1] * 27
            a = 1.0 / math.sqrt(fan)
            out[key] = rng.uniform(-a, a, shp).astype(np.float32)
        elif len(shp) == 4:        # 2D convolution of the feature extractor: Kaiming-uniform scale keeps activations O(1)
            a = math.sqrt(6.0 / (shp[1] * shp[2] * shp[3]))
            out[key] = rng.uniform(-a, a, shp).astype(np.float32)
        elif key.endswith("weight"):
            out[key] = rng.uniform(0.5, 1.5, shp).astype(np.float32)
        else:  # bias
            out[key] = (0.1 * rng.standard_normal(shp)).astype(np.float32)
    return out


# BASELINE.json configs (feature-map extents).  "stages": (C, D, h, w).
CONFIGS = {
    "cfg1": dict(family="mvsnet", n_views=4, batch=1, dtype="f32", stages=[(32, 48, 128, 160)]),
    "cfg2": dict(family="mvsnet", n_views=5, batch=4, dtype="bf16", stages=[(32, 192, 128, 160)]),
    "cfg3": dict(family="cas", n_views=5, batch=1, dtype="bf16",
                 stages=[(32, 48, 296, 400), (16, 32, 592, 800), (8, 8, 1184, 1600)]),
    "cfg5": dict(family="cas", n_views=7, batch=4, dtype="bf16",
                 stages=[(32, 64, 264, 480), (16, 32, 528, 960), (8, 8, 1056, 1920)]),
}


def warp_variance_bytes(n_views, batch, c, d, h, w, s_f, s_v, per_pixel_depth) -> int:
    """Algorithmic bytes of one fused warp+variance call:
    every feature map read once, the variance volume written once, hypotheses, cameras."""
    feat = n_views * batch * c * h * w * s_f
    vol = batch * c * d * h * w * s_v
    dep = batch * d * (h * w if per_pixel_depth else 1) * 4
    cam = batch * (n_views - 1) * 48
    return feat + vol + dep + cam


def costreg_flops(cin: int, d: int, h: int, w: int, base: int = 8) -> int:
    """Algorithmic FLOPs of the MVSNet/Cas CostRegNet."""
    v = d * h * w
    f = 2 * 27
    tot = f * cin * base * v
    tot += f * base * 2 * base * (v // 8) + f * 2 * base * 2 * base * (v // 8)
    tot += f * 2 * base * 4 * base * (v // 64) + f * 4 * base * 4 * base * (v // 64)
    tot += f * 4 * base * 8 * base * (v // 512) + f * 8 * base * 8 * base * (v // 512)
    tot += f * 8 * base * 4 * base * (v // 512) + f * 4 * base * 2 * base * (v // 64)
    tot += f * 2 * base * base * (v // 8) + f * base * 1 * v
    return tot
