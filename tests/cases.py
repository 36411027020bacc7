"""Seeded input builders shared by tests/golden/make_golden.py (which feeds them to the unmodified
reference) and by the parity tests (which feed them to the oracle and to the CUDA path).
Only inputs that cannot be regenerated bit-for-bit from a seed are stored in the fixtures."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from mvs_b200 import synth  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def golden(name: str):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


# ---- reference CostRegNet state-dict shapes (probed from the reference modules) ----
def _cbr_shapes(prefix_conv, prefix_bn, cin, cout, transposed=False):
    w = (cin, cout, 3, 3, 3) if transposed else (cout, cin, 3, 3, 3)
    return {
        prefix_conv + ".weight": w,
        prefix_bn + ".weight": (cout,), prefix_bn + ".bias": (cout,),
        prefix_bn + ".running_mean": (cout,), prefix_bn + ".running_var": (cout,),
        prefix_bn + ".num_batches_tracked": (),
    }


def costreg_shapes(family: str, cin: int = 32, base: int = 8) -> dict:
    s = {}
    if family in ("mvsnet", "cas"):
        if family == "mvsnet":
            cin, base = 32, 8
        ch = [(cin, base), (base, 2 * base), (2 * base, 2 * base), (2 * base, 4 * base),
              (4 * base, 4 * base), (4 * base, 8 * base), (8 * base, 8 * base)]
        for i, (a, b) in enumerate(ch):
            s.update(_cbr_shapes(f"conv{i}.conv", f"conv{i}.bn", a, b))
        for name, (a, b) in zip(("conv7", "conv9", "conv11"),
                                ((8 * base, 4 * base), (4 * base, 2 * base), (2 * base, base))):
            if family == "mvsnet":
                s.update(_cbr_shapes(name + ".0", name + ".1", a, b, True))
            else:
                s.update(_cbr_shapes(name + ".conv", name + ".bn", a, b, True))
        s["prob.weight"] = (1, base, 3, 3, 3)
        if family == "mvsnet":
            s["prob.bias"] = (1,)
    elif family == "cvp":
        for name, (a, b) in (("conv0", (16, 16)), ("conv0a", (16, 16)), ("conv1", (16, 32)),
                             ("conv2", (32, 32)), ("conv2a", (32, 32)), ("conv3", (32, 64)),
                             ("conv4", (64, 64)), ("conv4a", (64, 64))):
            s.update(_cbr_shapes(name + ".conv", name + ".bn", a, b))
        s.update(_cbr_shapes("conv5.0", "conv5.1", 64, 32, True))
        s.update(_cbr_shapes("conv6.0", "conv6.1", 32, 16, True))
        s["prob0.weight"] = (1, 16, 3, 3, 3)
        s["prob0.bias"] = (1,)
    else:
        raise ValueError(family)
    return s


def costreg_state(family: str, cin: int = 32, base: int = 8, seed: int = 0) -> dict:
    return synth.fill_state_dict(costreg_shapes(family, cin, base), seed)


# ---- CasMVSNet FeatureNet (fpn, base 8): state-dict shapes of CasMVSNet/models/module.py:304-365 ----
def featurenet_shapes(base: int = 8) -> dict:
    s = {}

    def blk(name, cin, cout, k):
        s[name + ".conv.weight"] = (cout, cin, k, k)
        for t, shp in (("weight", (cout,)), ("bias", (cout,)), ("running_mean", (cout,)), ("running_var", (cout,)),
                       ("num_batches_tracked", ())):
            s[f"{name}.bn.{t}"] = shp

    b = base
    blk("conv0.0", 3, b, 3); blk("conv0.1", b, b, 3)
    blk("conv1.0", b, 2 * b, 5); blk("conv1.1", 2 * b, 2 * b, 3); blk("conv1.2", 2 * b, 2 * b, 3)
    blk("conv2.0", 2 * b, 4 * b, 5); blk("conv2.1", 4 * b, 4 * b, 3); blk("conv2.2", 4 * b, 4 * b, 3)
    s["out1.weight"] = (4 * b, 4 * b, 1, 1)
    s["inner1.weight"] = (4 * b, 2 * b, 1, 1); s["inner1.bias"] = (4 * b,)
    s["inner2.weight"] = (4 * b, b, 1, 1); s["inner2.bias"] = (4 * b,)
    s["out2.weight"] = (2 * b, 4 * b, 3, 3)
    s["out3.weight"] = (b, 4 * b, 3, 3)
    return s


def featurenet_state(seed: int = 0) -> dict:
    return synth.fill_state_dict(featurenet_shapes(), seed)


def mvsnet_featurenet_shapes() -> dict:
    """MVSNet/models/mvsnet.py:8-45."""
    s = {}
    for name, cin, cout, k in (("conv0", 3, 8, 3), ("conv1", 8, 8, 3), ("conv2", 8, 16, 5), ("conv3", 16, 16, 3), ("conv4", 16, 16, 3),
                               ("conv5", 16, 32, 5), ("conv6", 32, 32, 3)):
        s[name + ".conv.weight"] = (cout, cin, k, k)
        for t, shp in (("weight", (cout,)), ("bias", (cout,)), ("running_mean", (cout,)), ("running_var", (cout,)),
                       ("num_batches_tracked", ())):
            s[f"{name}.bn.{t}"] = shp
    s["feature.weight"] = (32, 32, 3, 3)
    s["feature.bias"] = (32,)
    return s


def mvsnet_model_state(seed: int = 60) -> dict:
    sd = {"feature." + k: v for k, v in synth.fill_state_dict(mvsnet_featurenet_shapes(), seed).items()}
    sd.update({"cost_regularization." + k: v for k, v in costreg_state("mvsnet", seed=seed + 1).items()})
    return sd


def mvsnet_model_case(n_views=3, B=2, H=64, W=96, D=8, seed=13):
    return dict(imgs_u8=synth.images_u8(n_views, H, W, seed, B), proj=synth.proj_matrices(n_views, W // 4, seed, B),
                depth=synth.depth_planes(D, B))


def full_model_case(n_views=3, B=1, H=64, W=96, ndepths=(16, 8, 8), seed=9):
    """Whole CascadeMVSNet from uint8 images (the loader's /255 scaling applied by the caller / on the device)."""
    projs = {f"stage{i + 1}": synth.cas_proj_matrices(n_views, W // s, seed, B) for i, s in enumerate((4, 2, 1))}
    return dict(imgs_u8=synth.images_u8(n_views, H, W, seed, B), projs=projs, depth_values=synth.depth_planes(192, B),
                ndepths=list(ndepths), H=H, W=W)


def full_model_state(seed: int = 40) -> dict:
    """state_dict of the whole CascadeMVSNet: feature.* + cost_regularization.{0,1,2}.*"""
    sd = {"feature." + k: v for k, v in featurenet_state(seed).items()}
    for i, cin in enumerate((32, 16, 8)):
        sd.update({f"cost_regularization.{i}." + k: v for k, v in costreg_state("cas", cin=cin, base=8, seed=seed + 1 + i).items()})
    return sd


# ---- warp / cost-volume cases -------------------------------------------------------------------
def warp_plane_case(B=2, C=4, H=24, W=32, D=5, seed=1):
    fea = synth.features(2, C, H, W, seed, B)
    proj = synth.proj_matrices(2, W, seed, B)
    return dict(src_fea=fea[1], ref_fea=fea[0], ref_proj=proj[:, 0], src_proj=proj[:, 1],
                depth=synth.depth_planes(D, B))


def warp_pixel_case(B=2, C=4, H=24, W=32, D=5, seed=2):
    c = warp_plane_case(B, C, H, W, D, seed)
    c["depth"] = synth.depth_per_pixel(D, H, W, interval=21.25, batch=B)
    return c


def volume_case(n_views=4, B=1, C=8, H=16, W=24, D=8, seed=3, per_pixel=False):
    fea = synth.features(n_views, C, H, W, seed, B)
    proj = synth.proj_matrices(n_views, W, seed, B)
    depth = synth.depth_per_pixel(D, H, W, 10.6, B) if per_pixel else synth.depth_planes(D, B)
    return dict(feats=fea, proj=proj, depth=depth)


def cas_case(n_views=3, B=1, C=16, H=16, W=24, D=8, seed=4, per_pixel=True):
    fea = synth.features(n_views, C, H, W, seed, B)
    proj = synth.cas_proj_matrices(n_views, W, seed, B)
    depth = synth.depth_per_pixel(D, H, W, 10.6, B) if per_pixel else synth.depth_planes(D, B)
    return dict(feats=fea, proj=proj, depth=depth)


def cascade_case(n_views=3, B=1, H=64, W=96, ndepths=(16, 8, 8), seed=5):
    """Full 3-stage CasMVSNet from FPN-shaped features: stage k features [C_k, H/s_k, W/s_k]."""
    chans = (32, 16, 8)
    scales = (4, 2, 1)
    feats = {}
    projs = {}
    for i, (c, s) in enumerate(zip(chans, scales)):
        feats[f"stage{i + 1}"] = synth.features(n_views, c, H // s, W // s, seed + i, B)
        projs[f"stage{i + 1}"] = synth.cas_proj_matrices(n_views, W // s, seed, B)
    depth_values = synth.depth_planes(192, B)  # the dataset hands 192 planes; only [0] and [-1] are used
    return dict(feats=feats, projs=projs, depth_values=depth_values, ndepths=list(ndepths), H=H, W=W)


def cvp_case(n_src=2, B=1, C=16, H=16, W=24, D=8, seed=6, per_pixel=True):
    fea = synth.features(n_src + 1, C, H, W, seed, B)
    ref_in, src_in, ref_ex, src_ex = synth.cvp_cameras(n_src, W, seed, B)
    depth = synth.depth_per_pixel(D, H, W, 10.6, B) if per_pixel else synth.depth_planes(D, B)
    return dict(feats=fea, ref_in=ref_in, src_in=src_in, ref_ex=ref_ex, src_ex=src_ex, depth=depth)


def logits_case(B=2, D=12, H=10, W=14, seed=7, per_pixel=False):
    rng = np.random.RandomState(4000 + seed)
    logits = (3.0 * rng.standard_normal((B, D, H, W))).astype(np.float32)
    depth = synth.depth_per_pixel(D, H, W, 10.6, B) if per_pixel else synth.depth_planes(D, B)
    return dict(logits=logits, depth=depth)


def geo_case(n_src=4, H=96, W=128, seed=7):
    """Depth maps of a tilted world plane seen by a DTU-like rig (analytic per view), perturbed so that the
    geometric-consistency masks are mixed: smooth ripples everywhere, gross outliers in blocks, a zero-depth hole.
    float64 cameras [K 3x3, E 4x4] as read_camera_parameters returns them; float32 depth / confidence maps."""
    rng = np.random.RandomState(seed)
    n = n_src + 1
    f = 2892.33 * W / 1600.0
    K = np.array([[f, 0, W / 2 - 0.5 + 0.3], [0, f * 0.997, H / 2 - 0.5 - 0.2], [0, 0, 1]], np.float64)
    Ks, Es, depths = [], [], []
    normal = np.array([0.08, -0.05, 1.0]); normal /= np.linalg.norm(normal)
    c = 680.0
    for v in range(n):
        ang = rng.uniform(-0.08, 0.08, 3) if v else np.zeros(3)
        cx, sx, cy, sy, cz, sz = np.cos(ang[0]), np.sin(ang[0]), np.cos(ang[1]), np.sin(ang[1]), np.cos(ang[2]), np.sin(ang[2])
        R = (np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @
             np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]))
        t = rng.uniform(-60, 60, 3) * np.array([1, 1, 0.2]) if v else np.zeros(3)
        E = np.eye(4); E[:3, :3] = R; E[:3, 3] = t
        Kv = K.copy(); Kv[0, 2] += rng.uniform(-1, 1); Kv[1, 2] += rng.uniform(-1, 1)
        ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
        k = np.linalg.inv(Kv) @ np.stack([xs.ravel(), ys.ravel(), np.ones(H * W)])
        # plane normal . X_w = c with X_w = R^T (d k - t)  =>  d = (c + n.R^T t) / (n.R^T k)
        nr = normal @ R.T
        d = (c + nr @ t) / (nr @ k)
        d = d.reshape(H, W)
        d = d * (1 + 0.002 * np.sin(xs / 7.0 + v) * np.cos(ys / 5.0))           # ripples: borderline pixels
        Ks.append(Kv); Es.append(E); depths.append(d.astype(np.float32))
    # gross outliers / holes in the source maps, a hole in the reference map
    for v in range(1, n):
        y0, x0 = rng.randint(0, H - 24), rng.randint(0, W - 32)
        depths[v][y0:y0 + 24, x0:x0 + 32] *= np.float32(1.05)
        depths[v][rng.randint(0, H - 8):, :6] = 0
    depths[0][5:9, 10:20] = 0
    conf = rng.uniform(0.5, 1.0, (H, W)).astype(np.float32)
    return dict(K=np.stack(Ks), E=np.stack(Es), depth=np.stack(depths), conf=conf)
