"""CVP-MVSNet's coarse-to-fine pyramid driven through the sm_100a kernels, from feature pyramids to the
reference's output dict (CVP-MVSNet/models/net.py:99-207 after feature extraction).

Per level: fused warp+variance builder (with the reference's aliasing quirk, net.py:129-130 /
modules.py:228-229) -> shared CostRegNet -> softmax / depth regression; between levels the small
host-side geometry of the reference stays in PyTorch: intrinsics conditioning (modules.py:29-50), the
48-plane sweep (modules.py:57-78), bicubic x2 depth up-sampling (net.py:171) and the per-batch mean
depth interval of `calDepthHypo` (modules.py:122-219) -- an fp64 two-point epipolar construction that
yields ONE scalar per batch element.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn.functional as F

from . import ops, _lib as L
from .modules import stage_forward


def condition_intrinsics(intrinsics: torch.Tensor, img_hw, level_hws) -> torch.Tensor:
    """[B,3,3] -> [B,nscale,3,3]: fx, fy, cx, cy divided by the level's down-sampling ratio
    (image height / level height), modules.py:29-50."""
    outs = []
    for (h, _w) in level_hws:
        k = intrinsics.clone()
        k[:, :2, :] = k[:, :2, :] / (img_hw[0] / h)
        outs.append(k)
    return torch.stack(outs, 1)


def sweeping_depth_hypos(depth_min, depth_max, batch: int, n: int = 48) -> torch.Tensor:
    """[B,n] planes from depth_min[0] to depth_max[0] inclusive (modules.py:57-78; every batch row uses
    element 0's range, exactly like the reference).  depth_min/max arrive as float64 tensors from the
    DataLoader collation of Python floats."""
    assert n % 2 == 0
    lo, hi = depth_min[0].double(), depth_max[0].double()
    step = (hi - lo) / (n - 1)
    planes = lo + step * torch.arange(n, dtype=torch.float64, device=lo.device)
    return planes.unsqueeze(0).repeat(batch, 1)


def _mean_depth_interval(ref_depth, ref_k, src_k, ref_e, src_e, pixel_interval=1.0) -> torch.Tensor:
    """Scalar of calDepthHypo's test branch for ONE batch element (modules.py:157-209): back-project
    every pixel at depth D and D+1, project both into the first source view, step one pixel along that
    epipolar direction, and solve for the depth change that produces it; return mean |delta_d| (fp64)."""
    H, W = ref_depth.shape
    dev = ref_depth.device
    xx, yy = torch.meshgrid(torch.arange(W, device=dev), torch.arange(H, device=dev), indexing="ij")
    px = torch.stack([xx.reshape(-1).double(), yy.reshape(-1).double(), torch.ones(W * H, device=dev, dtype=torch.float64)], 0)
    d1 = ref_depth.t().reshape(-1).double()          # column-major pixel order, like the reference
    one = torch.ones(1, W * H, device=dev, dtype=torch.float64)

    def to_src(depth):
        ray = torch.inverse(ref_k) @ (px * depth)
        world = torch.inverse(ref_e) @ torch.cat([ray, one], 0)
        cam = (src_e @ world)[:3]
        img = src_k @ cam
        z = img[2].clone()
        return img / z, z

    x1, z1 = to_src(d1)
    x2, _ = to_src(d1 + 1)
    slope = (x2[1] - x1[1]) / (x2[0] - x1[0])
    theta = torch.atan(slope)
    x3 = x1 + torch.stack([torch.cos(theta) * pixel_interval, torch.sin(theta) * pixel_interval, torch.zeros_like(theta)], 0)
    a = (ref_k @ ref_e[:3, :3]) @ torch.inverse(src_k @ src_e[:3, :3])
    t1 = z1 * (a @ x1)
    t2 = a @ x3
    m1 = torch.cat([px.t().unsqueeze(2), t2.t().unsqueeze(2)], 2)[:, 1:, :]
    m2 = t1.t()[:, 1:]
    delta = (torch.inverse(m1) @ m2.unsqueeze(2))[:, 0, 0]
    return delta.abs().mean()


def mean_depth_interval_device(ref_depths, ref_in, src_in0, ref_ex, src_ex0, pixel_interval=1.0) -> torch.Tensor:
    """The same scalar per batch element in ONE kernel launch (mvs_cvp_depth_interval, float64 per pixel like the reference)
    instead of ~40 full-image float64 torch ops and a batched 2x2 inverse per batch element.  Returns [B] float64."""
    import ctypes as C
    B, H, W = ref_depths.shape
    rk, sk, re, se = ref_in.double(), src_in0.double(), ref_ex.double(), src_ex0.double()
    a = torch.matmul(torch.matmul(rk, re[:, :3, :3]), torch.inverse(torch.matmul(sk, se[:, :3, :3])))
    cams = torch.cat([torch.inverse(rk).reshape(B, 9), torch.inverse(re).reshape(B, 16), se.reshape(B, 16), sk.reshape(B, 9),
                      a.reshape(B, 9)], 1).contiguous()
    depth = ref_depths.float().contiguous()
    out = torch.zeros(B, dtype=torch.float64, device=depth.device)
    with torch.cuda.device(depth.device):
        L.check(L.lib().mvs_cvp_depth_interval(C.c_void_p(depth.data_ptr()), C.c_void_p(cams.data_ptr()), C.c_void_p(out.data_ptr()),
                                               B, H, W, float(pixel_interval), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                "mvs_cvp_depth_interval")
    return out / float(H * W)


def depth_hypos_refine(mode: str, ref_depths, ref_in, src_in, ref_ex, src_ex, d: int = 4) -> torch.Tensor:
    """calDepthHypo (modules.py:122-219): [B,H,W] up-sampled depth -> [B,2d,H,W] per-pixel hypotheses.
    train: fixed 6.8085 interval; test: the per-batch statistical interval above."""
    B = ref_depths.shape[0]
    levels = torch.arange(-d, d, device=ref_depths.device, dtype=ref_depths.dtype).view(1, 2 * d, 1, 1)
    if mode == "train":
        interval = torch.full((B, 1, 1, 1), 6.8085, device=ref_depths.device, dtype=ref_depths.dtype)
    else:
        with torch.no_grad():
            if ref_depths.is_cuda:
                interval = mean_depth_interval_device(ref_depths, ref_in, src_in[:, 0], ref_ex, src_ex[:, 0]).float().view(B, 1, 1, 1)
            else:      # CPU callers (golden generation / oracle side): the reference's op sequence in torch float64
                vals = [_mean_depth_interval(ref_depths[b], ref_in[b].double(), src_in[b, 0].double(), ref_ex[b].double(),
                                             src_ex[b, 0].double()) for b in range(B)]
                interval = torch.stack(vals).float().view(B, 1, 1, 1)
    return (ref_depths.unsqueeze(1) + levels * interval).float()


def cvp_hot_path(ref_pyramid: Sequence[torch.Tensor], src_pyramids: Sequence[Sequence[torch.Tensor]], ref_in, src_in,
                 ref_ex, src_ex, depth_min, depth_max, cost_reg, img_hw, mode: str = "test"):
    """network.forward after the FeaturePyramid (net.py:113-207).
    ref_pyramid: [level0 (finest) .. level nscale-1 (coarsest)] of [B,16,h,w]; src_pyramids[i] likewise.
    ref_in [B,3,3], src_in [B,nsrc,3,3], ref_ex [B,4,4], src_ex [B,nsrc,4,4].
    Returns {"depth_est_list": [finest ... coarsest], "prob_confidence": [B,H,W]}."""
    nscale, nsrc = len(ref_pyramid), len(src_pyramids)
    B = ref_in.shape[0]
    hws = [tuple(f.shape[2:]) for f in ref_pyramid]
    ref_ks = condition_intrinsics(ref_in, img_hw, hws)                                   # [B,nscale,3,3]
    src_ks = torch.stack([condition_intrinsics(src_in[:, i], img_hw, hws) for i in range(nsrc)], 1)   # [B,nsrc,nscale,3,3]

    def poses(level):
        with torch.no_grad():
            last = torch.tensor([[[0, 0, 0, 1.0]]], device=ref_in.device, dtype=ref_in.dtype).repeat(B, 1, 1)
            ref_proj = torch.cat((torch.matmul(ref_ks[:, level], ref_ex[:, 0:3, :]), last), 1)
            rots, transs = [], []
            for s in range(nsrc):
                src_proj = torch.cat((torch.matmul(src_ks[:, s, level], src_ex[:, s, 0:3, :]), last), 1)
                r, t = ops.relative_pose(src_proj, ref_proj)
                rots.append(r); transs.append(t)
        return torch.stack(rots, 1), torch.stack(transs, 1)

    depth_list: List[torch.Tensor] = []
    # coarsest level: fixed fronto-parallel sweep
    lvl = nscale - 1
    hyp = sweeping_depth_hypos(depth_min, depth_max, B).to(ref_in.device).float()
    rot, trans = poses(lvl)
    out = stage_forward([ref_pyramid[lvl]] + [p[lvl] for p in src_pyramids], rot, trans, hyp, cost_reg,
                        clamp_index=False, flags=L.REF_SUM_SQUARED)
    depth = out["depth"]
    depth_list.append(depth)
    # refinement levels, coarse to fine
    for lvl in range(nscale - 2, -1, -1):
        depth_up = F.interpolate(depth[None, :], size=None, scale_factor=2, mode="bicubic", align_corners=None).squeeze(0)
        hyp = depth_hypos_refine(mode, depth_up, ref_ks[:, lvl], src_ks[:, :, lvl], ref_ex, src_ex)
        rot, trans = poses(lvl)
        out = stage_forward([ref_pyramid[lvl]] + [p[lvl] for p in src_pyramids], rot, trans, hyp, cost_reg,
                            clamp_index=False, flags=L.REF_SUM_SQUARED)
        depth = out["depth"]
        depth_list.append(depth)
    depth_list.reverse()
    return {"depth_est_list": depth_list, "prob_confidence": out["photometric_confidence"]}


# ---- the caller side: CVP-MVSNet's feature pyramid and the whole `network` (net.py:22-50, 91-207) -----------------------
def _conv_lrelu(cin, cout):
    """`conv()` of CVP-MVSNet/models/modules.py:22-26: Conv2d(3x3, bias) + LeakyReLU(0.1); keys ``0.weight, 0.bias``."""
    return torch.nn.Sequential(torch.nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=1, dilation=1, bias=True),
                               torch.nn.LeakyReLU(0.1))


class FeaturePyramid(torch.nn.Module):
    """CVP-MVSNet/models/net.py:22-50: nine shared 3x3 conv + LeakyReLU layers applied to the image and its bilinear
    half-resolution copies.  Stays PyTorch (BASELINE north_star: 2D feature extractor)."""

    def __init__(self):
        super().__init__()
        chans = [("conv0aa", 3, 64), ("conv0ba", 64, 64), ("conv0bb", 64, 64), ("conv0bc", 64, 32), ("conv0bd", 32, 32),
                 ("conv0be", 32, 32), ("conv0bf", 32, 16), ("conv0bg", 16, 16), ("conv0bh", 16, 16)]
        for name, a, b in chans:
            setattr(self, name, _conv_lrelu(a, b))
        self._order = [c[0] for c in chans]

    def _net(self, img):
        f = img
        for name in self._order:
            f = getattr(self, name)(f)
        return f

    def forward(self, img, scales=5):
        fp = [self._net(img)]
        for _ in range(scales - 1):
            img = F.interpolate(img, scale_factor=0.5, mode="bilinear", align_corners=None).detach()
            fp.append(self._net(img))
        return fp


class network(torch.nn.Module):
    """Drop-in for CVP-MVSNet/models/net.py:91-207: `network(args)` with args.nsrc / args.nscale / args.mode, the
    reference's state-dict keys (featurePyramid.*, cost_reg_refine.*), its 8-argument forward and its output dict
    {"depth_est_list": [finest .. coarsest], "prob_confidence"}.  Training (args.mode == "train", module in .train())
    runs the strict fp32 kernels under autograd (mvs_b200/train.py)."""

    def __init__(self, args, mode="strict"):
        super().__init__()
        from .modules import CostRegNetCVP
        self.featurePyramid = FeaturePyramid()
        self.cost_reg_refine = CostRegNetCVP(mode=mode)
        self.args = args

    def forward(self, ref_img, src_imgs, ref_in, src_in, ref_ex, src_ex, depth_min, depth_max):
        nscale, nsrc = self.args.nscale, self.args.nsrc
        ref_pyr = self.featurePyramid(ref_img, nscale)
        src_pyrs = [self.featurePyramid(src_imgs[:, i], nscale) for i in range(nsrc)]
        return cvp_hot_path(ref_pyr, src_pyrs, ref_in, src_in, ref_ex, src_ex, depth_min, depth_max, self.cost_reg_refine,
                            tuple(ref_img.shape[2:]), mode=getattr(self.args, "mode", "test"))
