"""Multi-stage drivers of the hot path: CasMVSNet's coarse-to-fine cascade and CVP-MVSNet's
pyramid, from feature maps to depth maps.  The 2D feature extractors, losses and data loading stay
in the caller (PyTorch), as BASELINE.json's north_star prescribes.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import ops
from .modules import stage_forward, cas_relative_poses
from . import _lib as L

STAGE_SCALES = {"stage1": 4, "stage2": 2, "stage3": 1}   # CasMVSNet/models/cas_mvsnet.py:86-96


def cascade_hot_path(features: Sequence[Dict[str, torch.Tensor]], proj_matrices: Dict[str, torch.Tensor],
                     depth_values: torch.Tensor, cost_regularization, ndepths=(48, 32, 8),
                     depth_interals_ratio=(4, 2, 1), img_hw=None, depth_min: Optional[float] = None,
                     depth_max: Optional[float] = None, stage_hook=None):
    """CascadeMVSNet.forward after feature extraction (CasMVSNet/models/cas_mvsnet.py:109-165).

    features: one dict per view {"stage1": [B,32,H/4,W/4], "stage2": [B,16,H/2,W/2], "stage3": [B,8,H,W]}
    proj_matrices: {"stageK": [B,N,2,4,4]};  depth_values [B,Dd] (only [0,0] and [0,-1] are used,
    exactly like the reference, cas_mvsnet.py:110-112).  `depth_min/max` may be given as Python floats
    to skip the reference's device->host read.  `stage_hook(stage_idx)`, if given, is called right before a stage first
    touches its feature maps -- a streaming caller makes the compute stream wait there for that stage's input copy, so
    the coarse stages run while the fine stage's (4x larger) features are still on the PCIe bus.
    Returns the reference's dict: per-stage {"depth","photometric_confidence"} + last stage at top level.
    """
    if img_hw is None:
        s3 = features[0]["stage%d" % len(ndepths)]
        img_hw = (s3.shape[2] * STAGE_SCALES["stage%d" % len(ndepths)], s3.shape[3] * STAGE_SCALES["stage%d" % len(ndepths)])
    H, W = img_hw
    B = depth_values.shape[0]
    if depth_min is None:
        depth_min = float(depth_values[0, 0].cpu().numpy())     # the reference's D2H sync, cas_mvsnet.py:110-111
        depth_max = float(depth_values[0, -1].cpu().numpy())
    depth_interval = (depth_max - depth_min) / depth_values.size(1)
    # relative poses of all stages and views in one batched pass (a handful of launches instead of ~50 per stage)
    keys = ["stage%d" % (i + 1) for i in range(len(ndepths))]
    reg0 = cost_regularization[0] if isinstance(cost_regularization, (list, tuple, torch.nn.ModuleList)) else cost_regularization
    fast = getattr(reg0, "mode", "strict") == "fast" and not torch.is_grad_enabled()
    rot_all, trans_all = cas_relative_poses(torch.stack([proj_matrices[k] for k in keys], 0), fused_kernel=fast)   # [S,B,N-1,9|3]
    outputs = {}
    depth = None
    for stage_idx, nd in enumerate(ndepths):
        key = "stage%d" % (stage_idx + 1)
        scale = STAGE_SCALES[key]
        if stage_hook is not None:
            stage_hook(stage_idx)
        feats = [f[key] for f in features]
        fh, fw = feats[0].shape[2:4]           # NCHW [B,C,h,w] and C8 [B,CB,h,w,8] alike
        if (fh, fw) != (H // scale, W // scale):
            raise ValueError(f"{key}: feature maps are {fh}x{fw} but img_hw={H}x{W} at scale {scale} gives "
                             f"{H // scale}x{W // scale} (the reference fails with a shape error here, cas_mvsnet.py:150)")
        if depth is not None:
            # bilinear up-sampling + get_depth_range_samples + trilinear resampling (cas_mvsnet.py:129-151), fused
            hyp = ops.cas_hypotheses(depth.detach(), (H, W), (fh, fw), nd, depth_interals_ratio[stage_idx] * depth_interval)
        else:
            # first stage: uniform planes between depth_values[:,0] and [:, -1] (module.py:509-517).  The reference
            # repeats them to [B,D,H,W] at FULL resolution (364 MB at 1600x1184) and trilinearly resamples them to the
            # stage extent (cas_mvsnet.py:150-151); resampling a plane-uniform volume at the same D returns the same
            # constants (probed bitwise), so the [B,D] planes go to the kernel directly (MVS_DEPTH_PLANE).
            lo, hi = depth_values[:, 0], depth_values[:, -1]
            step = (hi - lo) / (nd - 1)
            hyp = lo.unsqueeze(1) + torch.arange(0, nd, device=lo.device, dtype=lo.dtype).reshape(1, -1) * step.unsqueeze(1)
        reg = cost_regularization if not isinstance(cost_regularization, (list, tuple, torch.nn.ModuleList)) \
            else cost_regularization[stage_idx]
        assert len(feats) == proj_matrices[key].shape[1], "Different number of images and projection matrices"
        out = stage_forward(feats, rot_all[stage_idx], trans_all[stage_idx], hyp, reg, clamp_index=True)
        depth = out["depth"]
        outputs[key] = out
        outputs.update(out)
    return outputs
