"""nn.Module mirrors of the reference's hot-path classes, running on the sm_100a kernels.

State-dict compatibility is part of the drop-in boundary: every class below builds
the same sub-module tree as its reference counterpart (plain nn.Conv3d / ConvTranspose3d /
BatchNorm3d used as PARAMETER HOLDERS, never called), so a reference checkpoint loads with
``strict=True``.  forward() folds eval-mode BatchNorm into a per-channel (scale, shift) pair and runs
every layer as one fused conv + affine + ReLU (+ skip) kernel.

Precision modes
  "strict": fp32 NCDHW end to end (parity mode: tracks the reference's fp32 CPU path).
  "fast":   bf16 C8 activations on the tcgen05 tensor cores (throughput mode).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import ops
from . import _lib as L


# ------------------------------------------------------------------------------------------------
# parameter holders with the reference's names
# ------------------------------------------------------------------------------------------------
class ConvBnReLU3D(nn.Module):
    """Keys ``conv.weight, bn.*`` -- MVSNet/models/module.py:26-33, CVP-MVSNet/models/modules.py."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, pad=1):
        super().__init__()
        assert kernel_size == 3 and pad == 1, "the hot path only has 3x3x3 / pad 1 convolutions"
        self.conv = nn.Conv3d(in_channels, out_channels, 3, stride=stride, padding=1, bias=False)
        self.bn = nn.BatchNorm3d(out_channels)
        self.stride, self.transposed = stride, False

    def forward(self, x, skip=None, mode="strict", layout=0):
        return _run_layer(x, self.conv.weight, self.bn, self.stride, False, True, skip, mode, self, layout)


class Conv3d(nn.Module):
    """Keys ``conv.weight, bn.*`` -- CasMVSNet/models/module.py:115-157 (bn momentum 0.1)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, relu=True, bn=True, bn_momentum=0.1,
                 init_method="xavier", **kwargs):
        super().__init__()
        assert kernel_size == 3 and kwargs.get("padding", 1) == 1 and stride in (1, 2) and bn
        self.conv = nn.Conv3d(in_channels, out_channels, 3, stride=stride, bias=False, padding=1)
        self.bn = nn.BatchNorm3d(out_channels, momentum=bn_momentum)
        self.stride, self.relu = stride, relu

    def forward(self, x, skip=None, mode="strict", layout=0):
        return _run_layer(x, self.conv.weight, self.bn, self.stride, False, self.relu, skip, mode, self, layout)


class Deconv3d(nn.Module):
    """Keys ``conv.weight, bn.*`` -- CasMVSNet/models/module.py:159-200."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, relu=True, bn=True, bn_momentum=0.1,
                 init_method="xavier", **kwargs):
        super().__init__()
        assert kernel_size == 3 and kwargs.get("padding", 1) == 1 and stride in (1, 2) and bn
        assert kwargs.get("output_padding", 0) == stride - 1
        self.conv = nn.ConvTranspose3d(in_channels, out_channels, 3, stride=stride, bias=False, padding=1,
                                       output_padding=stride - 1)
        self.bn = nn.BatchNorm3d(out_channels, momentum=bn_momentum)
        self.stride, self.relu = stride, relu

    def forward(self, x, skip=None, mode="strict", layout=0):
        return _run_layer(x, self.conv.weight, self.bn, self.stride, True, self.relu, skip, mode, self, layout)


def _deconv_seq(cin, cout, stride):
    """nn.Sequential(ConvTranspose3d, BatchNorm3d, ReLU): keys ``0.weight, 1.*`` (mvsnet.py:65-78)."""
    return nn.Sequential(
        nn.ConvTranspose3d(cin, cout, kernel_size=3, padding=1, output_padding=stride - 1, stride=stride, bias=False),
        nn.BatchNorm3d(cout), nn.ReLU(inplace=True))


def _fold_bn(bn: Optional[nn.BatchNorm3d], cache_owner: nn.Module):
    """Eval-mode BN -> (scale, shift) fp32, cached until any BN tensor is modified in place or replaced."""
    if bn is None:
        return None, None
    key = tuple((t.data_ptr(), t._version) for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var))
    cached = getattr(cache_owner, "_mvs_fold", None)
    if cached is not None and cached[0] == key:
        return cached[1], cached[2]
    with torch.no_grad():
        scale = (bn.weight.double() / torch.sqrt(bn.running_var.double() + bn.eps))
        shift = bn.bias.double() - bn.running_mean.double() * scale
        scale, shift = scale.float().contiguous(), shift.float().contiguous()
    cache_owner._mvs_fold = (key, scale, shift)
    return scale, shift


def _run_layer(x, weight, bn, stride, transposed, relu, skip, mode, owner, layout=0):
    """One conv + BN + ReLU (+ skip) block.  `layout` (fast mode only): L.X_DW / L.Y_DW / L.SKIP_DW of ops.conv3d_c8."""
    if owner.training and torch.is_grad_enabled():
        # training step (BASELINE configs[3]): strict fp32 kernels under autograd, batch-statistics BatchNorm (train.py)
        from . import train
        if x.dim() != 5:
            raise L.MvsError("training runs the strict fp32 NCDHW path: build the model with mode='strict'")
        return train.train_layer(x, weight, bn, stride, transposed, relu, skip)
    if owner.training and bn is not None:
        raise L.MvsError("CostRegNet is in training mode but gradients are disabled: call .eval() for inference "
                         "(train-mode BatchNorm would use and update batch statistics)")
    scale, shift = _fold_bn(bn, owner)
    if mode == "strict":
        return ops.conv3d(x, weight, scale, shift, skip, stride, transposed, relu)
    if mode == "fast":
        cin = weight.shape[0] if transposed else weight.shape[1]
        cout = weight.shape[1] if transposed else weight.shape[0]
        return ops.conv3d_c8(x, _packed(weight, stride, transposed, owner), cin, cout, scale, shift, skip, stride,
                             transposed, relu, layout=layout)
    raise L.MvsError(f"unknown CostRegNet mode {mode!r} (use 'strict' or 'fast')")


def _packed(weight, stride, transposed, owner):
    """tcgen05 weight blocks, re-packed only when the parameter is modified or replaced."""
    key = (weight.data_ptr(), weight._version, stride, transposed)
    cached = getattr(owner, "_mvs_packed", None)
    if cached is None or cached[0] != key:
        cached = (key, ops.pack_conv_weights(weight, stride, transposed))
        owner._mvs_packed = cached
    return cached[1]


def _seq_layer(seq: nn.Sequential, x, skip, mode, layout=0):
    return _run_layer(x, seq[0].weight, seq[1], seq[0].stride[0], True, True, skip, mode, seq, layout)


def _dw_flags(mode):
    """Fast mode: the skip tensors (conv0/2/4) are written W-de-interleaved -- their only readers are the stride-2 layer
    below (even / odd staged arrays become contiguous runs) and the transposed layer that adds them (one run per output
    parity).  (y, x, skip) flags; strict mode keeps plain NCDHW."""
    return (L.Y_DW, L.X_DW, L.SKIP_DW) if mode == "fast" else (0, 0, 0)


# ------------------------------------------------------------------------------------------------
# the three CostRegNet topologies
# ------------------------------------------------------------------------------------------------
class CostRegNetMVSNet(nn.Module):
    """MVSNet/models/mvsnet.py:48-93: 32->8->16(s2)->16->32(s2)->32->64(s2)->64, three stride-2
    transposed convs with skip adds, prob conv with bias.  [B,32,D,h,w] -> [B,1,D,h,w]."""

    def __init__(self, mode="strict"):
        super().__init__()
        self.mode = mode
        self.conv0 = ConvBnReLU3D(32, 8)
        self.conv1 = ConvBnReLU3D(8, 16, stride=2)
        self.conv2 = ConvBnReLU3D(16, 16)
        self.conv3 = ConvBnReLU3D(16, 32, stride=2)
        self.conv4 = ConvBnReLU3D(32, 32)
        self.conv5 = ConvBnReLU3D(32, 64, stride=2)
        self.conv6 = ConvBnReLU3D(64, 64)
        self.conv7 = _deconv_seq(64, 32, 2)
        self.conv9 = _deconv_seq(32, 16, 2)
        self.conv11 = _deconv_seq(16, 8, 2)
        self.prob = nn.Conv3d(8, 1, 3, stride=1, padding=1)

    def forward(self, x):
        m = self.mode
        _check_divisible(x, 8)
        yd, xd, sd = _dw_flags(m)
        conv0 = self.conv0(x, mode=m, layout=yd)
        conv2 = self.conv2(self.conv1(conv0, mode=m, layout=xd), mode=m, layout=yd)
        conv4 = self.conv4(self.conv3(conv2, mode=m, layout=xd), mode=m, layout=yd)
        x = self.conv6(self.conv5(conv4, mode=m, layout=xd), mode=m)
        x = _seq_layer(self.conv7, x, conv4, m, sd)
        x = _seq_layer(self.conv9, x, conv2, m, sd)
        x = _seq_layer(self.conv11, x, conv0, m, sd)
        return _prob_layer(self.prob, x, m)


class CostRegNetCas(nn.Module):
    """CasMVSNet/models/module.py:407-438: same topology, parametrised (in_channels, base_channels),
    Conv3d/Deconv3d blocks (keys convN.conv / convN.bn), prob without bias."""

    def __init__(self, in_channels, base_channels, mode="strict"):
        super().__init__()
        self.mode = mode
        b = base_channels
        self.conv0 = Conv3d(in_channels, b, padding=1)
        self.conv1 = Conv3d(b, b * 2, stride=2, padding=1)
        self.conv2 = Conv3d(b * 2, b * 2, padding=1)
        self.conv3 = Conv3d(b * 2, b * 4, stride=2, padding=1)
        self.conv4 = Conv3d(b * 4, b * 4, padding=1)
        self.conv5 = Conv3d(b * 4, b * 8, stride=2, padding=1)
        self.conv6 = Conv3d(b * 8, b * 8, padding=1)
        self.conv7 = Deconv3d(b * 8, b * 4, stride=2, padding=1, output_padding=1)
        self.conv9 = Deconv3d(b * 4, b * 2, stride=2, padding=1, output_padding=1)
        self.conv11 = Deconv3d(b * 2, b * 1, stride=2, padding=1, output_padding=1)
        self.prob = nn.Conv3d(b, 1, 3, stride=1, padding=1, bias=False)

    def forward(self, x):
        m = self.mode
        _check_divisible(x, 8)
        yd, xd, sd = _dw_flags(m)
        conv0 = self.conv0(x, mode=m, layout=yd)
        conv2 = self.conv2(self.conv1(conv0, mode=m, layout=xd), mode=m, layout=yd)
        conv4 = self.conv4(self.conv3(conv2, mode=m, layout=xd), mode=m, layout=yd)
        x = self.conv6(self.conv5(conv4, mode=m, layout=xd), mode=m)
        x = self.conv7(x, skip=conv4, mode=m, layout=sd)
        x = self.conv9(x, skip=conv2, mode=m, layout=sd)
        x = self.conv11(x, skip=conv0, mode=m, layout=sd)
        return _prob_layer(self.prob, x, m)


class CostRegNetCVP(nn.Module):
    """CVP-MVSNet/models/net.py:52-89: 16->16->16 | 32(s2)->32->32->64->64->64 | deconv s1 64->32
    (+conv2) | deconv s2 32->16 (+conv0) | prob0 16->1 (bias); returns the squeezed [B,D,h,w]."""

    def __init__(self, mode="strict"):
        super().__init__()
        self.mode = mode
        self.conv0 = ConvBnReLU3D(16, 16)
        self.conv0a = ConvBnReLU3D(16, 16)
        self.conv1 = ConvBnReLU3D(16, 32, stride=2)
        self.conv2 = ConvBnReLU3D(32, 32)
        self.conv2a = ConvBnReLU3D(32, 32)
        self.conv3 = ConvBnReLU3D(32, 64)
        self.conv4 = ConvBnReLU3D(64, 64)
        self.conv4a = ConvBnReLU3D(64, 64)
        self.conv5 = _deconv_seq(64, 32, 1)
        self.conv6 = _deconv_seq(32, 16, 2)
        self.prob0 = nn.Conv3d(16, 1, 3, stride=1, padding=1)

    def forward(self, x):
        m = self.mode
        _check_divisible(x, 2)
        yd, xd, sd = _dw_flags(m)
        conv0 = self.conv0a(self.conv0(x, mode=m), mode=m, layout=yd)
        conv2 = self.conv2a(self.conv2(self.conv1(conv0, mode=m, layout=xd), mode=m), mode=m)
        conv4 = self.conv4a(self.conv4(self.conv3(conv2, mode=m), mode=m), mode=m)
        conv5 = _seq_layer(self.conv5, conv4, conv2, m)
        conv6 = _seq_layer(self.conv6, conv5, conv0, m, sd)
        return _prob_layer(self.prob0, conv6, m).squeeze(1)


def _check_divisible(x, k):
    dims = x.shape[2:5]           # D,H,W for both NCDHW and C8 ([B,CB,D,H,W,8]) tensors
    if any(s % k for s in dims):
        raise ValueError(f"CostRegNet needs D,H,W divisible by {k} for its skip adds "
                         f"(the reference fails at the add, mvsnet.py:89-91); got {tuple(dims)}")


def _prob_layer(conv: nn.Conv3d, x, mode):
    if conv.training and torch.is_grad_enabled() and (x.requires_grad or conv.weight.requires_grad):
        from . import train
        y = train.Conv3dFn.apply(x, conv.weight, 1, False)
        return y if conv.bias is None else y + conv.bias.view(1, -1, 1, 1, 1)
    shift = conv.bias.detach().float() if conv.bias is not None else None
    if mode == "strict":
        return ops.conv3d(x, conv.weight, None, shift, None, 1, False, False)
    if mode == "fast":
        return ops.conv3d_c8(x, _packed(conv.weight, 1, False, conv), conv.weight.shape[1], 1, None, shift, None, 1,
                             False, False)
    raise L.MvsError(f"unknown CostRegNet mode {mode!r} (use 'strict' or 'fast')")


def CostRegNet(*args, **kwargs):
    """Reference-compatible constructor: ``CostRegNet()`` -> MVSNet topology (mvsnet.py:48),
    ``CostRegNet(in_channels, base_channels)`` -> CasMVSNet topology (module.py:407)."""
    if not args and "in_channels" not in kwargs:
        return CostRegNetMVSNet(**kwargs)
    return CostRegNetCas(*args, **kwargs)


# ------------------------------------------------------------------------------------------------
# builders + per-stage forward ("DepthNet")
# ------------------------------------------------------------------------------------------------
def build_cost_volume(ref_fea, src_feas: Sequence[torch.Tensor], ref_proj, src_projs: Sequence[torch.Tensor],
                      depth_values, flags: int = 0):
    """Fused replacement of the builder loop (MVSNet/models/mvsnet.py:152-170): the N-1 warped
    volumes are never materialised.  Projections are the reference's fused 4x4 matrices."""
    rots, transs = zip(*(ops.relative_pose(sp, ref_proj) for sp in src_projs))
    return ops.cost_volume(ref_fea, list(src_feas), list(rots), list(transs), depth_values, flags)


def build_cost_volume_c8(ref_fea, src_feas: Sequence[torch.Tensor], ref_proj, src_projs: Sequence[torch.Tensor],
                         depth_values, flags: int = 0):
    """Fast-path builder: feature maps (NCHW fp32/bf16, or already C8 bf16 [B,CB,h,w,8]) -> C8 bf16
    variance volume [B,CB,D,h,w,8] for the tensor-core CostRegNet."""
    rots, transs = zip(*(ops.relative_pose(sp, ref_proj) for sp in src_projs))
    as_c8 = lambda t: t if (t.dim() == 5 and ops.is_c8(t)) else ops.pack_c8(t, ops.FAST_FEATURE_DTYPE)
    return ops.cost_volume_c8(as_c8(ref_fea), [as_c8(s) for s in src_feas], list(rots), list(transs), depth_values,
                              flags)


def stage_forward(features, rot, trans, depth_values, cost_regularization, clamp_index=True, flags=0,
                  prob_volume_init=None):
    """One cascade stage from pre-computed relative poses (rot [B,nsrc,9], trans [B,nsrc,3]):
    fused builder -> CostRegNet -> softmax/regression/confidence.  Used by the multi-stage drivers,
    which batch the 4x4 pose algebra of all views and stages into a handful of launches."""
    nsrc = len(features) - 1
    rots = [rot[:, i].contiguous() for i in range(nsrc)]
    transs = [trans[:, i].contiguous() for i in range(nsrc)]
    if _is_fast(cost_regularization):
        as_c8 = lambda t: t if (t.dim() == 5 and ops.is_c8(t)) else ops.pack_c8(t, ops.FAST_FEATURE_DTYPE)
        var = ops.cost_volume_c8(as_c8(features[0]), [as_c8(f) for f in features[1:]], rots, transs, depth_values, flags)
    else:
        var = ops.cost_volume(features[0], list(features[1:]), rots, transs, depth_values, flags)
    cost_reg = cost_regularization(var)
    logits = cost_reg.squeeze(1) if cost_reg.dim() == 5 else cost_reg
    if prob_volume_init is not None:
        logits = logits + prob_volume_init
    depth, conf = regress(logits, depth_values, clamp_index=clamp_index)
    return {"depth": depth, "photometric_confidence": conf}


def cas_relative_poses(proj_matrices, fused_kernel=False):
    """[..., N, 2, 4, 4] CasMVSNet projection blocks -> rot [..., N-1, 9], trans [..., N-1, 3] of every
    source view relative to view 0: K[:3,:3] @ E[:3,:4] (cas_mvsnet.py:30-33), src @ inverse(ref)
    (module.py:257-259), batched over all leading dimensions.  `fused_kernel` (fast path): ONE launch of
    mvs_cas_poses (fp64 inside) instead of the ~10 ATen / cuSOLVER launches below; the strict path keeps the torch ops
    so that parity tests feed identical rot / trans bits to both sides."""
    if fused_kernel and proj_matrices.is_cuda and proj_matrices.dtype == torch.float32:
        import ctypes as C
        p = proj_matrices.contiguous()
        lead, N = p.shape[:-4], p.shape[-4]
        n_sets = 1
        for d in lead:
            n_sets *= d
        rot = torch.empty((*lead, N - 1, 9), dtype=torch.float32, device=p.device)
        trans = torch.empty((*lead, N - 1, 3), dtype=torch.float32, device=p.device)
        with torch.cuda.device(p.device):
            L.check(L.lib().mvs_cas_poses(C.c_void_p(p.data_ptr()), C.c_void_p(rot.data_ptr()), C.c_void_p(trans.data_ptr()),
                                          n_sets, N, C.c_void_p(torch.cuda.current_stream().cuda_stream)), "mvs_cas_poses")
        return rot, trans
    with torch.no_grad():
        fused = proj_matrices[..., 0, :, :].clone()
        fused[..., :3, :4] = torch.matmul(proj_matrices[..., 1, :3, :3], proj_matrices[..., 0, :3, :4])
        ref_inv = torch.linalg.inv_ex(fused[..., 0, :, :]).inverse   # == torch.inverse without its device->host "singular?" sync
        prod = torch.matmul(fused[..., 1:, :, :], ref_inv.unsqueeze(-3))
        rot = prod[..., :3, :3].reshape(*prod.shape[:-2], 9).float().contiguous()
        trans = prod[..., :3, 3].float().contiguous()
    return rot, trans


def _is_fast(cost_regularization) -> bool:
    return getattr(cost_regularization, "mode", "strict") == "fast"


def regress(cost_reg, depth_values, clamp_index):
    """softmax + depth_regression + photometric confidence on [B,1,D,h,w] or [B,D,h,w] logits."""
    logits = cost_reg.squeeze(1) if cost_reg.dim() == 5 else cost_reg
    if torch.is_grad_enabled() and logits.requires_grad:
        from . import train
        return train.regress_train(logits, depth_values, clamp_index)
    depth, conf, _, _ = ops.softargmin_conf(logits, depth_values, clamp_index=clamp_index)
    return depth, conf


def mvsnet_hot_path(features: List[torch.Tensor], proj_matrices, depth_values, cost_regularization):
    """MVSNet.forward from 'step 2' to the returned dict (MVSNet/models/mvsnet.py:149-194, refine=False).
    features: list of [B,32,h,w]; proj_matrices [B,N,4,4]; depth_values [B,D]."""
    projs = torch.unbind(proj_matrices, 1)
    assert len(features) == len(projs), "Different number of images and projection matrices"
    build = build_cost_volume_c8 if _is_fast(cost_regularization) else build_cost_volume
    var = build(features[0], features[1:], projs[0], projs[1:], depth_values)
    depth, conf = regress(cost_regularization(var), depth_values, clamp_index=False)
    return {"depth": depth, "photometric_confidence": conf}


class DepthNet(nn.Module):
    """Drop-in for CasMVSNet/models/cas_mvsnet.py:8-66: one cascade stage
    (builder + CostRegNet + softmax/regression/confidence).  proj_matrices [B,N,2,4,4]."""

    def forward(self, features, proj_matrices, depth_values, num_depth, cost_regularization, prob_volume_init=None):
        proj_matrices = torch.unbind(proj_matrices, 1)
        assert len(features) == len(proj_matrices), "Different number of images and projection matrices"
        assert depth_values.shape[1] == num_depth, f"depth_values.shape[1]:{depth_values.shape[1]}  num_depth:{num_depth}"
        fused = []
        for p in proj_matrices:      # K[:3,:3] @ E[:3,:4] per view (cas_mvsnet.py:30-33)
            with torch.no_grad():
                q = p[:, 0].clone()
                q[:, :3, :4] = torch.matmul(p[:, 1, :3, :3], p[:, 0, :3, :4])
            fused.append(q)
        build = build_cost_volume_c8 if _is_fast(cost_regularization) else build_cost_volume
        var = build(features[0], features[1:], fused[0], fused[1:], depth_values)
        cost_reg = cost_regularization(var)
        logits = cost_reg.squeeze(1)
        if prob_volume_init is not None:
            logits = logits + prob_volume_init
        depth, conf = regress(logits, depth_values, clamp_index=True)
        return {"depth": depth, "photometric_confidence": conf}


def proj_cost(settings, ref_feature, src_feature, level, ref_in, src_in, ref_ex, src_ex, depth_hypos):
    """Drop-in for CVP-MVSNet/models/modules.py:221-275 (per-pixel hypotheses; reproduces the
    reference's aliasing quirk: the running sum starts from ref**2)."""
    with torch.no_grad():
        B = ref_in.shape[0]
        last = torch.tensor([[[0, 0, 0, 1.0]]], device=ref_in.device, dtype=ref_in.dtype).repeat(B, 1, 1)
        ref_proj = torch.cat((torch.matmul(ref_in, ref_ex[:, 0:3, :]), last), 1)
        src_projs = [torch.cat((torch.matmul(src_in[:, s], src_ex[:, s, 0:3, :]), last), 1) for s in range(settings.nsrc)]
    srcs = [src_feature[s][level] for s in range(settings.nsrc)]
    if getattr(settings, "mvs_mode", "strict") == "fast":
        return build_cost_volume_c8(ref_feature, srcs, ref_proj, src_projs, depth_hypos, L.REF_SUM_SQUARED)
    return build_cost_volume(ref_feature, srcs, ref_proj, src_projs, depth_hypos, L.REF_SUM_SQUARED)
