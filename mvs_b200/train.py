"""Training step of the hot path (BASELINE.json configs[3]; SURVEY.md 8(a)-notes "Backward", 8(e)).

What the reference does per step: forward, `loss.backward()`, `optimizer.step()` (CasMVSNet/train.py:148-171,
CVP-MVSNet/train.py:184-219, MVSNet/train.py:204-227), gradients averaged over ranks by DistributedDataParallel
(CasMVSNet/train.py:367-372).  Here:

  * fused warp + variance builder: forward AND backward on the repo's kernels (`mvs_warp_variance_fwd / _bwd`, ops._CostVolumeFn);
    the sampling grid is built under no_grad in the reference (module.py:62), so only the feature maps receive gradients;
  * 3x3x3 convolutions: forward on `mvs_conv3d_fwd` (strict fp32), data gradient on the SAME kernel -- the data gradient of a
    strided convolution is the transposed convolution with the same weight tensor and vice versa (weight [Cout,Cin,3,3,3] of a
    Conv3d read as the [in,out,3,3,3] weight of a ConvTranspose3d); weight gradient: `mvs_conv3d_wgrad` (strict fp32 SIMT);
  * train-mode BatchNorm3d (batch statistics + running-stat update) + ReLU: `BnReluFn` on the streaming kernels of
    csrc/bn_train.cu (two full-grid passes each way; ATen's NCDHW batch-norm kernels run a few CTAs per channel and cost
    12.7 ms per backward call at the CVP coarse level); `MVS_TRAIN_BN=aten` restores `F.batch_norm` + `F.relu`;
  * softmax over D in ATen, depth regression through ops._DepthRegressionFn (kernel forward, analytic backward);
  * gradient all-reduce: ONE flat bucket (0.34-0.93 M parameters = 1.4-3.7 MB, latency-bound) over NCCL on a side stream,
    averaged over ranks -- `GradBucket`; the reference's DDP buckets the same tensors.
Training always runs the strict fp32 path (the fast bf16 / fp16 kernels are inference kernels).
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional

import torch
import torch.nn.functional as F

from . import _lib as L
from . import ops
from ._lib import check, lib
from .ops import _dev, _f32c, _p, _stream

import os
USE_BN_KERNELS = os.environ.get("MVS_TRAIN_BN", "native") != "aten"        # A/B knob


def conv3d_wgrad(x, grad_y, stride=1, transposed=False):
    """Weight gradient of conv3d / conv_transpose3d (kernel 3, padding 1[, output_padding stride-1]):
    x [B,Cin,D,H,W], grad_y = d loss / d y -> [Cout,Cin,3,3,3] (transposed: [Cin,Cout,3,3,3]), fp32."""
    x, grad_y = _f32c(x), _f32c(grad_y)
    _dev(x, grad_y)
    B, Cin, D, H, W = x.shape
    Cout = grad_y.shape[1]
    gw = torch.zeros((Cin, Cout, 3, 3, 3) if transposed else (Cout, Cin, 3, 3, 3), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().mvs_conv3d_wgrad(_p(x), _p(grad_y), _p(gw), B, Cin, Cout, D, H, W, stride, int(transposed), _stream()),
              "mvs_conv3d_wgrad")
    return gw


class Conv3dFn(torch.autograd.Function):
    """y = conv(x, w) (no affine, no activation) with both gradients on the repo's kernels."""

    @staticmethod
    def forward(ctx, x, weight, stride, transposed):
        x, weight = _f32c(x), _f32c(weight)
        ctx.save_for_backward(x, weight)
        ctx.stride, ctx.transposed = stride, transposed
        return ops.conv3d(x, weight, None, None, None, stride, transposed, False)

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        gy = _f32c(gy)
        gx = gw = None
        if ctx.needs_input_grad[0]:
            if ctx.stride == 2 and not ctx.transposed and any(s % 2 for s in x.shape[2:]):
                raise ValueError("stride-2 conv backward needs even D,H,W (CostRegNet's skip adds require it anyway)")
            # d/dx of a (strided) convolution = transposed convolution with the same weight tensor, and vice versa
            gx = ops.conv3d(gy, weight, None, None, None, ctx.stride, not ctx.transposed, False)
        if ctx.needs_input_grad[1]:
            gw = conv3d_wgrad(x, gy, ctx.stride, ctx.transposed)
        return gx, gw, None, None


class BnReluFn(torch.autograd.Function):
    """Train-mode BatchNorm3d (+ ReLU) on the repo's streaming kernels (csrc/bn_train.cu): batch statistics with biased variance
    for the normalisation, running statistics updated like torch (unbiased variance, momentum); analytic backward."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, momentum, eps, relu):
        x = x.contiguous()
        B, C = x.shape[:2]
        S = x[0, 0].numel()
        M = B * S
        L = lib()
        sums = torch.zeros(2 * C, dtype=torch.float64, device=x.device)
        with torch.cuda.device(x.device):
            check(L.mvs_bn_stats(_p(x), _p(sums), B, C, S, _stream()), "mvs_bn_stats")
        mean = sums[:C] / M
        var = (sums[C:] / M - mean * mean).clamp_min_(0.0)               # biased, float64
        invstd = torch.rsqrt(var + eps)
        if running_mean is not None:
            with torch.no_grad():
                running_mean.mul_(1 - momentum).add_(mean.to(running_mean.dtype), alpha=momentum)
                running_var.mul_(1 - momentum).add_((var * (M / max(M - 1, 1))).to(running_var.dtype), alpha=momentum)
        g64 = gamma.detach().double() if gamma is not None else torch.ones(C, dtype=torch.float64, device=x.device)
        b64 = beta.detach().double() if beta is not None else torch.zeros(C, dtype=torch.float64, device=x.device)
        a = (invstd * g64).float().contiguous()
        k = (b64 - mean * invstd * g64).float().contiguous()
        y = torch.empty_like(x)
        with torch.cuda.device(x.device):
            check(L.mvs_bn_apply(_p(x), _p(a), _p(k), _p(y), B, C, S, int(relu), _stream()), "mvs_bn_apply")
        ctx.save_for_backward(x, a, k, mean.float().contiguous(), invstd.float().contiguous(), (invstd * g64).float().contiguous())
        ctx.relu, ctx.dims, ctx.has_affine = bool(relu), (B, C, S, M), (gamma is not None, beta is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, a, k, mean, invstd, ga = ctx.saved_tensors
        B, C, S, M = ctx.dims
        dy = dy.contiguous()
        L = lib()
        sums = torch.zeros(2 * C, dtype=torch.float64, device=x.device)
        with torch.cuda.device(x.device):
            check(L.mvs_bn_bwd_stats(_p(x), _p(dy), _p(a), _p(k), _p(mean), _p(invstd), _p(sums), B, C, S, int(ctx.relu), _stream()),
                  "mvs_bn_bwd_stats")
        mg, mgx = (sums[:C] / M).float().contiguous(), (sums[C:] / M).float().contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            with torch.cuda.device(x.device):
                check(L.mvs_bn_bwd_apply(_p(x), _p(dy), _p(a), _p(k), _p(mean), _p(invstd), _p(ga), _p(mg), _p(mgx), _p(dx), B, C, S,
                                         int(ctx.relu), _stream()), "mvs_bn_bwd_apply")
        dgamma = sums[C:].float() if (ctx.has_affine[0] and ctx.needs_input_grad[1]) else None
        dbeta = sums[:C].float() if (ctx.has_affine[1] and ctx.needs_input_grad[2]) else None
        return dx, dgamma, dbeta, None, None, None, None, None


def bn_relu_train(x, bn, relu):
    """Batch-statistics BatchNorm3d (+ ReLU) of a conv block: the repo's kernels on CUDA fp32 tensors, ATen otherwise."""
    if x.is_cuda and x.dtype == torch.float32 and x.dim() == 5 and USE_BN_KERNELS:
        y = BnReluFn.apply(x, bn.weight, bn.bias, bn.running_mean if bn.track_running_stats else None,
                           bn.running_var if bn.track_running_stats else None, bn.momentum, bn.eps, relu)
    else:
        y = F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, True, bn.momentum, bn.eps)
        if relu:
            y = F.relu(y)
    return y


def train_layer(x, weight, bn, stride, transposed, relu, skip):
    """Conv3d / ConvTranspose3d + BatchNorm3d (batch statistics, running stats updated) + ReLU [+ skip], the training
    branch of MVSNet/models/module.py:26-33, CasMVSNet/models/module.py:115-200."""
    y = Conv3dFn.apply(x, weight, stride, transposed)
    if bn is not None:
        if bn.momentum is None:
            raise NotImplementedError("cumulative-average BatchNorm (momentum=None) is not used by the reference")
        y = bn_relu_train(y, bn, relu)
        if bn.track_running_stats and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
    elif relu:
        y = F.relu(y)
    return y if skip is None else skip + y


def regress_train(logits, depth_values, clamp_index):
    """softmax (ATen, autograd) + depth regression (kernel forward / analytic backward) + no_grad confidence."""
    prob = F.softmax(logits, dim=1)
    depth = ops.depth_regression(prob, depth_values)
    with torch.no_grad():
        _, conf, _, _ = ops.softargmin_conf(prob.detach(), depth_values.detach(), clamp_index=clamp_index, input_is_prob=True)
    return depth, conf


class GradBucket:
    """Single flat gradient bucket: pack -> one all-reduce (SUM) on a side stream -> average -> unpack.
    `reduce()` is launched right after backward and overlaps whatever the caller does before `wait()`."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=dt, device=dev)
        self.stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self.work = None

    def pack(self):
        o = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.flat[o:o + n].zero_()
            else:
                self.flat[o:o + n].copy_(p.grad.reshape(-1))
            o += n

    def unpack(self, scale: float):
        o = 0
        for p in self.params:
            n = p.numel()
            g = self.flat[o:o + n].view_as(p)
            if p.grad is None:
                p.grad = (g * scale).clone()
            else:
                p.grad.copy_(g).mul_(scale)
            o += n

    def reduce(self):
        import torch.distributed as dist
        self.pack()
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        if self.world == 1:
            return
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                self.work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=True)
        else:
            self.work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=True)

    def wait(self):
        if self.work is not None:
            self.work.wait()
            self.work = None
            if self.stream is not None:
                torch.cuda.current_stream().wait_stream(self.stream)
        self.unpack(1.0 / self.world)


def masked_smooth_l1(depth_est, depth_gt, mask):
    """mvsnet_loss / cas per-stage loss / CVP model_loss: smooth-L1 over mask (MVSNet/models/mvsnet.py:201-203)."""
    m = mask > 0.5 if mask.dtype != torch.bool else mask
    return F.smooth_l1_loss(depth_est[m], depth_gt[m], reduction="mean")


def train_step(model, optimizer, forward_loss, bucket: Optional[GradBucket] = None):
    """optimizer.zero_grad(); loss = forward_loss(model); loss.backward(); [all-reduce]; optimizer.step() -> loss tensor."""
    optimizer.zero_grad(set_to_none=False)
    loss = forward_loss(model)
    loss.backward()
    if bucket is not None:
        bucket.reduce()
        bucket.wait()
    optimizer.step()
    return loss.detach()
