"""Training step of the hot path (BASELINE.json configs[3]; SURVEY.md 8(a)-notes "Backward", 8(e)).

What the reference does per step: forward, `loss.backward()`, `optimizer.step()` (CasMVSNet/train.py:148-171,
CVP-MVSNet/train.py:184-219, MVSNet/train.py:204-227), gradients averaged over ranks by DistributedDataParallel
(CasMVSNet/train.py:367-372).  Here:

  * fused warp + variance builder: forward AND backward on the repo's kernels (`mvs_warp_variance_fwd / _bwd`, ops._CostVolumeFn);
    the sampling grid is built under no_grad in the reference (module.py:62), so only the feature maps receive gradients;
  * 3x3x3 convolutions: forward on `mvs_conv3d_fwd` (strict fp32), data gradient on the SAME kernel -- the data gradient of a
    strided convolution is the transposed convolution with the same weight tensor and vice versa (weight [Cout,Cin,3,3,3] of a
    Conv3d read as the [in,out,3,3,3] weight of a ConvTranspose3d); weight gradient: `mvs_conv3d_wgrad` (strict fp32 SIMT);
  * train-mode BatchNorm3d (batch statistics + running-stat update) and ReLU stay ATen ops on the conv output
    (`F.batch_norm(training=True)`): the gap to the fused eval epilogue is stated in DESIGN.md;
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


def train_layer(x, weight, bn, stride, transposed, relu, skip):
    """Conv3d / ConvTranspose3d + BatchNorm3d (batch statistics, running stats updated) + ReLU [+ skip], the training
    branch of MVSNet/models/module.py:26-33, CasMVSNet/models/module.py:115-200."""
    y = Conv3dFn.apply(x, weight, stride, transposed)
    if bn is not None:
        if bn.momentum is None:
            raise NotImplementedError("cumulative-average BatchNorm (momentum=None) is not used by the reference")
        y = F.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, True, bn.momentum, bn.eps)
        if bn.track_running_stats and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
    if relu:
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
