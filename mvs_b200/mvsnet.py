"""MVSNet (the canonical family) from images: the 2D `FeatureNet` of MVSNet/models/mvsnet.py:8-45 and the `MVSNet` model of
:124-194 (refine=False, the only setting the reference's drivers use: train.py:93, eval.py:103) behind the reference's
sub-module tree / `state_dict` keys / call signature / output dict.

strict mode: the reference's op sequence in fp32 ATen (parity mode, trainable).  fast mode: the same native fp16 C8 engine
as the CasMVSNet extractor (featurenet.py): 3x3 layers on the tcgen05 convolution with the images folded onto its row axis,
the 5x5 stride-2 layers as 3x3 layers over a space-to-depth map, all N views in one batched pass, output in the builder's
C8H layout.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import featurenet, modules, ops
from .featurenet import _NativeLayer


class ConvBnReLU(nn.Module):
    """Keys ``conv.weight, bn.*`` -- MVSNet/models/module.py:6-13."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, pad=1):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=pad, bias=False)
        self.bn = nn.BatchNorm2d(out_channels)

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)), inplace=True)


class FeatureNet(nn.Module):
    """MVSNet/models/mvsnet.py:8-45: [B,3,H,W] -> [B,32,H/4,W/4]."""

    def __init__(self, mode="strict"):
        super().__init__()
        self.mode = mode
        self.inplanes = 32
        self.conv0 = ConvBnReLU(3, 8, 3, 1, 1)
        self.conv1 = ConvBnReLU(8, 8, 3, 1, 1)
        self.conv2 = ConvBnReLU(8, 16, 5, 2, 2)
        self.conv3 = ConvBnReLU(16, 16, 3, 1, 1)
        self.conv4 = ConvBnReLU(16, 16, 3, 1, 1)
        self.conv5 = ConvBnReLU(16, 32, 5, 2, 2)
        self.conv6 = ConvBnReLU(32, 32, 3, 1, 1)
        self.feature = nn.Conv2d(32, 32, 3, 1, 1)
        self._native = None

    def _reference_sequence(self, x):
        x = self.conv1(self.conv0(x))
        x = self.conv4(self.conv3(self.conv2(x)))
        return self.feature(self.conv6(self.conv5(x)))

    def _forward_native(self, x):
        if self._native is None:
            self._native = [_NativeLayer(b.conv, b.bn, True) for b in (self.conv0, self.conv1, self.conv2, self.conv3, self.conv4,
                                                                      self.conv5, self.conv6)] + [_NativeLayer(self.feature, None, False)]
        L = self._native
        N = x.shape[0]
        if featurenet.FLAT2D:                                                    # batch-major maps, flat 2D convolution mode
            c1 = L[1](L[0](ops.img_to_c8h(x)))
            c4 = L[4](L[3](L[2](ops.s2d_c8(c1))))
            return L[7](L[6](L[5](ops.s2d_c8(c4))))                              # C8H [N,4,h,w,8]
        t = ops.img_to_c8h(x).view(1, N, *x.shape[2:], 8)                       # CB = 1: folded == batch-major
        c1 = L[1](L[0](t, True), True)
        c4 = L[4](L[3](L[2](ops.s2d_c8(c1, True, True), True), True), True)      # [2,N,H/2,W/2,8] folded
        c6 = L[6](L[5](ops.s2d_c8(c4, True, False)))                             # [N,4,H/4,W/4,8] batch-major
        return L[7](c6)                                                          # C8H [N,4,h,w,8]

    def forward(self, x, mode=None, emit_c8h=False):
        """x [B,3,H,W] float32, or uint8 (normalised as float32(u8) / 255).  strict: NCHW fp32; fast: fp16, C8H on request."""
        mode = mode or self.mode
        native = mode == "fast" and not self.training and x.is_cuda
        if x.dtype == torch.uint8 and not native:
            x = x.float() / 255.0
        if not native:
            if emit_c8h:
                raise ValueError("emit_c8h needs mode='fast' (eval) on a CUDA device")
            if x.is_cuda and not self.training:
                with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                    return self._reference_sequence(x.float())
            return self._reference_sequence(x.float())
        out = self._forward_native(x if x.dtype in (torch.uint8, torch.float32) else x.float())
        return out if emit_c8h else out.permute(0, 1, 4, 2, 3).reshape(out.shape[0], -1, out.shape[2], out.shape[3])


class MVSNet(nn.Module):
    """Drop-in for MVSNet/models/mvsnet.py:124-194 with refine=False: model(imgs [B,N,3,H,W], proj_matrices [B,N,4,4],
    depth_values [B,D]) -> {"depth", "photometric_confidence"} at H/4 x W/4."""

    def __init__(self, refine=False, mode="strict"):
        super().__init__()
        if refine:
            raise NotImplementedError("RefineNet is not on the hot path (the reference runs refine=False: train.py:93)")
        self.refine, self.mode = refine, mode
        self.feature = FeatureNet(mode=mode)
        self.cost_regularization = modules.CostRegNetMVSNet(mode=mode)

    def extract(self, imgs):
        B, N = imgs.shape[:2]
        if self.mode == "fast" and not self.training:
            flat = imgs.transpose(0, 1).reshape(N * B, *imgs.shape[2:])
            out = self.feature(flat, mode="fast", emit_c8h=True)
            return [out[v * B:(v + 1) * B] for v in range(N)]
        return [self.feature(imgs[:, v], mode="strict") for v in range(N)]

    def forward(self, imgs, proj_matrices, depth_values):
        assert imgs.shape[1] == proj_matrices.shape[1], "Different number of images and projection matrices"
        return modules.mvsnet_hot_path(self.extract(imgs), proj_matrices, depth_values, self.cost_regularization)
