"""The caller side of the hot path: CasMVSNet's 2D FPN feature extractor and the whole
`CascadeMVSNet` model behind the reference's names, state-dict keys and call signature, so that a driver can hand the
SAME host inputs the reference takes -- images + projection matrices + depth range (cas_mvsnet.py:109-118) -- and get
the reference's output dict back.

BASELINE.json's north_star keeps the 2D extractor in PyTorch; what is built here is the hand-off:
  * `FeatureNet` (CasMVSNet/models/module.py:304-405, arch_mode "fpn") with the reference's exact sub-module tree
    (conv0.0.conv / conv0.0.bn ... out1, inner1, inner2, out2, out3) so reference checkpoints load with strict=True;
  * mode "strict": the reference's op sequence in fp32 (NCHW) -- parity mode;
  * mode "fast": ALL N views of a reference view in ONE batched call (as MVSNet_pl/models/mvsnet.py:85-88 does) instead
    of a Python loop over views (cas_mvsnet.py:115-118), fp16 C8 activations end to end, eval-mode BatchNorm folded into the
    convolution epilogue, and the stage outputs emitted directly in the builder's C8H layout [B,C/8,h,w,8] fp16 (no repack):
      - the 3x3 layers run on the tcgen05 convolution of csrc/conv3d_umma.cu with D = 1 (MVS_ACT_F16);
      - the two 5x5 / stride-2 / pad-2 layers become 3x3 / stride-1 layers over a 2x2 space-to-depth map
        (k = 2j + p + 2: tap j in {-1,0,1} of parity p in {0,1}; mvs_s2d_c8 + re-laid weights);
      - the FPN lateral steps (1x1 conv + nearest up-sampling + add, module.py:393-398) are one fused kernel each
        (mvs_fpn_merge_c8h): the 32-channel full-resolution lateral / up-sampled maps are never materialised;
    `engine="torch"` keeps the previous formulation (fp16 channels-last cuDNN) for A/B timing: 6.9 ms per cfg3 reference view
    against ~1 ms for the native engine (bench.py from_images);
  * uint8 images are accepted and normalised on the device exactly as the loader does on the host
    (CasMVSNet/datasets/general_eval.py:81-86: float32(u8) / 255): 4x fewer bytes over PCIe, identical fp32 values.
"""
from __future__ import annotations

import os
from typing import Dict, List, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, modules, ops
from .cascade import cascade_hot_path


class Conv2d(nn.Module):
    """Parameter holder with the reference's keys ``conv.weight[, conv.bias], bn.*`` (CasMVSNet/models/module.py:26-66)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, relu=True, bn=True, bn_momentum=0.1,
                 init_method="xavier", **kwargs):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, bias=(not bn), **kwargs)
        self.kernel_size, self.stride = kernel_size, stride
        self.bn = nn.BatchNorm2d(out_channels, momentum=bn_momentum) if bn else None
        self.relu = relu

    def forward(self, x):           # the reference's own sequence (strict mode / training)
        x = self.conv(x)
        if self.bn is not None:
            x = self.bn(x)
        if self.relu:
            x = F.relu(x, inplace=True)
        return x


def _fold(block: Conv2d, dtype, channels_last):
    """conv + eval BatchNorm -> (weight', bias') in `dtype`; cached until a parameter / buffer changes."""
    tensors = [block.conv.weight] + ([block.conv.bias] if block.conv.bias is not None else [])
    if block.bn is not None:
        tensors += [block.bn.weight, block.bn.bias, block.bn.running_mean, block.bn.running_var]
    key = (dtype, channels_last) + tuple((t.data_ptr(), t._version) for t in tensors)
    cached = getattr(block, "_mvs_fold", None)
    if cached is not None and cached[0] == key:
        return cached[1], cached[2]
    with torch.no_grad():
        w = block.conv.weight.double()
        b = block.conv.bias.double() if block.conv.bias is not None else torch.zeros(w.shape[0], dtype=torch.float64, device=w.device)
        if block.bn is not None:
            s = block.bn.weight.double() / torch.sqrt(block.bn.running_var.double() + block.bn.eps)
            w = w * s.view(-1, 1, 1, 1)
            b = (b - block.bn.running_mean.double()) * s + block.bn.bias.double()
        w = w.to(dtype)
        if channels_last:
            w = w.contiguous(memory_format=torch.channels_last)
        b = b.to(dtype)
    block._mvs_fold = (key, w, b)
    return w, b


def _plain(conv: nn.Conv2d, dtype, channels_last):
    key = (dtype, channels_last, conv.weight.data_ptr(), conv.weight._version,
           None if conv.bias is None else (conv.bias.data_ptr(), conv.bias._version))
    cached = getattr(conv, "_mvs_cast", None)
    if cached is not None and cached[0] == key:
        return cached[1], cached[2]
    with torch.no_grad():
        w = conv.weight.to(dtype)
        if channels_last:
            w = w.contiguous(memory_format=torch.channels_last)
        b = None if conv.bias is None else conv.bias.to(dtype)
    conv._mvs_cast = (key, w, b)
    return w, b


def _bn_affine(bn):
    if bn is None:
        return None, None
    with torch.no_grad():
        s = bn.weight.double() / torch.sqrt(bn.running_var.double() + bn.eps)
        return s.float().contiguous(), (bn.bias.double() - bn.running_mean.double() * s).float().contiguous()


def _as_3x3x3(w2d: torch.Tensor) -> torch.Tensor:
    """[Cout,Cin,k,k] (k = 1, 3 or 5-as-stride-2) -> [Cout,Cin',3,3,3] with only the kd = 1 slice populated, Cin' padded to 8:
    the 3D kernel with D = 1 then IS the 2D convolution.  k = 5 is the stride-2 / pad-2 layer over a 2x2 space-to-depth
    input: Cin' = 4*Cin, channel (py*2+px)*Cin + c, tap (jy, jx) <- original tap (2jy+py+2, 2jx+px+2)."""
    co, ci, k, _ = w2d.shape
    w2d = w2d.detach().float()
    if k == 5:
        w = torch.zeros(co, 4 * ci, 3, 3, dtype=torch.float32, device=w2d.device)
        for py in range(2):
            for px in range(2):
                for jy in range(-1, 2):
                    for jx in range(-1, 2):
                        ky, kx = 2 * jy + py + 2, 2 * jx + px + 2
                        if 0 <= ky < 5 and 0 <= kx < 5:
                            w[:, (py * 2 + px) * ci:(py * 2 + px + 1) * ci, jy + 1, jx + 1] = w2d[:, :, ky, kx]
    elif k == 3:
        w = w2d
    elif k == 1:
        w = torch.zeros(co, ci, 3, 3, dtype=torch.float32, device=w2d.device)
        w[:, :, 1, 1] = w2d[:, :, 0, 0]
    else:
        raise ValueError(k)
    cin = w.shape[1]
    cpad = (cin + 7) // 8 * 8
    w3 = torch.zeros(co, cpad, 3, 3, 3, dtype=torch.float32, device=w2d.device)
    w3[:, :cin, 1] = w
    return w3


class _FusedOut3:
    """The last FPN stage by linearity.  The reference computes (CasMVSNet/models/module.py:396-398)

        out3(up2(intra) + inner2(conv0))        up2 = nearest x2, inner2 = 1x1 conv 8 -> 32 (+ bias), out3 = 3x3 conv 32 -> 8

    through a 32-channel full-resolution map (600 MB written and re-read at 5 x 1600 x 1184).  Equivalent, without that map:
      A   = conv3x3(intra; Wc)  at HALF resolution, 32 -> 4 x 8: output pixel (2y + py, 2x + px) of out3(up2(intra)) only sees
            intra rows {y-1, y} (py = 0) or {y, y+1} (py = 1), likewise columns -- one 3x3 half-resolution kernel per parity class
            (py, px), its taps the sums of the out3 taps that fall on the same intra pixel;
      out = conv3x3(conv0; W') + shift' + A[class(h, w)][h/2][w/2]     with W' = out3 o inner2 composed (8 -> 8), the lateral bias
            pushed through out3 into shift' -- and taken out again on the image border, where the zero padding of out3 hides it.
    Weights are composed in float64 on the host and cached until a parameter changes."""

    def __init__(self, out3: nn.Conv2d, inner2: nn.Conv2d):
        self.out3, self.inner2, self.key = out3, inner2, None

    def get(self):
        ts = [self.out3.weight, self.inner2.weight] + [t for t in (self.out3.bias, self.inner2.bias) if t is not None]
        key = tuple((t.data_ptr(), t._version) for t in ts)
        if key == self.key:
            return self
        dev = self.out3.weight.device
        W3 = self.out3.weight.detach().double().cpu()                       # [Co, 32, 3, 3]
        W2 = self.inner2.weight.detach().double().cpu()[:, :, 0, 0]         # [32, Ci]
        co, cm = W3.shape[:2]
        ci = W2.shape[1]
        b2 = self.inner2.bias.detach().double().cpu() if self.inner2.bias is not None else torch.zeros(cm, dtype=torch.float64)
        b3 = self.out3.bias.detach().double().cpu() if self.out3.bias is not None else torch.zeros(co, dtype=torch.float64)
        # parity-class kernels over the half-resolution map: full-resolution tap k of class p lands on half-resolution offset
        # floor((p + k - 1) / 2)  (p = 0: k = 0 -> -1, k = 1, 2 -> 0;  p = 1: k = 0, 1 -> 0, k = 2 -> +1)
        Wc = torch.zeros(4 * co, cm, 3, 3, dtype=torch.float64)
        for py in range(2):
            for px in range(2):
                for ky in range(3):
                    for kx in range(3):
                        r, c = (py + ky - 1) // 2 + 1, (px + kx - 1) // 2 + 1
                        Wc[(py * 2 + px) * co:(py * 2 + px + 1) * co, :, r, c] += W3[:, :, ky, kx]
        Wp = torch.einsum("omhw,mc->ochw", W3, W2)                          # out3 o inner2
        tapb = torch.einsum("omhw,m->ohw", W3, b2)                          # the lateral bias seen through each out3 tap
        corr = torch.zeros(9, co, dtype=torch.float64)
        for rc in range(3):
            for cc in range(3):
                for ky in range(3):
                    for kx in range(3):
                        if (rc == 0 and ky == 0) or (rc == 2 and ky == 2) or (cc == 0 and kx == 0) or (cc == 2 and kx == 2):
                            corr[rc * 3 + cc] -= tapb[:, ky, kx]
        w3d = torch.zeros(4 * co, cm, 3, 3, 3); w3d[:, :, 1] = Wc.float()
        self.packed_a = ops.pack_conv_weights(w3d.to(dev), 1, False, act_f16=True, flat2d=True)
        w3d = torch.zeros(co, (ci + 7) // 8 * 8, 3, 3, 3); w3d[:, :ci, 1] = Wp.float()
        self.packed_b = ops.pack_conv_weights(w3d.to(dev), 1, False, act_f16=True, flat2d=True)
        self.shift = (b3 + tapb.sum(dim=(1, 2))).float().to(dev).contiguous()
        self.corr = corr.float().contiguous()
        self.co, self.cm, self.ci, self.key = co, cm, (ci + 7) // 8 * 8, key
        return self

    def __call__(self, conv0, intra):
        """conv0 [N,Ci/8,H,W,8], intra [N,Cm/8,H/2,W/2,8] (fp16 C8) -> [N,Co/8,H,W,8]."""
        F3 = self.get()
        N, _, H, W, _ = conv0.shape
        Hh, Wh = intra.shape[2:4]
        a = ops.conv3d_c8(intra.view(N, F3.cm // 8, 1, Hh, Wh, 8), F3.packed_a, F3.cm, 4 * F3.co, None, None, None, 1, False, False,
                          act_f16=True, layout=_lib.FLAT2D)
        y = ops.conv3d_c8(conv0.view(N, F3.ci // 8, 1, H, W, 8), F3.packed_b, F3.ci, F3.co, None, F3.shift, a, 1, False, False,
                          act_f16=True, layout=_lib.FLAT2D | _lib.SKIP_PS).view(N, -1, H, W, 8)
        return ops.border_add_c8h(y, F3.corr)


# Layout of the native engine's maps: True = batch-major [N,CB,H,W,8] with the flat 2D mode of the convolution kernel
# (MVS_FLAT2D: up to 15 image rows per pipeline step); False = full- / half-resolution maps folded onto the kernel's depth
# axis ([CB,N,H,W,8], MVS_KD1: N rows per step) -- the round-2 form, kept for A/B runs (MVS_FEATURE_FLAT2D=0).
FLAT2D = os.environ.get("MVS_FEATURE_FLAT2D", "1") != "0"
FUSED_OUT3 = os.environ.get("MVS_FEATURE_FUSED_OUT3", "1") != "0"      # the last FPN stage by linearity (_FusedOut3); A/B knob


class _NativeLayer:
    """Packed fp16 tcgen05 weights + folded-BN affine of one extractor layer, rebuilt when a parameter changes."""

    def __init__(self, conv: nn.Conv2d, bn, relu):
        self.conv, self.bn, self.relu, self.key = conv, bn, relu, None

    def get(self):
        ts = [self.conv.weight] + ([self.conv.bias] if self.conv.bias is not None else []) + \
            ([self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var] if self.bn is not None else [])
        key = tuple((t.data_ptr(), t._version) for t in ts)
        if key != self.key:
            w3 = _as_3x3x3(self.conv.weight)
            self.cout, self.cin = w3.shape[0], w3.shape[1]
            self.packed = ops.pack_conv_weights(w3, 1, False, act_f16=True)
            self.packed_flat = ops.pack_conv_weights(w3, 1, False, act_f16=True, flat2d=True)
            self.scale, self.shift = _bn_affine(self.bn)
            if self.conv.bias is not None:          # conv bias (no BN on such layers in the reference: it goes into the shift)
                b = self.conv.bias.detach().float()
                self.shift = b.contiguous() if self.shift is None else (self.shift + self.scale * b).contiguous()
            self.key = key
        return self

    def __call__(self, x, folded=False):
        """x [N,CB,H,W,8] fp16 -> [N,Cout/8,H,W,8]; folded: x [CB,N,H,W,8] -> [Cout/8,N,H,W,8], run as ONE volume with
        D = N (the 3D weights live in the kd = 1 slice only, so images do not mix)."""
        L = self.get()
        if folded:
            CB, N, H, W, _ = x.shape
            y = ops.conv3d_c8(x.view(1, CB, N, H, W, 8), L.packed, L.cin, L.cout, L.scale, L.shift, None, 1, False, self.relu, act_f16=True,
                              layout=_lib.KD1)
            return y.view(y.shape[1], N, H, W, 8)
        N, CB, H, W, _ = x.shape
        if FLAT2D:      # batch-major maps, one image per batch element: the kernel tiles the image rows (no depth axis)
            y = ops.conv3d_c8(x.view(N, CB, 1, H, W, 8), L.packed_flat, L.cin, L.cout, L.scale, L.shift, None, 1, False, self.relu,
                              act_f16=True, layout=_lib.FLAT2D)
        else:
            y = ops.conv3d_c8(x.view(N, CB, 1, H, W, 8), L.packed, L.cin, L.cout, L.scale, L.shift, None, 1, False, self.relu,
                              act_f16=True, layout=_lib.KD1)
        return y.view(N, y.shape[1], H, W, 8)


def to_c8h(x: torch.Tensor) -> torch.Tensor:
    """[B,C,h,w] fp16 channels-last (C % 8 == 0) -> C8H [B,C/8,h,w,8] fp16 contiguous.  C == 8: a view, no copy."""
    B, C, h, w = x.shape
    if x.dtype != torch.float16 or C % 8:
        raise ValueError("to_c8h takes fp16 maps with a multiple of 8 channels")
    nhwc = x.permute(0, 2, 3, 1)
    if not nhwc.is_contiguous():
        nhwc = nhwc.contiguous()
    if C == 8:
        return nhwc.view(B, 1, h, w, 8)
    return nhwc.view(B, h, w, C // 8, 8).permute(0, 3, 1, 2, 4).contiguous()


class FeatureNet(nn.Module):
    """CasMVSNet/models/module.py:304-405 with arch_mode="fpn" (the model's default, cas_mvsnet.py:72,100)."""

    def __init__(self, base_channels=8, num_stage=3, stride=4, arch_mode="fpn", mode="strict", engine="native"):
        super().__init__()
        if arch_mode != "fpn" or num_stage != 3:
            raise NotImplementedError("mvs_b200.FeatureNet mirrors the fpn / 3-stage extractor CascadeMVSNet constructs")
        self.arch_mode, self.stride, self.base_channels, self.num_stage, self.mode = arch_mode, stride, base_channels, num_stage, mode
        self.engine = engine            # fast mode: "native" (tcgen05 + fused FPN kernels, base_channels 8) | "torch" (cuDNN fp16)
        self._native = None
        b = base_channels
        self.conv0 = nn.Sequential(Conv2d(3, b, 3, 1, padding=1), Conv2d(b, b, 3, 1, padding=1))
        self.conv1 = nn.Sequential(Conv2d(b, b * 2, 5, stride=2, padding=2), Conv2d(b * 2, b * 2, 3, 1, padding=1),
                                   Conv2d(b * 2, b * 2, 3, 1, padding=1))
        self.conv2 = nn.Sequential(Conv2d(b * 2, b * 4, 5, stride=2, padding=2), Conv2d(b * 4, b * 4, 3, 1, padding=1),
                                   Conv2d(b * 4, b * 4, 3, 1, padding=1))
        self.out1 = nn.Conv2d(b * 4, b * 4, 1, bias=False)
        self.out_channels = [4 * b]
        final_chs = b * 4
        self.inner1 = nn.Conv2d(b * 2, final_chs, 1, bias=True)
        self.inner2 = nn.Conv2d(b * 1, final_chs, 1, bias=True)
        self.out2 = nn.Conv2d(final_chs, b * 2, 3, padding=1, bias=False)
        self.out3 = nn.Conv2d(final_chs, b, 3, padding=1, bias=False)
        self.out_channels += [b * 2, b]

    # -- strict: the reference's forward, op for op (module.py:366-405) ------------------------------------------
    def _forward_strict(self, x):
        if x.is_cuda and not self.training:      # parity mode means fp32 arithmetic: no TF32 inside cuDNN
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                return self._reference_sequence(x)
        return self._reference_sequence(x)

    def _reference_sequence(self, x):
        conv0 = self.conv0(x)
        conv1 = self.conv1(conv0)
        conv2 = self.conv2(conv1)
        intra_feat = conv2
        outputs = {"stage1": self.out1(intra_feat)}
        intra_feat = F.interpolate(intra_feat, scale_factor=2, mode="nearest") + self.inner1(conv1)
        outputs["stage2"] = self.out2(intra_feat)
        intra_feat = F.interpolate(intra_feat, scale_factor=2, mode="nearest") + self.inner2(conv0)
        outputs["stage3"] = self.out3(intra_feat)
        return outputs

    # -- fast: folded BN, fp16 channels-last -----------------------------------------------------------------------
    def _forward_fast(self, x):
        dt = torch.float16
        x = x.to(dt).contiguous(memory_format=torch.channels_last)

        def seq(blocks, t):
            for blk in blocks:
                w, b = _fold(blk, dt, True)
                t = F.conv2d(t, w, b, stride=blk.stride, padding=blk.conv.padding)
                if blk.relu:
                    t = F.relu_(t)
            return t

        conv0 = seq(self.conv0, x)
        conv1 = seq(self.conv1, conv0)
        conv2 = seq(self.conv2, conv1)
        outputs = {"stage1": F.conv2d(conv2, _plain(self.out1, dt, True)[0])}
        w, b = _plain(self.inner1, dt, True)
        intra = F.interpolate(conv2, scale_factor=2, mode="nearest") + F.conv2d(conv1, w, b)
        outputs["stage2"] = F.conv2d(intra, _plain(self.out2, dt, True)[0], padding=1)
        w, b = _plain(self.inner2, dt, True)
        intra = F.interpolate(intra, scale_factor=2, mode="nearest") + F.conv2d(conv0, w, b)
        outputs["stage3"] = F.conv2d(intra, _plain(self.out3, dt, True)[0], padding=1)
        return outputs

    # -- fast, native engine: fp16 C8 end to end on the repo's own kernels --------------------------------------------
    def _forward_native(self, x):
        if self._native is None:
            blocks = list(self.conv0) + list(self.conv1) + list(self.conv2)
            self._native = dict(blocks=[_NativeLayer(b.conv, b.bn, b.relu) for b in blocks],
                                out1=_NativeLayer(self.out1, None, False), out2=_NativeLayer(self.out2, None, False),
                                out3=_NativeLayer(self.out3, None, False), lateral=None)
        nv = self._native
        lk = tuple((t.data_ptr(), t._version) for t in (self.inner1.weight, self.inner1.bias, self.inner2.weight, self.inner2.bias))
        if nv["lateral"] is None or nv["lateral"][0] != lk:      # host copies of the 1x1 weights (kernel-parameter operands)
            with torch.no_grad():
                nv["lateral"] = (lk, [t.detach().float().cpu() for t in (self.inner1.weight, self.inner1.bias, self.inner2.weight, self.inner2.bias)])
        w1, b1, w2, b2 = nv["lateral"][1]
        L = nv["blocks"]
        # The full- and half-resolution layers run FOLDED ([CB,N,H,W,8] = one volume with D = N images on the convolution's
        # row axis); maps with one channel block have the same bytes in both layouts, the builder's inputs are batch-major.
        N = x.shape[0]
        if FLAT2D:
            conv0 = L[1](L[0](ops.img_to_c8h(x)))                                          # [N,1,H,W,8]
            conv1 = L[4](L[3](L[2](ops.s2d_c8(conv0))))                                    # [N,2,H/2,W/2,8]
            conv2 = L[7](L[6](L[5](ops.s2d_c8(conv1))))                                    # [N,4,H/4,W/4,8]
            out = {"stage1": nv["out1"](conv2)}
            intra = ops.fpn_merge_c8h(conv1, w1, b1, conv2)
            out["stage2"] = nv["out2"](intra)
            H, W = conv0.shape[2:4]
            if FUSED_OUT3 and H % 2 == 0 and W % 2 == 0 and H >= 2 and W >= 2 and tuple(intra.shape[2:4]) == (H // 2, W // 2):
                if "fused3" not in nv:
                    nv["fused3"] = _FusedOut3(self.out3, self.inner2)
                out["stage3"] = nv["fused3"](conv0, intra)                                 # never builds the 32-channel map
            else:
                out["stage3"] = nv["out3"](ops.fpn_merge_c8h(conv0, w2, b2, intra))
            return out
        t = ops.img_to_c8h(x).view(1, N, *x.shape[2:], 8)         # [1,N,H,W,8]: CB = 1, folded == batch-major
        conv0 = L[1](L[0](t, True), True)                         # [1,N,H,W,8]
        conv1 = L[4](L[3](L[2](ops.s2d_c8(conv0, True, True), True), True), True)          # [2,N,H/2,W/2,8] folded
        conv2 = L[7](L[6](L[5](ops.s2d_c8(conv1, True, False))))                           # [N,4,H/4,W/4,8] batch-major
        out = {"stage1": nv["out1"](conv2)}
        intra = ops.fpn_merge_c8h(conv1, w1, b1, conv2, x_folded=True)                     # [N,4,H/2,W/2,8]
        out["stage2"] = nv["out2"](intra)
        intra = ops.fpn_merge_c8h(conv0, w2, b2, intra, x_folded=True, out_folded=True)    # [4,N,H,W,8] folded
        out["stage3"] = nv["out3"](intra, True).view(N, 1, *x.shape[2:], 8)                # [1,N,..] == [N,1,..]
        return out                                                # C8H [N,C/8,h,w,8] fp16

    def forward(self, x, mode=None, emit_c8h=False):
        """x [B,3,H,W] float (or uint8: normalised as float32(u8)/255 like the loader).  Returns the reference's dict
        {"stage1": [B,32,H/4,W/4], "stage2": [B,16,H/2,W/2], "stage3": [B,8,H,W]}; with emit_c8h (fast mode only) the
        maps come in the builder's C8H layout [B,C/8,h,w,8] fp16."""
        mode = mode or self.mode
        native = mode == "fast" and not self.training and self.engine == "native" and x.is_cuda and self.base_channels == 8
        if x.dtype == torch.uint8 and not native:
            x = x.float() / 255.0
        if mode == "strict" or self.training:
            if emit_c8h:
                raise ValueError("emit_c8h needs mode='fast' (eval)")
            return self._forward_strict(x.float())
        if native:
            out = self._forward_native(x if x.dtype in (torch.uint8, torch.float32) else x.float())
            if emit_c8h:
                return out
            return {k: v.permute(0, 1, 4, 2, 3).reshape(v.shape[0], -1, v.shape[2], v.shape[3]) for k, v in out.items()}
        out = self._forward_fast(x)
        return {k: to_c8h(v) for k, v in out.items()} if emit_c8h else out


class CascadeMVSNet(nn.Module):
    """Drop-in for CasMVSNet/models/cas_mvsnet.py:69-165 (refine=False, share_cr supported): same constructor
    arguments, same state-dict keys (feature.*, cost_regularization.N.*), same forward signature and output dict.
    `mode`: "strict" (fp32 parity mode) or "fast" (fp16 extractor + C8H hand-off + tcgen05 CostRegNet)."""

    def __init__(self, refine=False, ndepths=(48, 32, 8), depth_interals_ratio=(4, 2, 1), share_cr=False, grad_method="detach",
                 arch_mode="fpn", cr_base_chs=(8, 8, 8), mode="strict"):
        super().__init__()
        if refine:
            raise NotImplementedError("RefineNet is not on the hot path (the reference never enables it: train.py)")
        assert len(ndepths) == len(depth_interals_ratio)
        self.refine, self.share_cr, self.grad_method, self.arch_mode = refine, share_cr, grad_method, arch_mode
        self.ndepths, self.depth_interals_ratio, self.cr_base_chs = list(ndepths), list(depth_interals_ratio), list(cr_base_chs)
        self.num_stage, self.mode = len(ndepths), mode
        self.feature = FeatureNet(base_channels=8, stride=4, num_stage=self.num_stage, arch_mode=arch_mode, mode=mode)
        if share_cr:
            raise NotImplementedError("share_cr=True needs equal stage channels, which the fpn extractor does not have")
        self.cost_regularization = nn.ModuleList([
            modules.CostRegNetCas(self.feature.out_channels[i], self.cr_base_chs[i], mode=mode) for i in range(self.num_stage)])
        self.DepthNet = modules.DepthNet()

    def extract(self, imgs: torch.Tensor) -> List[Dict[str, torch.Tensor]]:
        """imgs [B,N,3,H,W] -> one feature dict per view.  fast: one batched extractor call over all B*N images
        (view-major, so every view's maps are contiguous), C8H out; strict: the reference's loop (cas_mvsnet.py:115-118)."""
        B, N = imgs.shape[:2]
        if self.mode == "fast" and not self.training:
            flat = imgs.transpose(0, 1).reshape(N * B, *imgs.shape[2:])
            out = self.feature(flat, mode="fast", emit_c8h=True)
            return [{k: t[v * B:(v + 1) * B] for k, t in out.items()} for v in range(N)]
        return [self.feature(imgs[:, v], mode="strict") for v in range(N)]

    def forward(self, imgs, proj_matrices, depth_values, depth_min=None, depth_max=None, stage_hook=None):
        features = self.extract(imgs)
        return cascade_hot_path(features, proj_matrices, depth_values, self.cost_regularization, ndepths=self.ndepths,
                                depth_interals_ratio=self.depth_interals_ratio, img_hw=tuple(imgs.shape[-2:]),
                                depth_min=depth_min, depth_max=depth_max, stage_hook=stage_hook)
