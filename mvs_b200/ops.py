"""Python face of the C-ABI: the reference's operator names and signatures over device tensors.

Every function here validates its arguments, allocates the output with torch, and enqueues ONE
call into libmvs_b200.so on torch's current CUDA stream.  There is no CPU path: tensors that are
not on a CUDA device raise.  The tiny 4x4 projection algebra (inverse / matmul) stays in PyTorch
exactly where the reference has it (MVSNet/models/module.py:63), so both sides of a parity test
consume identical rot / trans bits.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib as L
from ._lib import lib, check

__all__ = [
    "relative_pose", "homo_warping", "homo_warping_cvp", "homo_warp", "warp_taps", "cost_volume",
    "cost_volume_c8", "pack_c8", "unpack_c8", "conv3d", "softargmin_conf", "depth_regression",
    "depth_regression_refine", "depth_range_samples",
]


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# Optional per-kernel CUDA-event timers (bench.py's roofline leg): events are recorded on the
# launching stream around the named launches while KERNEL_TIMERS is a dict.
KERNEL_TIMERS = None


def _tic(name):
    if KERNEL_TIMERS is None:
        return None
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    KERNEL_TIMERS.setdefault(name, []).append((a, b))
    a.record()
    return b


def _toc(ev):
    if ev is not None:
        ev.record()


def _dev(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise L.MvsError("mvs_b200 runs on CUDA tensors only (no CPU fallback); got a tensor on "
                             f"{t.device}")
    lib()


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _p(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _ptr_array(ts: Sequence[Optional[torch.Tensor]]):
    arr = (C.c_void_p * len(ts))(*[0 if t is None else t.data_ptr() for t in ts])
    return arr


def _depth_mode(depth: torch.Tensor, B: int, hw=None):
    """MVS_DEPTH_PLANE for [B,D], MVS_DEPTH_PIXEL for [B,D,H,W]; `hw` = the (H, W) extent the kernel will index the
    per-pixel hypotheses with (a mismatch would be an out-of-bounds device read; the reference raises a shape error)."""
    if depth.dim() == 2:
        if depth.shape[0] != B:
            raise ValueError(f"depth_values batch {depth.shape[0]} != {B}")
        return L.DEPTH_PLANE
    if depth.dim() == 4:
        if depth.shape[0] != B or (hw is not None and tuple(depth.shape[2:]) != tuple(hw)):
            raise ValueError(f"per-pixel depth_values must be [{B},D,{hw[0] if hw else 'H'},{hw[1] if hw else 'W'}], "
                             f"got {tuple(depth.shape)}")
        return L.DEPTH_PIXEL
    raise ValueError(f"depth_values must be [B,D] or [B,D,H,W], got {tuple(depth.shape)}")


def relative_pose(src_proj: torch.Tensor, ref_proj: torch.Tensor, ref_is_inverse: bool = False):
    """proj = src_proj @ inverse(ref_proj) -> (rot [B,9], trans [B,3]) fp32 contiguous.
    Same two torch ops as the reference (MVSNet/models/module.py:63-65)."""
    with torch.no_grad():
        proj = torch.matmul(src_proj, ref_proj if ref_is_inverse else torch.linalg.inv_ex(ref_proj).inverse)   # torch.inverse minus its D2H singularity check
        rot = proj[:, :3, :3].reshape(-1, 9).float().contiguous()
        trans = proj[:, :3, 3].float().contiguous()
    return rot, trans


def _warp(src_fea, rot, trans, depth_values, flags):
    src_fea = _f32c(src_fea)
    depth_values = _f32c(depth_values)
    _dev(src_fea, rot, trans, depth_values)
    B, Cc, H, W = src_fea.shape
    D = depth_values.shape[1]
    mode = _depth_mode(depth_values, B, (H, W))
    out = torch.empty((B, Cc, D, H, W), dtype=torch.float32, device=src_fea.device)
    with torch.cuda.device(src_fea.device):
        check(lib().mvs_warp_fwd(_p(src_fea), _p(rot), _p(trans), _p(depth_values), mode, _p(out), B, Cc, D, H, W,
                                 flags, _stream()), "mvs_warp_fwd")
    return out


def homo_warping(src_fea, src_proj, ref_proj, depth_values):
    """Drop-in for MVSNet/models/module.py:46 and CasMVSNet/models/module.py:245
    (depth_values [B,D] or [B,D,H,W]) -> [B,C,D,H,W]."""
    rot, trans = relative_pose(src_proj, ref_proj)
    return _WarpFn.apply(src_fea, rot, trans, depth_values, 0)


def homo_warping_cvp(src_feature, ref_in, src_in, ref_ex, src_ex, depth_hypos):
    """Drop-in for CVP-MVSNet/models/modules.py:81 (K and E given separately, composed as
    K @ E[:3] with a [0,0,0,1] row, modules.py:90-94)."""
    with torch.no_grad():
        last = torch.tensor([[[0, 0, 0, 1.0]]], device=src_in.device, dtype=src_in.dtype).repeat(len(src_in), 1, 1)
        src_proj = torch.cat((torch.matmul(src_in, src_ex[:, 0:3, :]), last), 1)
        ref_proj = torch.cat((torch.matmul(ref_in, ref_ex[:, 0:3, :]), last), 1)
    return homo_warping(src_feature, src_proj, ref_proj, depth_hypos)


def homo_warp(src_feat, src_proj, ref_proj_inv, depth_values):
    """Drop-in for MVSNet_pl/models/modules.py:25 (pre-inverted ref matrix, align_corners=True,
    R @ (xyz*d) + T op order)."""
    rot, trans = relative_pose(src_proj, ref_proj_inv, ref_is_inverse=True)
    return _WarpFn.apply(src_feat, rot, trans, depth_values, L.ALIGN_CORNERS | L.PL_ORDER)


def warp_taps(rot, trans, depth_values, H, W, flags=0, want_ixy=True):
    """Integer tap indices / in-bounds masks of the warp: x0, y0 int32, mask uint8 [B,D,H,W]."""
    depth_values = _f32c(depth_values)
    _dev(rot, trans, depth_values)
    B, D = depth_values.shape[0], depth_values.shape[1]
    mode = _depth_mode(depth_values, B, (H, W))
    dev = depth_values.device
    x0 = torch.empty((B, D, H, W), dtype=torch.int32, device=dev)
    y0 = torch.empty_like(x0)
    mask = torch.empty((B, D, H, W), dtype=torch.uint8, device=dev)
    ixy = torch.empty((B, D, H, W, 2), dtype=torch.float32, device=dev) if want_ixy else None
    with torch.cuda.device(dev):
        check(lib().mvs_warp_taps(_p(rot), _p(trans), _p(depth_values), mode, _p(x0), _p(y0), _p(mask), _p(ixy), B, D,
                                  H, W, flags, _stream()), "mvs_warp_taps")
    return x0, y0, mask, ixy


def _stack_pose(rots, transs):
    rot = torch.stack(rots, 1).contiguous()      # [B,nsrc,9]
    trans = torch.stack(transs, 1).contiguous()  # [B,nsrc,3]
    return rot, trans


def _cost_volume_fwd(ref, srcs, rot, trans, depth_values, flags):
    B, Cc, H, W = ref.shape
    D = depth_values.shape[1]
    mode = _depth_mode(depth_values, B, (H, W))
    nsrc = len(srcs)
    if nsrc < 1:
        raise ValueError("need at least one source view")
    if nsrc > L.MAX_SRC:
        raise ValueError(f"at most {L.MAX_SRC} source views per fused call, got {nsrc}")
    for s in srcs:
        if s.shape != ref.shape:
            raise ValueError("source feature maps must have the reference map's shape")
    out = torch.empty((B, Cc, D, H, W), dtype=torch.float32, device=ref.device)
    with torch.cuda.device(ref.device):
        ev = _tic("warp_variance")
        check(lib().mvs_warp_variance_fwd(_p(ref), _ptr_array(srcs), nsrc, _p(rot), _p(trans), _p(depth_values), mode,
                                          _p(out), B, Cc, D, H, W, flags, _stream()), "mvs_warp_variance_fwd")
        _toc(ev)
    return out


class _WarpFn(torch.autograd.Function):
    """homo_warping with the reference's gradient structure: the sampling grid is built under
    no_grad (module.py:62), so only src_fea receives a gradient."""

    @staticmethod
    def forward(ctx, src_fea, rot, trans, depth_values, flags):
        depth_values = _f32c(depth_values)
        ctx.save_for_backward(rot, trans, depth_values)
        ctx.flags = flags
        ctx.shape = tuple(src_fea.shape)
        return _warp(src_fea, rot, trans, depth_values, flags)

    @staticmethod
    def backward(ctx, grad_out):
        rot, trans, depth_values = ctx.saved_tensors
        grad_out = _f32c(grad_out)
        B, Cc, H, W = ctx.shape
        D = depth_values.shape[1]
        g_src = torch.zeros(ctx.shape, dtype=torch.float32, device=grad_out.device)
        with torch.cuda.device(grad_out.device):
            check(lib().mvs_warp_bwd(_p(grad_out), _p(rot), _p(trans), _p(depth_values), _depth_mode(depth_values, B, (H, W)),
                                     _p(g_src), B, Cc, D, H, W, ctx.flags, _stream()), "mvs_warp_bwd")
        return g_src, None, None, None, None


class _CostVolumeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ref, rot, trans, depth_values, flags, *srcs):
        ref = _f32c(ref)
        srcs = [_f32c(s) for s in srcs]
        depth_values = _f32c(depth_values)
        _dev(ref, rot, trans, depth_values, *srcs)
        ctx.save_for_backward(ref, rot, trans, depth_values, *srcs)
        ctx.flags = flags
        return _cost_volume_fwd(ref, srcs, rot, trans, depth_values, flags)

    @staticmethod
    def backward(ctx, grad_out):
        ref, rot, trans, depth_values, *srcs = ctx.saved_tensors
        grad_out = _f32c(grad_out)
        B, Cc, H, W = ref.shape
        D = depth_values.shape[1]
        g_ref = torch.zeros_like(ref)
        g_srcs = [torch.zeros_like(s) for s in srcs]
        with torch.cuda.device(ref.device):
            check(lib().mvs_warp_variance_bwd(_p(grad_out), _p(ref), _ptr_array(srcs), len(srcs), _p(rot), _p(trans),
                                              _p(depth_values), _depth_mode(depth_values, B, (H, W)), _p(g_ref),
                                              _ptr_array(g_srcs), B, Cc, D, H, W, ctx.flags, _stream()),
                  "mvs_warp_variance_bwd")
        return (g_ref, None, None, None, None, *g_srcs)


def cost_volume(ref_fea, src_feas, rots, transs, depth_values, flags=0):
    """Fused builder, strict fp32: variance over (ref, warped srcs) -> [B,C,D,H,W] fp32.
    rots/transs: per-source lists from relative_pose().  Differentiable w.r.t. the feature maps."""
    rot, trans = _stack_pose(rots, transs)
    return _CostVolumeFn.apply(ref_fea, rot, trans, depth_values, flags, *src_feas)


def pack_c8(x: torch.Tensor, dtype=torch.bfloat16) -> torch.Tensor:
    """NC(D)HW fp32/bf16 -> C8 [B, ceil(C/8), *spatial, 8] in `dtype`: bf16 (activations of the conv stack) or
    fp16 ("C8H", the feature-map format of the fast builder; values saturate at +-65504)."""
    _dev(x)
    if dtype not in (torch.bfloat16, torch.float16):
        raise ValueError("pack_c8: dtype must be torch.bfloat16 or torch.float16")
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    x = x.contiguous()
    B, Cc = x.shape[:2]
    spatial = tuple(x.shape[2:])
    inner = 1
    for s in spatial:
        inner *= s
    out = torch.empty((B, (Cc + 7) // 8, *spatial, 8), dtype=dtype, device=x.device)
    with torch.cuda.device(x.device):
        fn = lib().mvs_pack_c8 if dtype == torch.bfloat16 else lib().mvs_pack_c8h
        check(fn(_p(x), L.F32 if x.dtype == torch.float32 else L.BF16, _p(out), B, Cc, inner, _stream()), "mvs_pack_c8")
    return out


def is_c8(t: torch.Tensor) -> bool:
    """A packed C8 / C8H feature map or volume (trailing dimension of 8 sixteen-bit elements)."""
    return t.dtype in (torch.bfloat16, torch.float16) and t.dim() >= 4 and t.shape[-1] == 8


# feature-map format the fast path packs NCHW features into (MVS_C8_FEATURES=bf16 restores bf16 features + fp32 blend)
FAST_FEATURE_DTYPE = torch.bfloat16 if __import__("os").environ.get("MVS_C8_FEATURES", "f16") == "bf16" else torch.float16


def unpack_c8(x_c8: torch.Tensor, channels: int, dtype=torch.float32) -> torch.Tensor:
    """C8 bf16 [B, CB, *spatial, 8] -> NC(D)HW `dtype` (fp32 or bf16).  fp16 "C8H" feature maps are a one-way hand-off
    format of the builder and cannot be unpacked."""
    _dev(x_c8)
    if dtype not in (torch.float32, torch.bfloat16):
        raise ValueError(f"unpack_c8: dtype must be torch.float32 or torch.bfloat16, got {dtype}")
    if x_c8.dtype != torch.bfloat16 or x_c8.shape[-1] != 8 or not x_c8.is_contiguous():
        raise ValueError("unpack_c8: input must be a contiguous bf16 C8 tensor [B,CB,*spatial,8] "
                         "(fp16 C8H feature maps are not supported)")
    B = x_c8.shape[0]
    spatial = tuple(x_c8.shape[2:-1])
    inner = 1
    for s in spatial:
        inner *= s
    out = torch.empty((B, channels, *spatial), dtype=dtype, device=x_c8.device)
    with torch.cuda.device(x_c8.device):
        check(lib().mvs_unpack_c8(_p(x_c8), _p(out), L.F32 if dtype == torch.float32 else L.BF16, B, channels, inner,
                                  _stream()), "mvs_unpack_c8")
    return out


def cost_volume_c8(ref_c8, srcs_c8, rots, transs, depth_values, flags=0):
    """Fused builder, fast path: C8 feature maps [B,CB,H,W,8] (all bf16, or all fp16 = C8H) -> C8 bf16 volume
    [B,CB,D,H,W,8]."""
    if any(t.dtype != ref_c8.dtype for t in srcs_c8) or ref_c8.dtype not in (torch.bfloat16, torch.float16):
        raise ValueError("cost_volume_c8: feature maps must all be bf16 C8 or all be fp16 C8H")
    if ref_c8.dtype == torch.float16:
        flags |= L.FEAT_F16
    depth_values = _f32c(depth_values)
    rot, trans = _stack_pose(rots, transs)
    _dev(ref_c8, rot, trans, depth_values, *srcs_c8)
    B, CB, H, W, _ = ref_c8.shape
    D = depth_values.shape[1]
    mode = _depth_mode(depth_values, B, (H, W))
    nsrc = len(srcs_c8)
    if not 1 <= nsrc <= L.MAX_SRC:
        raise ValueError(f"1..{L.MAX_SRC} source views per fused call, got {nsrc}")
    out = torch.empty((B, CB, D, H, W, 8), dtype=torch.bfloat16, device=ref_c8.device)
    with torch.cuda.device(ref_c8.device):
        ev = _tic("warp_variance")
        check(lib().mvs_warp_variance_c8_fwd(_p(ref_c8), _ptr_array(srcs_c8), nsrc, _p(rot), _p(trans),
                                             _p(depth_values), mode, _p(out), B, CB * 8, D, H, W, flags, _stream()),
              "mvs_warp_variance_c8_fwd")
        _toc(ev)
    return out


def conv3d(x, weight, scale=None, shift=None, skip=None, stride=1, transposed=False, relu=False):
    """Strict fp32 NCDHW: y = [skip +] act(conv(x, w) * scale + shift); kernel 3, padding 1
    (transposed: ConvTranspose3d(stride, padding=1, output_padding=stride-1))."""
    x = _f32c(x)
    weight = _f32c(weight)
    _dev(x, weight, scale, shift, skip)
    B, Cin, D, H, W = x.shape
    Cout = weight.shape[1] if transposed else weight.shape[0]
    if (weight.shape[0] if transposed else weight.shape[1]) != Cin or tuple(weight.shape[2:]) != (3, 3, 3):
        raise ValueError(f"weight {tuple(weight.shape)} does not match Cin={Cin} / 3x3x3")
    if transposed:
        oshape = (B, Cout, D * stride, H * stride, W * stride)
    else:
        oshape = (B, Cout, (D - 1) // stride + 1, (H - 1) // stride + 1, (W - 1) // stride + 1)
    if skip is not None:
        skip = _f32c(skip)
        if tuple(skip.shape) != oshape:
            raise ValueError(f"skip {tuple(skip.shape)} != output {oshape}")
    y = torch.empty(oshape, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().mvs_conv3d_fwd(_p(x), _p(weight), _p(scale), _p(shift), _p(skip), _p(y), B, Cin, Cout, D, H, W,
                                   stride, int(transposed), L.RELU if relu else 0, _stream()), "mvs_conv3d_fwd")
    return y


def softargmin_conf(logits, depth_values, clamp_index=False, want_prob=False, want_index=False, input_is_prob=False):
    """Fused softmax over D + depth regression + photometric confidence.
    logits [B,D,H,W] -> (depth [B,H,W], conf [B,H,W], prob|None, index|None)."""
    logits = _f32c(logits)
    depth_values = _f32c(depth_values)
    _dev(logits, depth_values)
    B, D, H, W = logits.shape
    if depth_values.dim() == 1:
        depth_values = depth_values.unsqueeze(0).expand(B, D).contiguous()
    mode = _depth_mode(depth_values, B, (H, W))
    if depth_values.shape[1] != D:
        raise ValueError(f"depth_values has {depth_values.shape[1]} hypotheses, logits have {D}")
    dev = logits.device
    depth = torch.empty((B, H, W), dtype=torch.float32, device=dev)
    conf = torch.empty((B, H, W), dtype=torch.float32, device=dev)
    prob = torch.empty((B, D, H, W), dtype=torch.float32, device=dev) if want_prob else None
    index = torch.empty((B, H, W), dtype=torch.int32, device=dev) if want_index else None
    flags = (L.CLAMP_INDEX if clamp_index else 0) | (L.INPUT_IS_PROB if input_is_prob else 0)
    with torch.cuda.device(dev):
        check(lib().mvs_softargmin_conf_fwd(_p(logits), _p(depth_values), mode, _p(depth), _p(conf), _p(prob),
                                            _p(index), B, D, H, W, flags, _stream()), "mvs_softargmin_conf_fwd")
    return depth, conf, prob, index


class _DepthRegressionFn(torch.autograd.Function):
    """depth = sum_d p * depth_values with the reference's gradient structure (MVSNet/models/module.py:91-103 is plain
    autograd): d depth / d p = depth_values, d depth / d depth_values = p.  The forward is the fused kernel."""

    @staticmethod
    def forward(ctx, p, depth_values):
        ctx.save_for_backward(p, depth_values)
        return softargmin_conf(p, depth_values, input_is_prob=True)[0]

    @staticmethod
    def backward(ctx, g):
        p, dv = ctx.saved_tensors
        g = g.unsqueeze(1)                                   # [B,1,H,W]
        gp = gdv = None
        if ctx.needs_input_grad[0]:
            dvb = dv if dv.dim() == 4 else (dv.view(*dv.shape, 1, 1) if dv.dim() == 2 else dv.view(1, -1, 1, 1))
            gp = (g * dvb).to(p.dtype)
        if ctx.needs_input_grad[1]:
            gd = g * p
            gdv = gd if dv.dim() == 4 else (gd.sum((2, 3)) if dv.dim() == 2 else gd.sum((0, 2, 3)))
            gdv = gdv.to(dv.dtype)
        return gp, gdv


def _regress_prob(p, depth_values):
    if torch.is_grad_enabled() and (p.requires_grad or depth_values.requires_grad):
        return _DepthRegressionFn.apply(p, depth_values)
    return softargmin_conf(p, depth_values, input_is_prob=True)[0]


def depth_regression(p, depth_values):
    """Drop-in for depth_regression(p, depth_values): MVSNet/models/module.py:91,
    CasMVSNet/models/module.py:455 ([B,D] or [B,D,H,W]), CVP-MVSNet/models/modules.py:338, MVSNet_pl/models/modules.py:64.
    `p` is a probability volume (already soft-maxed), as in the reference.  Differentiable w.r.t. `p` and
    `depth_values` (the training path of MVSNet_pl/models/mvsnet.py:112 back-propagates through it)."""
    return _regress_prob(p, depth_values)


def depth_regression_refine(prob_volume, depth_hypothesis):
    """Drop-in for CVP-MVSNet/models/modules.py:352 (differentiable like depth_regression)."""
    return _regress_prob(prob_volume, depth_hypothesis)


def depth_range_samples(cur_depth, ndepth: int, depth_interval_pixel: float):
    """Per-pixel branch of get_depth_range_samples (CasMVSNet/models/module.py:485-504):
    cur_depth [B,H,W] -> [B,ndepth,H,W]."""
    cur_depth = _f32c(cur_depth)
    _dev(cur_depth)
    B, H, W = cur_depth.shape
    out = torch.empty((B, ndepth, H, W), dtype=torch.float32, device=cur_depth.device)
    with torch.cuda.device(cur_depth.device):
        check(lib().mvs_depth_range_samples(_p(cur_depth), float(depth_interval_pixel), ndepth, _p(out), B, H, W,
                                            _stream()), "mvs_depth_range_samples")
    return out


def cas_hypotheses(prev_depth, img_hw, stage_hw, ndepth: int, depth_interval_pixel: float):
    """Fused inter-stage step of CascadeMVSNet.forward (cas_mvsnet.py:129-151): prev_depth [B,hp,wp] ->
    per-pixel hypotheses [B,ndepth,h,w] at the stage extent (bilinear up-sampling to img_hw, +-ndepth/2
    samples, trilinear resampling) without the full-resolution intermediate volumes."""
    prev_depth = _f32c(prev_depth)
    _dev(prev_depth)
    B, hp, wp = prev_depth.shape
    (H, W), (h, w) = img_hw, stage_hw
    out = torch.empty((B, ndepth, h, w), dtype=torch.float32, device=prev_depth.device)
    with torch.cuda.device(prev_depth.device):
        check(lib().mvs_cas_hypotheses(_p(prev_depth), hp, wp, H, W, h, w, ndepth, float(depth_interval_pixel), _p(out),
                                       B, _stream()), "mvs_cas_hypotheses")
    return out


# ---- fast path: bf16 C8 convolution on the tcgen05 tensor cores -----------------------------------
def pack_conv_weights(weight: torch.Tensor, stride: int = 1, transposed: bool = False, act_f16: bool = False,
                      flat2d: bool = False) -> torch.Tensor:
    """fp32 [Cout,Cin,3,3,3] (or [Cin,Cout,3,3,3] when transposed) -> opaque packed bf16 (act_f16: fp16) blocks for
    conv3d_c8 (uint8 tensor on the weight's device).  flat2d: blocks for conv3d_c8(..., layout=L.FLAT2D) -- only the centre
    depth slice weight[:, :, 1] is used (a plain 2D convolution)."""
    weight = _f32c(weight.detach())
    _dev(weight)
    cin = weight.shape[0] if transposed else weight.shape[1]
    cout = weight.shape[1] if transposed else weight.shape[0]
    nbytes = int(lib().mvs_conv3d_c8_packed_weight_bytes(cin, cout, stride, int(transposed)))
    if nbytes <= 0:
        raise ValueError(f"unsupported layer shape Cin={cin} Cout={cout} stride={stride}")
    packed = torch.empty(nbytes, dtype=torch.uint8, device=weight.device)
    with torch.cuda.device(weight.device):
        check(lib().mvs_conv3d_c8_pack_weights_ex(_p(weight), _p(packed), cin, cout, stride, int(transposed),
                                                  (L.ACT_F16 if act_f16 else 0) | (L.FLAT2D if flat2d else 0), _stream()),
              "mvs_conv3d_c8_pack_weights")
    return packed


def conv3d_c8(x_c8, packed_w, cin: int, cout: int, scale=None, shift=None, skip_c8=None, stride=1, transposed=False,
              relu=False, act_f16=False, layout=0):
    """y = [skip +] act(conv(x, w) * scale + shift) on C8 bf16 activations [B,CB,D,H,W,8] (act_f16: fp16 activations,
    weights packed with act_f16 -- the 2D feature extractor runs on this with D = 1).
    layout: OR of L.X_DW / L.Y_DW / L.SKIP_DW -- the input / output / skip tensor keeps W de-interleaved (column w at
    (w & 1) * ceil(W / 2) + (w >> 1)); CostRegNet's fast path writes conv0/2/4 that way for their stride-2 consumers.
    Returns C8 [B,ceil(Cout/8),Do,Ho,Wo,8] in the activation dtype, or fp32 [B,1,Do,Ho,Wo] when Cout == 1 (`prob`)."""
    _dev(x_c8, packed_w, scale, shift, skip_c8)
    adt = torch.float16 if act_f16 else torch.bfloat16
    if x_c8.dtype != adt or x_c8.dim() != 6 or x_c8.shape[-1] != 8 or not x_c8.is_contiguous():
        raise ValueError(f"x_c8 must be a contiguous C8 {adt} tensor [B,CB,D,H,W,8]")
    B, CB, D, H, W, _ = x_c8.shape
    if CB != (cin + 7) // 8:
        raise ValueError(f"x_c8 has {CB} channel blocks, Cin={cin} needs {(cin + 7) // 8}")
    if transposed and stride == 2:
        Do, Ho, Wo = 2 * D, 2 * H, 2 * W
    elif stride == 2:
        Do, Ho, Wo = (D - 1) // 2 + 1, (H - 1) // 2 + 1, (W - 1) // 2 + 1
    else:
        Do, Ho, Wo = D, H, W
    if cout == 1:
        y = torch.empty((B, 1, Do, Ho, Wo), dtype=torch.float32, device=x_c8.device)
    else:
        y = torch.empty((B, (cout + 7) // 8, Do, Ho, Wo, 8), dtype=adt, device=x_c8.device)
    if skip_c8 is not None and (int(layout) & L.SKIP_PS):
        want = (B, 4 * ((cout + 7) // 8), Do, Ho // 2, Wo // 2, 8)       # pixel-shuffled half-resolution operand (flat 2D layers)
        if tuple(skip_c8.shape) != want or skip_c8.dtype != adt or not skip_c8.is_contiguous() or Ho % 2 or Wo % 2:
            raise ValueError(f"SKIP_PS: skip_c8 must be a contiguous C8 tensor {want} of the output's dtype (even H, W)")
    elif skip_c8 is not None and (skip_c8.shape != y.shape or skip_c8.dtype != adt or not skip_c8.is_contiguous()):
        raise ValueError("skip_c8 must be a contiguous C8 tensor of the output's shape and dtype")
    with torch.cuda.device(x_c8.device):
        check(lib().mvs_conv3d_c8_fwd(_p(x_c8), _p(packed_w), _p(scale), _p(shift), _p(skip_c8), _p(y), B, cin, cout, D,
                                      H, W, stride, int(transposed),
                                      (L.RELU if relu else 0) | (L.ACT_F16 if act_f16 else 0) | int(layout),
                                      _stream()), "mvs_conv3d_c8_fwd")
    return y


# ---- FeatureNet hand-off: the kernels around the extractor's 3x3 convolutions ------------------
def img_to_c8h(imgs: torch.Tensor) -> torch.Tensor:
    """[N,3,H,W] uint8 (scaled by 1/255 like the loader) or float32 -> fp16 C8 [N,1,H,W,8], channels 3..7 zero."""
    _dev(imgs)
    if imgs.dim() != 4 or imgs.shape[1] != 3 or imgs.dtype not in (torch.uint8, torch.float32):
        raise ValueError("img_to_c8h takes [N,3,H,W] uint8 or float32 images")
    imgs = imgs.contiguous()
    N, _, H, W = imgs.shape
    out = torch.empty((N, 1, H, W, 8), dtype=torch.float16, device=imgs.device)
    with torch.cuda.device(imgs.device):
        check(lib().mvs_img_to_c8h(_p(imgs), L.U8 if imgs.dtype == torch.uint8 else L.F32, _p(out), N, H, W, _stream()),
              "mvs_img_to_c8h")
    return out


def border_add_c8h(y: torch.Tensor, corr_host: torch.Tensor) -> torch.Tensor:
    """In place: y[n, c, h, w] += corr[rc * 3 + cc, c] on the one-pixel border of an fp16 C8 map [N,C/8,H,W,8] (rc / cc: 0 first,
    1 inner, 2 last row / column).  corr_host: float32 CPU tensor [9, C] (kernel-parameter operand)."""
    _dev(y)
    N, CB, H, W = _map_dims(y, False)
    if y.dtype != torch.float16:
        raise ValueError("y must be an fp16 C8 map")
    corr = corr_host.detach().to("cpu", torch.float32).reshape(9, CB * 8).contiguous()
    with torch.cuda.device(y.device):
        check(lib().mvs_border_add_c8h(_p(y), _p(corr), N, CB * 8, H, W, _stream()), "mvs_border_add_c8h")
    return y


def _map_dims(x, folded):
    """(N, CB, H, W) of a C8 map in batch-major [N,CB,H,W,8] or folded [CB,N,H,W,8] layout."""
    if x.dim() != 5 or x.shape[-1] != 8 or x.element_size() != 2 or not x.is_contiguous():
        raise ValueError("expected a contiguous C8 map [N,CB,H,W,8] (folded: [CB,N,H,W,8])")
    a, b, H, W, _ = x.shape
    return (b, a, H, W) if folded else (a, b, H, W)


def s2d_c8(x: torch.Tensor, src_folded=False, dst_folded=False) -> torch.Tensor:
    """2x2 space-to-depth of a C8 map [N,CB,H,W,8] -> [N,4*CB,ceil(H/2),ceil(W/2),8] (block order (py*2+px)*CB + cb).
    `*_folded`: the map is laid out [CB,N,H,W,8] (one volume [1,CB,D=N,H,W,8] for the convolution kernel)."""
    _dev(x)
    N, CB, H, W = _map_dims(x, src_folded)
    Ho, Wo = (H + 1) // 2, (W + 1) // 2
    out = torch.empty((4 * CB, N, Ho, Wo, 8) if dst_folded else (N, 4 * CB, Ho, Wo, 8), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().mvs_s2d_c8(_p(x), _p(out), N, CB, H, W, (1 if src_folded else 0) | (2 if dst_folded else 0), _stream()), "mvs_s2d_c8")
    return out


def fpn_merge_c8h(x: torch.Tensor, w_host: torch.Tensor, bias_host, prev=None, x_folded=False, out_folded=False,
                  prev_folded=False) -> torch.Tensor:
    """FPN lateral step: nearest_up2(prev) + conv1x1(x) + bias -> [N,4,H,W,8] fp16.  x [N,Cin/8,H,W,8] fp16 (Cin 8 | 16);
    w_host [32,Cin] / bias_host [32] float32 CPU tensors (they travel in the kernel parameter block: pass CPU copies, a
    CUDA tensor would cost a device->host sync per call); prev [N,4,Hp,Wp,8].  `*_folded`: [CB,N,H,W,8] layouts."""
    _dev(x, prev)
    if x.dtype != torch.float16:
        raise ValueError("x must be an fp16 C8 map")
    N, CB, H, W = _map_dims(x, x_folded)
    cin = CB * 8
    w_host = w_host.detach().to("cpu", torch.float32).reshape(32, cin).contiguous()
    b_host = None if bias_host is None else bias_host.detach().to("cpu", torch.float32).reshape(32).contiguous()
    Hp = Wp = 0
    if prev is not None:
        Np, CBp, Hp, Wp = _map_dims(prev, prev_folded)
        if prev.dtype != torch.float16 or (Np, CBp) != (N, 4):
            raise ValueError("prev must be an fp16 C8 map with 32 channels and x's batch")
    out = torch.empty((4, N, H, W, 8) if out_folded else (N, 4, H, W, 8), dtype=torch.float16, device=x.device)
    flags = (1 if x_folded else 0) | (2 if out_folded else 0) | (4 if prev_folded else 0)
    with torch.cuda.device(x.device):
        check(lib().mvs_fpn_merge_c8h(_p(x), _p(w_host), _p(b_host), _p(prev), _p(out), N, cin, H, W, Hp, Wp, flags, _stream()),
              "mvs_fpn_merge_c8h")
    return out
