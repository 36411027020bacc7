"""PyTorch-CPU restatement of the reference hot path, op for op (TEST / BASELINE INFRASTRUCTURE).

Why this exists next to mvs_oracle.c: the reference's arithmetic engine IS PyTorch (ATen grid_sample,
oneDNN conv3d, MKL matmul).  /root/reference cannot travel to the GPU box, so bench.py's
`cpu_baseline` leg and `--impl reference` arm time THIS port on the box's host cores -- the same
ATen kernels, with all host threads, that the reference's own modules would run there
(`cpu_baseline.kind = "port"`).  tests/test_torch_port.py checks it against the golden fixtures
generated from the unmodified reference (bit-exact for the warp / variance, since the op sequence
is identical).

Only tests/, bench.py (cpu_baseline / --impl reference) and __graft_entry__.smoke() import this.
"""
from __future__ import annotations

import warnings

import torch
import torch.nn.functional as F


def warp_volume(src_fea, src_proj, ref_proj, depth_values, align_corners=None):
    """MVSNet/models/module.py:46-87 and CasMVSNet/models/module.py:245-280 (depth [B,D] | [B,D,H,W])."""
    B, C, H, W = src_fea.shape
    D = depth_values.shape[1]
    with torch.no_grad():
        proj = src_proj @ torch.inverse(ref_proj)
        rot, trans = proj[:, :3, :3], proj[:, :3, 3:4]
        dev = src_fea.device
        ys, xs = torch.meshgrid(torch.arange(0, H, dtype=torch.float32, device=dev),
                                torch.arange(0, W, dtype=torch.float32, device=dev), indexing="ij")
        pix = torch.stack((xs.reshape(-1), ys.reshape(-1), torch.ones(H * W, device=dev)))   # [3, H*W]
        rot_xyz = torch.matmul(rot, pix.unsqueeze(0).repeat(B, 1, 1))                   # [B, 3, H*W]
        pts = rot_xyz.unsqueeze(2).repeat(1, 1, D, 1) * depth_values.view(B, 1, D, -1)  # [B, 3, D, H*W]
        pts = pts + trans.view(B, 3, 1, 1)
        uv = pts[:, :2] / pts[:, 2:3]
        gx = uv[:, 0] / ((W - 1) / 2) - 1
        gy = uv[:, 1] / ((H - 1) / 2) - 1
        grid = torch.stack((gx, gy), dim=3)                                             # [B, D, H*W, 2]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if align_corners is None:
            out = F.grid_sample(src_fea, grid.view(B, D * H, W, 2), mode="bilinear", padding_mode="zeros")
        else:
            out = F.grid_sample(src_fea, grid.view(B, D * H, W, 2), mode="bilinear", padding_mode="zeros",
                                align_corners=align_corners)
    return out.view(B, C, D, H, W)


def variance_volume(ref_fea, src_feas, ref_proj, src_projs, depth_values, ref_sum_squared=False):
    """Builder loop, eval branch: MVSNet/models/mvsnet.py:152-170 (in-place accumulation)."""
    D = depth_values.shape[1]
    n = len(src_feas) + 1
    ref_volume = ref_fea.unsqueeze(2).repeat(1, 1, D, 1, 1)
    if ref_sum_squared:   # CVP aliasing: net.py:129-130
        vol_sq = ref_volume.pow_(2)
        vol_sum = ref_volume
        for fea, prj in zip(src_feas, src_projs):
            w = warp_volume(fea, prj, ref_proj, depth_values)
            vol_sum = vol_sum + w
            vol_sq = vol_sq + w ** 2
    else:
        vol_sum = ref_volume
        vol_sq = ref_volume ** 2
        for fea, prj in zip(src_feas, src_projs):
            w = warp_volume(fea, prj, ref_proj, depth_values)
            vol_sum += w
            vol_sq += w.pow_(2)
    return vol_sq.div_(n).sub_(vol_sum.div_(n).pow_(2))


def _cbr(x, sd, conv, bn, stride=1, transposed=False, eps=1e-5):
    w = sd[conv + ".weight"]
    if transposed:
        y = F.conv_transpose3d(x, w, None, stride=stride, padding=1, output_padding=stride - 1)
    else:
        y = F.conv3d(x, w, None, stride=stride, padding=1)
    y = F.batch_norm(y, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"], sd[bn + ".bias"],
                     False, 0.1, eps)
    return F.relu(y, inplace=True)


def costreg(x, sd, family):
    """CostRegNet.forward: MVSNet/models/mvsnet.py:84-93 | CasMVSNet/models/module.py:429-438 |
    CVP-MVSNet/models/net.py:78-89, driven by a reference-keyed state dict of tensors."""
    if family == "cvp":
        c0 = _cbr(_cbr(x, sd, "conv0.conv", "conv0.bn"), sd, "conv0a.conv", "conv0a.bn")
        c2 = _cbr(_cbr(_cbr(c0, sd, "conv1.conv", "conv1.bn", 2), sd, "conv2.conv", "conv2.bn"), sd, "conv2a.conv", "conv2a.bn")
        c4 = _cbr(_cbr(_cbr(c2, sd, "conv3.conv", "conv3.bn"), sd, "conv4.conv", "conv4.bn"), sd, "conv4a.conv", "conv4a.bn")
        c5 = c2 + _cbr(c4, sd, "conv5.0", "conv5.1", 1, True)
        c6 = c0 + _cbr(c5, sd, "conv6.0", "conv6.1", 2, True)
        return F.conv3d(c6, sd["prob0.weight"], sd.get("prob0.bias"), padding=1).squeeze(1)
    dk = (lambda n: (n + ".0", n + ".1")) if family == "mvsnet" else (lambda n: (n + ".conv", n + ".bn"))
    c0 = _cbr(x, sd, "conv0.conv", "conv0.bn")
    c2 = _cbr(_cbr(c0, sd, "conv1.conv", "conv1.bn", 2), sd, "conv2.conv", "conv2.bn")
    c4 = _cbr(_cbr(c2, sd, "conv3.conv", "conv3.bn", 2), sd, "conv4.conv", "conv4.bn")
    y = _cbr(_cbr(c4, sd, "conv5.conv", "conv5.bn", 2), sd, "conv6.conv", "conv6.bn")
    y = c4 + _cbr(y, sd, *dk("conv7"), 2, True)
    y = c2 + _cbr(y, sd, *dk("conv9"), 2, True)
    y = c0 + _cbr(y, sd, *dk("conv11"), 2, True)
    return F.conv3d(y, sd["prob.weight"], sd.get("prob.bias"), padding=1)


def regress(cost_reg, depth_values, clamp_index):
    """softmax + depth_regression + photometric confidence: mvsnet.py:183-191 / cas_mvsnet.py:51-64."""
    logits = cost_reg.squeeze(1) if cost_reg.dim() == 5 else cost_reg
    p = F.softmax(logits, dim=1)
    dv = depth_values.view(*depth_values.shape, 1, 1) if depth_values.dim() <= 2 else depth_values
    depth = torch.sum(p * dv, 1)
    D = p.shape[1]
    with torch.no_grad():
        sum4 = 4 * F.avg_pool3d(F.pad(p.unsqueeze(1), pad=(0, 0, 0, 0, 1, 2)), (4, 1, 1), stride=1, padding=0).squeeze(1)
        idx = torch.sum(p * torch.arange(D, dtype=torch.float, device=p.device).view(1, D, 1, 1), 1).long()
        if clamp_index:
            idx = idx.clamp(min=0, max=D - 1)
        conf = torch.gather(sum4, 1, idx.unsqueeze(1)).squeeze(1)
    return depth, conf


def cas_fuse_proj(p):
    """K[:3,:3] @ E[:3,:4] composition per view, CasMVSNet/models/cas_mvsnet.py:30-33. p [B,2,4,4]."""
    q = p[:, 0].clone()
    q[:, :3, :4] = torch.matmul(p[:, 1, :3, :3], p[:, 0, :3, :4])
    return q


def cas_stage(feats, proj_matrices, depth_values, sd):
    """DepthNet.forward (cas_mvsnet.py:12-66). feats: list of [B,C,h,w]; proj_matrices [B,N,2,4,4]."""
    projs = [cas_fuse_proj(p) for p in torch.unbind(proj_matrices, 1)]
    var = variance_volume(feats[0], feats[1:], projs[0], projs[1:], depth_values)
    return regress(costreg(var, sd, "cas"), depth_values, clamp_index=True)


def cas_cascade(features, proj_matrices, depth_values, sds, ndepths=(48, 32, 8), ratios=(4, 2, 1), img_hw=None):
    """CascadeMVSNet.forward after feature extraction (cas_mvsnet.py:109-165)."""
    H, W = img_hw
    B = depth_values.shape[0]
    depth_min = float(depth_values[0, 0]); depth_max = float(depth_values[0, -1])
    depth_interval = (depth_max - depth_min) / depth_values.size(1)
    out, depth = {}, None
    for i, nd in enumerate(ndepths):
        key = f"stage{i + 1}"
        scale = (4, 2, 1)[i]
        if depth is None:
            lo, hi = depth_values[:, 0], depth_values[:, -1]
            step = (hi - lo) / (nd - 1)
            samples = lo.unsqueeze(1) + torch.arange(0, nd, dtype=lo.dtype, device=lo.device).reshape(1, -1) * step.unsqueeze(1)
            samples = samples.unsqueeze(-1).unsqueeze(-1).repeat(1, 1, H, W)
        else:
            cur = F.interpolate(depth.detach().unsqueeze(1), [H, W], mode="bilinear", align_corners=False).squeeze(1)
            half = nd / 2 * (ratios[i] * depth_interval)
            lo, hi = cur - half, cur + half
            step = (hi - lo) / (nd - 1)
            samples = lo.unsqueeze(1) + torch.arange(0, nd, dtype=cur.dtype, device=cur.device).reshape(1, -1, 1, 1) * step.unsqueeze(1)
        hyp = F.interpolate(samples.unsqueeze(1), [nd, H // scale, W // scale], mode="trilinear",
                            align_corners=False).squeeze(1)
        d, c = cas_stage([f[key] for f in features], proj_matrices[key], hyp, sds[i])
        depth = d
        out[key] = {"depth": d, "photometric_confidence": c}
    out.update(out[f"stage{len(ndepths)}"])
    return out


# ---- the caller side: CasMVSNet's FPN extractor and the whole model from images ----------------
def _c2d(x, sd, name, stride=1, pad=1, eps=1e-5):
    """Conv2d block = conv (no bias) + BatchNorm2d (eval) + ReLU: CasMVSNet/models/module.py:26-66."""
    y = F.conv2d(x, sd[name + ".conv.weight"], None, stride=stride, padding=pad)
    y = F.batch_norm(y, sd[name + ".bn.running_mean"], sd[name + ".bn.running_var"], sd[name + ".bn.weight"],
                     sd[name + ".bn.bias"], False, 0.1, eps)
    return F.relu(y, inplace=True)


def featurenet(x, sd, prefix=""):
    """FeatureNet.forward, arch_mode="fpn", num_stage=3: CasMVSNet/models/module.py:366-405.  x [B,3,H,W]."""
    p = prefix
    conv0 = _c2d(_c2d(x, sd, p + "conv0.0"), sd, p + "conv0.1")
    conv1 = _c2d(_c2d(_c2d(conv0, sd, p + "conv1.0", 2, 2), sd, p + "conv1.1"), sd, p + "conv1.2")
    conv2 = _c2d(_c2d(_c2d(conv1, sd, p + "conv2.0", 2, 2), sd, p + "conv2.1"), sd, p + "conv2.2")
    intra = conv2
    out = {"stage1": F.conv2d(intra, sd[p + "out1.weight"])}
    intra = F.interpolate(intra, scale_factor=2, mode="nearest") + F.conv2d(conv1, sd[p + "inner1.weight"], sd[p + "inner1.bias"])
    out["stage2"] = F.conv2d(intra, sd[p + "out2.weight"], padding=1)
    intra = F.interpolate(intra, scale_factor=2, mode="nearest") + F.conv2d(conv0, sd[p + "inner2.weight"], sd[p + "inner2.bias"])
    out["stage3"] = F.conv2d(intra, sd[p + "out3.weight"], padding=1)
    return out


def cas_model(imgs, proj_matrices, depth_values, sd, ndepths=(48, 32, 8), ratios=(4, 2, 1)):
    """CascadeMVSNet.forward from images (cas_mvsnet.py:109-165): per-view extractor loop (:115-118) + the cascade.
    imgs [B,N,3,H,W] float32; sd = the whole model's state dict (feature.*, cost_regularization.N.*)."""
    features = [featurenet(imgs[:, v], sd, "feature.") for v in range(imgs.shape[1])]
    sds = [{k[len(f"cost_regularization.{i}."):]: v for k, v in sd.items() if k.startswith(f"cost_regularization.{i}.")}
           for i in range(len(ndepths))]
    return cas_cascade(features, proj_matrices, depth_values, sds, ndepths=ndepths, ratios=ratios, img_hw=tuple(imgs.shape[-2:]))


# ---- CVP-MVSNet: the whole `network` (feature pyramid + coarse-to-fine hot path), eval or training -------------------
def _cvp_feature(img, sd, p):
    """FeaturePyramid's shared nine-layer CNN (CVP-MVSNet/models/net.py:28-43; `conv()` = Conv2d + LeakyReLU(0.1))."""
    f = img
    for name in ("conv0aa", "conv0ba", "conv0bb", "conv0bc", "conv0bd", "conv0be", "conv0bf", "conv0bg", "conv0bh"):
        f = F.leaky_relu(F.conv2d(f, sd[f"{p}{name}.0.weight"], sd[f"{p}{name}.0.bias"], padding=1), 0.1)
    return f


def cvp_feature_pyramid(img, sd, scales, p="featurePyramid."):
    """net.py:39-50: the image and its bilinear half-resolution copies through the same CNN; finest level first."""
    fp = [_cvp_feature(img, sd, p)]
    for _ in range(scales - 1):
        img = F.interpolate(img, scale_factor=0.5, mode="bilinear", align_corners=None).detach()
        fp.append(_cvp_feature(img, sd, p))
    return fp


def _cvp_cbr(x, sd, conv, bn, stride, transposed, train):
    w = sd[conv + ".weight"]
    y = (F.conv_transpose3d(x, w, None, stride=stride, padding=1, output_padding=stride - 1) if transposed
         else F.conv3d(x, w, None, stride=stride, padding=1))
    if train:
        y = F.batch_norm(y, None, None, sd[bn + ".weight"], sd[bn + ".bias"], True, 0.1, 1e-5)
    else:
        y = F.batch_norm(y, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"], sd[bn + ".bias"], False, 0.1, 1e-5)
    return F.relu(y)


def cvp_costreg(x, sd, p="cost_reg_refine.", train=False):
    """CVP CostRegNet.forward (net.py:78-89), eval or train-mode BatchNorm."""
    c = lambda t, n, s=1: _cvp_cbr(t, sd, f"{p}{n}.conv", f"{p}{n}.bn", s, False, train)
    c0 = c(c(x, "conv0"), "conv0a")
    c2 = c(c(c(c0, "conv1", 2), "conv2"), "conv2a")
    c4 = c(c(c(c2, "conv3"), "conv4"), "conv4a")
    c5 = c2 + _cvp_cbr(c4, sd, f"{p}conv5.0", f"{p}conv5.1", 1, True, train)
    c6 = c0 + _cvp_cbr(c5, sd, f"{p}conv6.0", f"{p}conv6.1", 2, True, train)
    return F.conv3d(c6, sd[f"{p}prob0.weight"], sd[f"{p}prob0.bias"], padding=1).squeeze(1)


def cvp_network(ref_img, src_imgs, ref_in, src_in, ref_ex, src_ex, depth_min, depth_max, sd, nscale=2, train=True):
    """network.forward (net.py:99-207) with args.mode = "train" | "test" semantics for the hypotheses (train: the fixed
    6.8085 interval of modules.py:134-143).  Returns depth_est_list, finest first.  Out-of-place accumulation (the training
    branch, net.py:141-143); the aliasing quirk (sum starts from ref**2, net.py:129-130 / modules.py:228-229) is kept."""
    nsrc = src_imgs.shape[1]
    B = ref_img.shape[0]
    ref_p = cvp_feature_pyramid(ref_img, sd, nscale)
    src_p = [cvp_feature_pyramid(src_imgs[:, i], sd, nscale) for i in range(nsrc)]
    H = ref_img.shape[2]

    def cond(k, level):                                         # conditionIntrinsics, modules.py:29-50
        k = k.clone()
        k[:, :2, :] = k[:, :2, :] / (H / ref_p[level].shape[2])
        return k

    def proj(k, e):
        last = torch.tensor([[[0, 0, 0, 1.0]]], device=k.device, dtype=k.dtype).repeat(B, 1, 1)
        return torch.cat((torch.matmul(k, e[:, 0:3, :]), last), 1)

    def volume(level, hyp):
        D = hyp.shape[1]
        r2 = ref_p[level].unsqueeze(2).repeat(1, 1, D, 1, 1) ** 2
        vs, vq = r2, r2
        rp = proj(cond(ref_in, level), ref_ex)
        for i in range(nsrc):
            w = warp_volume(src_p[i][level], proj(cond(src_in[:, i], level), src_ex[:, i]), rp, hyp)
            vs = vs + w
            vq = vq + w ** 2
        return vq / (nsrc + 1) - (vs / (nsrc + 1)) ** 2

    lo, hi = depth_min[0].double(), depth_max[0].double()
    planes = (lo + (hi - lo) / 47 * torch.arange(48, dtype=torch.float64, device=ref_img.device)).float().unsqueeze(0).repeat(B, 1)
    depths = []
    p = F.softmax(cvp_costreg(volume(nscale - 1, planes), sd, train=train), 1)
    depth = torch.sum(p * planes.view(B, 48, 1, 1), 1)
    depths.append(depth)
    for level in range(nscale - 2, -1, -1):
        up = F.interpolate(depth[None, :], size=None, scale_factor=2, mode="bicubic", align_corners=None).squeeze(0)
        assert train, "the test-mode statistical interval of calDepthHypo is not ported (mvs_b200/pyramid.py has it)"
        hyp = up.unsqueeze(1) + torch.arange(-4, 4, device=up.device, dtype=up.dtype).view(1, 8, 1, 1) * 6.8085
        p = F.softmax(cvp_costreg(volume(level, hyp), sd, train=train), 1)
        depth = torch.sum(p * hyp, 1)
        depths.append(depth)
    depths.reverse()
    return depths


# ---- MVSNet from images (MVSNet/models/mvsnet.py:8-45, 136-194) ----------------------------------------------------------
def mvsnet_featurenet(x, sd, p="feature."):
    def cbr(t, n, stride=1, pad=1):
        y = F.conv2d(t, sd[f"{p}{n}.conv.weight"], None, stride=stride, padding=pad)
        y = F.batch_norm(y, sd[f"{p}{n}.bn.running_mean"], sd[f"{p}{n}.bn.running_var"], sd[f"{p}{n}.bn.weight"], sd[f"{p}{n}.bn.bias"],
                         False, 0.1, 1e-5)
        return F.relu(y, inplace=True)
    x = cbr(cbr(x, "conv0"), "conv1")
    x = cbr(cbr(cbr(x, "conv2", 2, 2), "conv3"), "conv4")
    x = cbr(cbr(x, "conv5", 2, 2), "conv6")
    return F.conv2d(x, sd[p + "feature.weight"], sd[p + "feature.bias"], padding=1)


def mvsnet_model(imgs, proj_matrices, depth_values, sd):
    """MVSNet.forward, refine=False, eval branch.  imgs [B,N,3,H,W]; proj_matrices [B,N,4,4]; depth_values [B,D]."""
    feats = [mvsnet_featurenet(imgs[:, v], sd) for v in range(imgs.shape[1])]
    projs = torch.unbind(proj_matrices, 1)
    var = variance_volume(feats[0], feats[1:], projs[0], projs[1:], depth_values)
    csd = {k[len("cost_regularization."):]: v for k, v in sd.items() if k.startswith("cost_regularization.")}
    depth, conf = regress(costreg(var, csd, "mvsnet"), depth_values, clamp_index=False)
    return {"depth": depth, "photometric_confidence": conf}
