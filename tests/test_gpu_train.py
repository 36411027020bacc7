"""Training step on the repo's kernels (BASELINE configs[3]): gradients against torch autograd through the reference's op
sequence (ATen fp32, TF32 off) -- conv / deconv data + weight gradients, train-mode BatchNorm CostRegNet, the whole
single-stage path, and a CVP pyramid step."""
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import cases

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _fp32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("cin,cout,stride,transposed,dhw", [
    (8, 16, 1, False, (4, 6, 70)), (16, 32, 2, False, (4, 6, 10)), (32, 16, 2, True, (2, 3, 5)), (64, 32, 1, True, (2, 4, 9)),
    (8, 1, 1, False, (3, 5, 133)), (3, 5, 2, False, (6, 8, 12)), (16, 8, 2, True, (3, 5, 66))])
def test_conv3d_fn_gradients_vs_torch(cin, cout, stride, transposed, dhw):
    from mvs_b200.train import Conv3dFn
    g = torch.Generator(device=DEV).manual_seed(cin * 7 + cout)
    x = torch.randn(2, cin, *dhw, device=DEV, generator=g, requires_grad=True)
    wshape = (cin, cout, 3, 3, 3) if transposed else (cout, cin, 3, 3, 3)
    w = (torch.randn(wshape, device=DEV, generator=g) / (27 * cin) ** 0.5).requires_grad_()
    y = Conv3dFn.apply(x, w, stride, transposed)
    x2, w2 = x.detach().clone().requires_grad_(), w.detach().clone().requires_grad_()
    y2 = (F.conv_transpose3d(x2, w2, stride=stride, padding=1, output_padding=stride - 1) if transposed
          else F.conv3d(x2, w2, stride=stride, padding=1))
    gy = torch.randn(y2.shape, device=DEV, generator=g)
    y.backward(gy); y2.backward(gy)
    np.testing.assert_allclose(y.detach().cpu().numpy(), y2.detach().cpu().numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(x.grad.cpu().numpy(), x2.grad.cpu().numpy(), rtol=1e-4, atol=1e-5)
    scale = float(w2.grad.abs().max())
    np.testing.assert_allclose(w.grad.cpu().numpy(), w2.grad.cpu().numpy(), rtol=2e-4, atol=2e-5 * max(scale, 1.0))


def _ref_cbr(x, sd, conv, bn, stride=1, transposed=False):
    w = sd[conv + ".weight"]
    y = (F.conv_transpose3d(x, w, None, stride=stride, padding=1, output_padding=stride - 1) if transposed
         else F.conv3d(x, w, None, stride=stride, padding=1))
    y = F.batch_norm(y, None, None, sd[bn + ".weight"], sd[bn + ".bias"], True, 0.1, 1e-5)          # batch statistics
    return F.relu(y)


def _ref_costreg_cas_train(x, sd):
    """CasMVSNet CostRegNet.forward (module.py:429-438) in TRAINING mode, functional."""
    c0 = _ref_cbr(x, sd, "conv0.conv", "conv0.bn")
    c2 = _ref_cbr(_ref_cbr(c0, sd, "conv1.conv", "conv1.bn", 2), sd, "conv2.conv", "conv2.bn")
    c4 = _ref_cbr(_ref_cbr(c2, sd, "conv3.conv", "conv3.bn", 2), sd, "conv4.conv", "conv4.bn")
    y = _ref_cbr(_ref_cbr(c4, sd, "conv5.conv", "conv5.bn", 2), sd, "conv6.conv", "conv6.bn")
    y = c4 + _ref_cbr(y, sd, "conv7.conv", "conv7.bn", 2, True)
    y = c2 + _ref_cbr(y, sd, "conv9.conv", "conv9.bn", 2, True)
    y = c0 + _ref_cbr(y, sd, "conv11.conv", "conv11.bn", 2, True)
    return F.conv3d(y, sd["prob.weight"], None, padding=1)


def test_single_stage_training_gradients_vs_reference_ops():
    """feature maps -> fused builder -> CostRegNet (train-mode BN) -> softmax -> depth regression -> smooth-L1: every
    gradient (feature maps, conv weights, BN affine) against autograd through the reference's ATen op sequence."""
    from mvs_b200 import modules
    from mvs_b200.train import masked_smooth_l1
    from oracle import torch_port as TP
    v = cases.cas_case(n_views=3, B=2, C=16, H=16, W=24, D=8, seed=4, per_pixel=True)
    sd_np = cases.costreg_state("cas", cin=16, base=8, seed=12)
    feats = [torch.from_numpy(f).to(DEV) for f in v["feats"]]
    proj = torch.from_numpy(v["proj"]).to(DEV)
    depth = torch.from_numpy(v["depth"]).to(DEV)
    gt = depth[:, 3] + 2.0
    mask = torch.ones_like(gt)
    # ours
    reg = modules.CostRegNetCas(16, 8, mode="strict")
    reg.load_state_dict({k: torch.from_numpy(np.asarray(a)) for k, a in sd_np.items()}, strict=True)
    reg = reg.to(DEV).train()
    fo = [f.clone().requires_grad_() for f in feats]
    out = modules.DepthNet()(fo, proj, depth, depth.shape[1], reg)
    loss = masked_smooth_l1(out["depth"], gt, mask)
    loss.backward()
    # reference op sequence with autograd (training branch: out-of-place accumulation, mvsnet.py:158-161)
    sd = {k: torch.from_numpy(np.asarray(a)).to(DEV).requires_grad_(a.dtype == np.float32 and "running" not in k) for k, a in sd_np.items()}
    fr = [f.clone().requires_grad_() for f in feats]
    projs = [TP.cas_fuse_proj(p) for p in torch.unbind(proj, 1)]
    D = depth.shape[1]
    ref_vol = fr[0].unsqueeze(2).repeat(1, 1, D, 1, 1)
    vs, vq = ref_vol, ref_vol ** 2
    for f, p in zip(fr[1:], projs[1:]):
        w = TP.warp_volume(f, p, projs[0], depth)
        vs = vs + w
        vq = vq + w ** 2
    var = vq / 3 - (vs / 3) ** 2
    logits = _ref_costreg_cas_train(var, sd).squeeze(1)
    d_ref = torch.sum(F.softmax(logits, 1) * depth, 1)
    loss_ref = F.smooth_l1_loss(d_ref[mask > 0.5], gt[mask > 0.5], reduction="mean")
    loss_ref.backward()
    assert abs(float(loss) - float(loss_ref)) <= 1e-4 * abs(float(loss_ref))
    for a, b in zip(fo, fr):
        s = float(b.grad.abs().max())
        np.testing.assert_allclose(a.grad.cpu().numpy(), b.grad.cpu().numpy(), rtol=2e-3, atol=2e-4 * s)
    checked = 0
    for name, p in reg.named_parameters():
        gref = sd[name].grad
        s = float(gref.abs().max())
        np.testing.assert_allclose(p.grad.cpu().numpy(), gref.cpu().numpy(), rtol=5e-3, atol=5e-4 * max(s, 1e-6), err_msg=name)
        checked += 1
    assert checked == 31          # 11 conv weights + 10 x (BN weight, bias)
    # running statistics were updated like nn.BatchNorm3d(momentum=0.1) does in train mode
    assert int(reg.conv0.bn.num_batches_tracked) == 1
    assert not torch.equal(reg.conv0.bn.running_mean.cpu(), torch.from_numpy(sd_np["conv0.bn.running_mean"]))


def test_cvp_network_training_steps_reduce_the_loss():
    """CVP-MVSNet `network` mirror in train mode (BASELINE configs[3] shape family, tiny extents): three Adam steps through
    FeaturePyramid (ATen) + builder / CostRegNet / regression on the repo's kernels; finite gradients for every parameter,
    decreasing loss."""
    from mvs_b200 import pyramid
    from mvs_b200.train import masked_smooth_l1, train_step, GradBucket
    torch.manual_seed(0)
    args = types.SimpleNamespace(nsrc=2, nscale=2, mode="train")
    net = pyramid.network(args).to(DEV).train()
    B, H, W = 2, 32, 48
    ref_in, src_in, ref_ex, src_ex = [torch.from_numpy(a).to(DEV) for a in cases.synth.cvp_cameras(2, W, seed=6, batch=B)]
    imgs = torch.from_numpy(cases.synth.images_u8(3, H, W, seed=3, batch=B)).to(DEV).float() / 255.0
    dmin = torch.full((B,), 425.0, dtype=torch.float64, device=DEV)
    dmax = torch.full((B,), 935.0, dtype=torch.float64, device=DEV)
    gts = [torch.full((B, H >> i, W >> i), 650.0, device=DEV) for i in range(2)]
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    bucket = GradBucket(net.parameters())

    def fwd(m):
        out = m(imgs[:, 0], imgs[:, 1:], ref_in, src_in, ref_ex, src_ex, dmin, dmax)
        assert [tuple(d.shape) for d in out["depth_est_list"]] == [(B, H, W), (B, H // 2, W // 2)]
        return sum(masked_smooth_l1(d, g, torch.ones_like(g)) for d, g in zip(out["depth_est_list"], gts))

    losses = [float(train_step(net, opt, fwd, bucket)) for _ in range(4)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())


def test_cvp_network_training_loss_and_gradients_vs_reference_ops():
    """The CVP `network` mirror in train mode against the torch port of the reference's forward (oracle/torch_port.py
    cvp_network: FeaturePyramid, aliasing quirk, train-mode BatchNorm, fixed 6.8085 refinement interval): same loss, same
    gradients for the feature-pyramid weights (the longest chain: through builder backward, CostRegNet and regression)."""
    from mvs_b200 import pyramid
    from mvs_b200.train import masked_smooth_l1
    from oracle import torch_port as TP
    torch.manual_seed(1)
    net = pyramid.network(types.SimpleNamespace(nsrc=2, nscale=2, mode="train")).to(DEV).train()
    B, H, W = 1, 32, 48
    ref_in, src_in, ref_ex, src_ex = [torch.from_numpy(a).to(DEV) for a in cases.synth.cvp_cameras(2, W, seed=6, batch=B)]
    imgs = torch.from_numpy(cases.synth.images_u8(3, H, W, seed=3, batch=B)).to(DEV).float() / 255.0
    dmin = torch.full((B,), 425.0, dtype=torch.float64, device=DEV)
    dmax = torch.full((B,), 935.0, dtype=torch.float64, device=DEV)
    gts = [torch.full((B, H >> i, W >> i), 650.0, device=DEV) for i in range(2)]
    sd = {k: v.detach().clone().requires_grad_(v.dtype == torch.float32 and "running" not in k) for k, v in net.state_dict().items()}
    out = net(imgs[:, 0], imgs[:, 1:], ref_in, src_in, ref_ex, src_ex, dmin, dmax)
    loss = sum(masked_smooth_l1(d, g, torch.ones_like(g)) for d, g in zip(out["depth_est_list"], gts))
    loss.backward()
    ref = TP.cvp_network(imgs[:, 0], imgs[:, 1:], ref_in, src_in, ref_ex, src_ex, dmin, dmax, sd, 2, True)
    loss_ref = sum(F.smooth_l1_loss(d, g, reduction="mean") for d, g in zip(ref, gts))
    loss_ref.backward()
    for a, b in zip(out["depth_est_list"], ref):
        np.testing.assert_allclose(a.detach().cpu().numpy(), b.detach().cpu().numpy(), rtol=2e-5)
    assert abs(loss.item() - loss_ref.item()) <= 2e-5 * abs(loss_ref.item())
    n = 0
    for name, p in net.named_parameters():
        g = sd[name].grad
        if name.endswith("prob0.bias"):
            # softmax is invariant to a constant added to every logit: this gradient is exactly 0 in exact arithmetic,
            # both sides hold rounding noise
            assert float(p.grad.abs().max()) < 1e-3 and float(g.abs().max()) < 1e-3
        else:
            s = float(g.abs().max())
            np.testing.assert_allclose(p.grad.cpu().numpy(), g.cpu().numpy(), rtol=2e-2, atol=2e-3 * max(s, 1e-8), err_msg=name)
        n += 1
    assert n == 50


@pytest.mark.parametrize("shape,relu", [((2, 8, 5, 17, 33), True), ((1, 16, 3, 40, 50), False), ((3, 4, 2, 9, 700), True)])
def test_bn_relu_train_kernels_vs_aten(shape, relu):
    """csrc/bn_train.cu behind train.BnReluFn: output, running statistics and all three gradients against
    F.batch_norm(training=True) [+ F.relu] under torch autograd (fp32; sums in different orders)."""
    from mvs_b200 import train
    g = torch.Generator().manual_seed(7)
    x = (torch.randn(shape, generator=g) * 1.7 + 0.3).to(DEV).requires_grad_(True)
    C = shape[1]
    gamma = (torch.rand(C, generator=g) + 0.5).to(DEV).requires_grad_(True)
    beta = (torch.randn(C, generator=g) * 0.2).to(DEV).requires_grad_(True)
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    rm2, rv2 = rm.clone(), rv.clone()
    y = train.BnReluFn.apply(x, gamma, beta, rm, rv, 0.1, 1e-5, relu)
    x2, g2, b2 = (t.detach().clone().requires_grad_(True) for t in (x, gamma, beta))
    yr = F.batch_norm(x2, rm2, rv2, g2, b2, True, 0.1, 1e-5)
    if relu:
        yr = F.relu(yr)
    np.testing.assert_allclose(y.detach().cpu().numpy(), yr.detach().cpu().numpy(), rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(rm.cpu().numpy(), rm2.cpu().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(rv.cpu().numpy(), rv2.cpu().numpy(), rtol=1e-5, atol=1e-6)
    up = torch.randn(shape, generator=g).to(DEV)
    y.backward(up)
    yr.backward(up)
    for a, b, name in ((x.grad, x2.grad, "dx"), (gamma.grad, g2.grad, "dgamma"), (beta.grad, b2.grad, "dbeta")):
        scale = b.abs().max().item() + 1e-12
        assert (a - b).abs().max().item() <= 2e-5 * scale + 1e-6, name
