"""GPU parity tests: the CUDA path (through the C-ABI, via mvs_b200.ops / modules) against the CPU
oracle and the reference-generated golden fixtures.  Bit-exact for the warp / cost volume / tap
indices; stated tolerances for conv / softmax (fp32 summation order) and for the bf16 fast path."""
import numpy as np
import pytest
import torch

import cases
from oracle import oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def npy(t):
    return t.detach().cpu().numpy()


def assert_bitexact(a, b):
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    same = (a.view(np.uint32) == b.view(np.uint32)) | (a == b) | (np.isnan(a) & np.isnan(b))   # NaN payloads may differ
    assert same.all(), f"{(~same).sum()} of {same.size} differ"


def rt(proj):
    return np.ascontiguousarray(proj[..., :3, :3]), np.ascontiguousarray(proj[..., :3, 3])


# ---- a1: warp ------------------------------------------------------------------------------------
def test_homo_warping_plane_bitexact():
    from mvs_b200 import ops
    g = cases.golden("mvsnet_warp_plane"); c = cases.warp_plane_case()
    out = ops.homo_warping(cu(c["src_fea"]), cu(c["src_proj"]).cpu().to(DEV), cu(c["ref_proj"]), cu(c["depth"]))
    # projection product computed by torch on the GPU may differ in the last bit from the CPU's;
    # feed the CPU product for the bit-exact statement, and check the drop-in call to tolerance.
    rot, tr = rt(g["proj"])
    exact = ops._warp(cu(c["src_fea"]), cu(rot.reshape(-1, 9)), cu(tr), cu(c["depth"]), 0)
    assert_bitexact(npy(exact), g["out"])
    np.testing.assert_allclose(npy(out), g["out"], rtol=0, atol=2e-2)


def test_homo_warping_pixel_bitexact():
    from mvs_b200 import ops
    g = cases.golden("cas_warp_pixel"); c = cases.warp_pixel_case()
    rot, tr = rt(g["proj"])
    out = ops._warp(cu(c["src_fea"]), cu(rot.reshape(-1, 9)), cu(tr), cu(c["depth"]), 0)
    assert_bitexact(npy(out), g["out"])


def test_homo_warp_pl_bitexact():
    from mvs_b200 import ops, _lib as L
    g = cases.golden("pl_homo_warp"); c = cases.warp_plane_case(seed=9)
    rot, tr = rt(g["proj"])
    out = ops._warp(cu(c["src_fea"]), cu(rot.reshape(-1, 9)), cu(tr), cu(c["depth"]), L.ALIGN_CORNERS | L.PL_ORDER)
    assert_bitexact(npy(out), g["out"])
    # drop-in signature (ref_proj_inv supplied by the dataset)
    inv = torch.inverse(torch.from_numpy(c["ref_proj"]))
    out2 = ops.homo_warp(cu(c["src_fea"]), cu(c["src_proj"]), inv.to(DEV), cu(c["depth"]))
    np.testing.assert_allclose(npy(out2), g["out"], rtol=0, atol=2e-2)


def test_homo_warping_cvp_signature():
    from mvs_b200 import ops
    g = cases.golden("cvp_warp"); c = cases.cvp_case(per_pixel=False)
    out = ops.homo_warping_cvp(cu(c["feats"][1]), cu(c["ref_in"]), cu(c["src_in"][:, 0]), cu(c["ref_ex"]),
                               cu(c["src_ex"][:, 0]), cu(c["depth"]))
    np.testing.assert_allclose(npy(out), g["out"], rtol=0, atol=2e-2)


@pytest.mark.parametrize("shape", [(1, 48, 128, 160), (2, 3, 37, 50), (1, 1, 1, 1), (1, 2, 5, 33)])
@pytest.mark.parametrize("pixel", [False, True])
@pytest.mark.parametrize("flags", [0, 3])
def test_tap_indices_bitexact(shape, pixel, flags):
    """Integer tap indices, in-bounds masks and fp32 sample positions identical to the oracle
    (cfg1 extent 48x128x160 and ragged / degenerate extents)."""
    from mvs_b200 import ops
    B, D, H, W = shape
    proj = cases.synth.proj_matrices(2, max(W, 2), seed=3, batch=B)
    p = torch.from_numpy(proj)
    prod = (p[:, 1] @ torch.inverse(p[:, 0])).numpy()
    rot, tr = rt(prod)
    depth = cases.synth.depth_per_pixel(D, H, W, 10.6, B) if pixel else cases.synth.depth_planes(D, B)
    x0, y0, mask, ixy = ops.warp_taps(cu(rot.reshape(-1, 9)), cu(tr), cu(depth), H, W, flags)
    ox0, oy0, omask, oixy = O.warp_taps(rot, tr, depth, H, W, flags)
    assert np.array_equal(npy(x0), ox0) and np.array_equal(npy(y0), oy0)
    assert np.array_equal(npy(mask), omask)
    assert_bitexact(npy(ixy), oixy)


# ---- a1+a2: fused builder --------------------------------------------------------------------------
def _volume(feats, proj, depth, flags=0):
    from mvs_b200 import ops
    rot, tr = rt(proj)
    nsrc = rot.shape[1]
    rots = [cu(rot[:, i].reshape(-1, 9)) for i in range(nsrc)]
    trs = [cu(tr[:, i]) for i in range(nsrc)]
    return ops.cost_volume(cu(feats[0]), [cu(f) for f in feats[1:]], rots, trs, cu(depth), flags)


def test_cost_volume_mvsnet_bitexact():
    g = cases.golden("mvsnet_forward"); v = cases.volume_case(n_views=4, C=32, H=16, W=24, D=8, seed=3)
    assert_bitexact(npy(_volume(v["feats"], g["proj"], v["depth"])), g["var"])


@pytest.mark.parametrize("tag,cin,pp,seed", [("s2", 16, True, 4), ("s1", 32, False, 8)])
def test_cost_volume_cas_bitexact(tag, cin, pp, seed):
    g = cases.golden(f"cas_depthnet_{tag}"); c = cases.cas_case(C=cin, per_pixel=pp, seed=seed)
    assert_bitexact(npy(_volume(c["feats"], g["proj"], c["depth"])), g["var"])


def test_cost_volume_cvp_quirk_bitexact():
    from mvs_b200 import _lib as L
    from test_oracle_golden import cvp_products
    g = cases.golden("cvp_proj_cost"); c = cases.cvp_case(per_pixel=True)
    assert_bitexact(npy(_volume(c["feats"], cvp_products(c), c["depth"], L.REF_SUM_SQUARED)), g["out"])


@pytest.mark.parametrize("nsrc", [1, 2, 6, 8])
def test_cost_volume_view_counts_vs_oracle(nsrc):
    v = cases.volume_case(n_views=nsrc + 1, C=5, H=13, W=35, D=3, seed=20 + nsrc, per_pixel=True)
    p = torch.from_numpy(v["proj"])
    prod = torch.stack([p[:, i] @ torch.inverse(p[:, 0]) for i in range(1, nsrc + 1)], 1).numpy()
    rot, tr = rt(prod)
    assert_bitexact(npy(_volume(v["feats"], prod, v["depth"])), O.cost_volume(v["feats"][0], v["feats"][1:], rot, tr, v["depth"]))


def test_cost_volume_edge_cases():
    from mvs_b200 import ops
    # empty batch -> empty volume, no launch error
    z = torch.zeros(0, 4, 8, 8, device=DEV)
    out = ops.cost_volume(z, [z], [torch.zeros(0, 9, device=DEV)], [torch.zeros(0, 3, device=DEV)], torch.zeros(0, 3, device=DEV))
    assert out.shape == (0, 4, 3, 8, 8)
    # camera looking away: every tap out of bounds -> variance of (ref, 0, 0)
    fea = cases.synth.features(3, 4, 8, 12, seed=1)
    rot = np.tile(np.eye(3, dtype=np.float32).reshape(1, 9), (1, 1)); tr = np.array([[1e6, 1e6, 1.0]], np.float32)
    depth = cases.synth.depth_planes(2)
    out = ops.cost_volume(cu(fea[0]), [cu(fea[1]), cu(fea[2])], [cu(rot)] * 2, [cu(tr)] * 2, cu(depth))
    ref = fea[0][:, :, None]
    expect = (ref * ref) / np.float32(3) - (ref / np.float32(3)) ** 2
    np.testing.assert_allclose(npy(out), np.broadcast_to(expect, out.shape), rtol=1e-6, atol=1e-7)
    # non-finite sample positions propagate NaN like the oracle / ATen CPU kernel
    bad = ops._warp(cu(fea[1]), torch.zeros(1, 9, device=DEV), torch.zeros(1, 3, device=DEV), cu(depth), 0)
    assert torch.isnan(bad).all()
    # too many views -> ValueError (the fused launch takes <= MVS_MAX_SRC)
    with pytest.raises(ValueError):
        ops.cost_volume(cu(fea[0]), [cu(fea[1])] * 9, [cu(rot)] * 9, [cu(tr)] * 9, cu(depth))


# ---- a3: convolutions ------------------------------------------------------------------------------
@pytest.mark.parametrize("cin,cout,stride,transposed", [
    (8, 8, 1, False), (32, 8, 1, False), (8, 16, 2, False), (16, 32, 2, False), (64, 64, 1, False),
    (64, 32, 2, True), (16, 8, 2, True), (64, 32, 1, True), (8, 1, 1, False), (5, 3, 1, False)])
def test_conv3d_layer_vs_oracle(cin, cout, stride, transposed):
    from mvs_b200 import ops
    rng = np.random.RandomState(cin * 100 + cout)
    D, H, W = (4, 6, 10) if not (stride == 2 and not transposed) else (5, 7, 11)   # odd extents for stride 2
    x = rng.standard_normal((2, cin, D, H, W)).astype(np.float32)
    wshape = (cin, cout, 3, 3, 3) if transposed else (cout, cin, 3, 3, 3)
    w = (rng.standard_normal(wshape) / np.sqrt(27 * cin)).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32); shift = rng.standard_normal(cout).astype(np.float32)
    ref = O.conv3d(x, w, None, stride, transposed)
    ref = ref * scale.reshape(1, -1, 1, 1, 1) + shift.reshape(1, -1, 1, 1, 1)
    ref = np.maximum(ref, 0)
    skip = rng.standard_normal(ref.shape).astype(np.float32)
    ref = skip + ref
    out = ops.conv3d(cu(x), cu(w), cu(scale), cu(shift), cu(skip), stride, transposed, relu=True)
    np.testing.assert_allclose(npy(out), ref, rtol=1e-5, atol=1e-5)


def _load(module, sd):
    module.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    return module.to(DEV).eval()


def test_costreg_mvsnet_golden():
    from mvs_b200 import modules
    net = _load(modules.CostRegNet(), cases.costreg_state("mvsnet", seed=11))
    x = np.random.RandomState(21).standard_normal((1, 32, 8, 16, 24)).astype(np.float32)
    with torch.no_grad():
        out = net(cu(x))
    np.testing.assert_allclose(npy(out), cases.golden("mvsnet_costreg")["out"], rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("cin", [32, 16, 8])
def test_costreg_cas_golden(cin):
    from mvs_b200 import modules
    net = _load(modules.CostRegNet(cin, 8), cases.costreg_state("cas", cin=cin, seed=13))
    x = np.random.RandomState(22 + cin).standard_normal((1, cin, 8, 16, 24)).astype(np.float32)
    with torch.no_grad():
        out = net(cu(x))
    np.testing.assert_allclose(npy(out), cases.golden(f"cas_costreg_c{cin}")["out"], rtol=1e-4, atol=2e-5)


def test_costreg_cvp_golden():
    from mvs_b200 import modules
    net = _load(modules.CostRegNetCVP(), cases.costreg_state("cvp", seed=15))
    x = np.random.RandomState(23).standard_normal((1, 16, 8, 16, 24)).astype(np.float32)
    with torch.no_grad():
        out = net(cu(x))
    np.testing.assert_allclose(npy(out), cases.golden("cvp_costreg")["out"], rtol=1e-4, atol=2e-5)


# ---- a4-a6: softmax / regression / confidence --------------------------------------------------------
@pytest.mark.parametrize("pixel", [False, True])
def test_softargmin_conf(pixel):
    from mvs_b200 import ops
    c = cases.logits_case(per_pixel=pixel)
    ref = O.softargmin_conf(c["logits"], c["depth"], clamp_index=True, want_prob=True)
    depth, conf, prob, index = ops.softargmin_conf(cu(c["logits"]), cu(c["depth"]), clamp_index=True, want_prob=True,
                                                   want_index=True)
    np.testing.assert_allclose(npy(depth), ref["depth"], rtol=1e-6, atol=0)      # fp32 relative L-inf
    np.testing.assert_allclose(npy(prob), ref["prob"], rtol=1e-5, atol=1e-7)
    stable = np.abs(ref["expect_idx"] - np.round(ref["expect_idx"])) > 1e-3
    assert (npy(index) == ref["index"])[stable].all()
    np.testing.assert_allclose(npy(conf)[stable], ref["conf"][stable], rtol=1e-5, atol=1e-6)
    if not pixel:
        g = cases.golden("mvsnet_regress")
        np.testing.assert_allclose(npy(depth), g["depth"], rtol=1e-6, atol=0)
        # drop-in depth_regression(p, depth_values) on a probability volume
        d2 = ops.depth_regression(cu(g["prob"]), cu(c["depth"]))
        np.testing.assert_allclose(npy(d2), g["depth"], rtol=1e-6, atol=0)


def test_one_hot_known_answer():
    from mvs_b200 import ops
    D = 6
    depth = cases.synth.depth_planes(D)
    logits = np.full((1, D, 3, 5), -1e4, np.float32); logits[:, 4] = 0
    d, c, _, idx = ops.softargmin_conf(cu(logits), cu(depth), want_index=True)
    assert np.allclose(npy(d), depth[0, 4]) and np.allclose(npy(c), 1.0) and (npy(idx) == 4).all()


# ---- whole-path goldens ------------------------------------------------------------------------------
def test_mvsnet_hot_path_golden():
    """MVSNet.forward from features to depth (depth within 1e-4 relative L-inf, north_star)."""
    from mvs_b200 import modules
    g = cases.golden("mvsnet_forward"); v = cases.volume_case(n_views=4, C=32, H=16, W=24, D=8, seed=3)
    reg = _load(modules.CostRegNet(), cases.costreg_state("mvsnet", seed=11))
    with torch.no_grad():
        out = modules.mvsnet_hot_path([cu(f) for f in v["feats"]], cu(v["proj"]), cu(v["depth"]), reg)
    np.testing.assert_allclose(npy(out["depth"]), g["depth"], rtol=1e-4, atol=0)
    np.testing.assert_allclose(npy(out["photometric_confidence"]), g["conf"], rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("tag,cin,pp,seed", [("s2", 16, True, 4), ("s1", 32, False, 8)])
def test_cas_depthnet_golden(tag, cin, pp, seed):
    from mvs_b200 import modules
    g = cases.golden(f"cas_depthnet_{tag}"); c = cases.cas_case(C=cin, per_pixel=pp, seed=seed)
    reg = _load(modules.CostRegNet(cin, 8), cases.costreg_state("cas", cin=cin, seed=12))
    depth = cu(c["depth"])
    with torch.no_grad():
        out = modules.DepthNet()([cu(f) for f in c["feats"]], cu(c["proj"]), depth, depth.shape[1], reg)
    np.testing.assert_allclose(npy(out["depth"]), g["depth"], rtol=1e-4, atol=0)
    np.testing.assert_allclose(npy(out["photometric_confidence"]), g["conf"], rtol=1e-3, atol=1e-4)


def test_cas_cascade_golden():
    """Full 3-stage CascadeMVSNet hot path vs the reference's forward (FeatureNet stubbed)."""
    from mvs_b200 import modules, cascade
    g = cases.golden("cas_cascade"); k = cases.cascade_case()
    regs = [_load(modules.CostRegNet(cin, 8), cases.costreg_state("cas", cin=cin, seed=14 + i))
            for i, cin in enumerate((32, 16, 8))]
    n_views = k["feats"]["stage1"].shape[0]
    feats = [{s: cu(k["feats"][s][v]) for s in k["feats"]} for v in range(n_views)]
    with torch.no_grad():
        out = cascade.cascade_hot_path(feats, {s: cu(p) for s, p in k["projs"].items()}, cu(k["depth_values"]), regs,
                                       ndepths=k["ndepths"], img_hw=(k["H"], k["W"]))
    for s in ("stage1", "stage2", "stage3"):
        np.testing.assert_allclose(npy(out[s]["depth"]), g[s + "_depth"], rtol=1e-4, atol=0)
        np.testing.assert_allclose(npy(out[s]["photometric_confidence"]), g[s + "_conf"], rtol=2e-3, atol=2e-4)
    assert out["depth"] is out["stage3"]["depth"]


def test_cvp_proj_cost_golden():
    import types
    from mvs_b200 import modules
    g = cases.golden("cvp_proj_cost"); c = cases.cvp_case(per_pixel=True)
    settings = types.SimpleNamespace(nsrc=2, mode="test")
    srcs = [[cu(c["feats"][1 + i])] for i in range(2)]
    vol = modules.proj_cost(settings, cu(c["feats"][0]), srcs, 0, cu(c["ref_in"]), cu(c["src_in"]), cu(c["ref_ex"]),
                            cu(c["src_ex"]), cu(c["depth"]))
    np.testing.assert_allclose(npy(vol), g["out"], rtol=1e-4, atol=1e-4)


def test_range_samples_bitexact():
    from mvs_b200 import ops
    cur = np.random.RandomState(31).uniform(500, 800, (2, 12, 20)).astype(np.float32)
    assert_bitexact(npy(ops.depth_range_samples(cu(cur), 8, 5.3)), cases.golden("cas_range_samples")["out"])


# ---- fast path: C8 / bf16 ---------------------------------------------------------------------------
@pytest.mark.parametrize("C", [8, 16, 32, 5])
def test_c8_pack_roundtrip(C):
    from mvs_b200 import ops
    x = torch.randn(2, C, 3, 7, 9, device=DEV)
    p = ops.pack_c8(x)
    assert p.shape == (2, (C + 7) // 8, 3, 7, 9, 8) and p.dtype == torch.bfloat16
    back = ops.unpack_c8(p, C)
    assert torch.equal(back, x.bfloat16().float())
    assert torch.equal(ops.unpack_c8(ops.pack_c8(x.bfloat16()), C, torch.bfloat16), x.bfloat16())


@pytest.mark.parametrize("C,pixel,nsrc", [(32, False, 4), (16, True, 4), (8, True, 2), (8, False, 6)])
def test_cost_volume_c8_vs_oracle(C, pixel, nsrc):
    """bf16 storage, fp32 math: against the oracle fed the SAME bf16-rounded features the error is
    one bf16 rounding of the result (rel 2^-8) plus fp32 reassociation."""
    from mvs_b200 import ops
    v = cases.volume_case(n_views=nsrc + 1, C=C, H=21, W=40, D=6, seed=30 + C, per_pixel=pixel)
    feats = torch.from_numpy(v["feats"]).bfloat16().float().numpy()
    p = torch.from_numpy(v["proj"])
    prod = torch.stack([p[:, i] @ torch.inverse(p[:, 0]) for i in range(1, nsrc + 1)], 1).numpy()
    rot, tr = rt(prod)
    ref = O.cost_volume(feats[0], feats[1:], rot, tr, v["depth"])
    rots = [cu(rot[:, i].reshape(-1, 9)) for i in range(nsrc)]
    trs = [cu(tr[:, i]) for i in range(nsrc)]
    packed = [ops.pack_c8(cu(f)) for f in feats]
    vol = ops.cost_volume_c8(packed[0], packed[1:], rots, trs, cu(v["depth"]))
    out = npy(ops.unpack_c8(vol, C))
    np.testing.assert_allclose(out, ref, rtol=2 ** -7, atol=2e-3)


# ---- backward (training hook; torch fp32 autograd as the reference for this floating-point kernel) --
def test_cost_volume_backward_vs_torch_autograd():
    import torch.nn.functional as F
    from mvs_b200 import ops
    v = cases.volume_case(n_views=3, C=4, H=10, W=14, D=3, seed=40, per_pixel=True)
    p = torch.from_numpy(v["proj"])
    prod = torch.stack([p[:, i] @ torch.inverse(p[:, 0]) for i in range(1, 3)], 1).numpy()
    rot, tr = rt(prod)
    B, C, H, W = v["feats"][0].shape
    D = v["depth"].shape[1]
    x0, y0, mask, ixy = O.warp_taps(rot[:, 0], tr[:, 0], v["depth"], H, W)
    feats = [cu(f).requires_grad_(True) for f in v["feats"]]
    rots = [cu(rot[:, i].reshape(-1, 9)) for i in range(2)]
    trs = [cu(tr[:, i]) for i in range(2)]
    vol = ops.cost_volume(feats[0], feats[1:], rots, trs, cu(v["depth"]))
    gout = torch.randn_like(vol)
    vol.backward(gout)
    got = [f.grad.clone() for f in feats]
    # torch reference: grid_sample with the oracle's sample positions (align_corners=False unnormalised -> normalised)
    tf = [cu(f).requires_grad_(True) for f in v["feats"]]
    vols = [tf[0].unsqueeze(2).expand(B, C, D, H, W)]
    for i in range(2):
        _, _, _, ixy_i = O.warp_taps(rot[:, i], tr[:, i], v["depth"], H, W)
        g = cu(ixy_i)
        gx = (g[..., 0] + 0.5) * 2 / W - 1
        gy = (g[..., 1] + 0.5) * 2 / H - 1
        grid = torch.stack((gx, gy), -1).view(B, D * H, W, 2)
        vols.append(F.grid_sample(tf[i + 1], grid, mode="bilinear", padding_mode="zeros", align_corners=False).view(B, C, D, H, W))
    s = sum(vols) / 3
    var = sum(x * x for x in vols) / 3 - s * s
    np.testing.assert_allclose(npy(vol), npy(var), rtol=1e-4, atol=1e-4)
    var.backward(gout)
    for a, b in zip(got, tf):
        np.testing.assert_allclose(npy(a), npy(b.grad), rtol=1e-3, atol=1e-3)


# ---- size-independent properties at BASELINE.json's full extents (cfg3 stage shapes) ----------------
@pytest.mark.parametrize("C,D,H,W", [(32, 48, 296, 400), (8, 8, 1184, 1600)])
def test_full_size_properties(C, D, H, W):
    from mvs_b200 import ops, _lib as L
    torch.manual_seed(0)
    fea = torch.randn(1, C, H, W, device=DEV)
    eye = torch.eye(3, device=DEV).reshape(1, 9); zero = torch.zeros(1, 3, device=DEV)
    depth = cu(cases.synth.depth_planes(D))
    # identical views + identity pose + align_corners => variance == 0 up to rounding
    var = ops.cost_volume(fea, [fea, fea], [eye, eye], [zero, zero], depth, L.ALIGN_CORNERS)
    assert var.abs().max().item() < 1e-4
    del var
    # scaling the features by k scales the variance by k^2 (k a power of two: exact)
    proj = cases.synth.proj_matrices(3, W, seed=2)
    p = torch.from_numpy(proj)
    prod = torch.stack([p[:, i] @ torch.inverse(p[:, 0]) for i in range(1, 3)], 1)
    rots = [prod[:, i, :3, :3].reshape(1, 9).contiguous().to(DEV) for i in range(2)]
    trs = [prod[:, i, :3, 3].contiguous().to(DEV) for i in range(2)]
    srcs = [torch.randn(1, C, H, W, device=DEV) for _ in range(2)]
    v1 = ops.cost_volume(fea, srcs, rots, trs, depth)
    v2 = ops.cost_volume(fea * 2, [s * 2 for s in srcs], rots, trs, depth)
    assert torch.equal(v2, v1 * 4)
    del v2
    # fast path agrees with the strict path to bf16 rounding on bf16-representable inputs
    fb = [t.bfloat16().float() for t in [fea] + srcs]
    strict = ops.cost_volume(fb[0], fb[1:], rots, trs, depth)
    fast = ops.unpack_c8(ops.cost_volume_c8(ops.pack_c8(fb[0]), [ops.pack_c8(t) for t in fb[1:]], rots, trs, depth), C)
    err = (fast - strict).abs()
    assert (err <= strict.abs() * 2 ** -7 + 2e-3).all()


# ---- fast-path tap arithmetic (C8 builder): indices must still be the reference's -------------------
@pytest.mark.parametrize("shape", [(1, 48, 128, 160), (1, 8, 296, 400), (2, 3, 37, 50), (1, 2, 5, 33)])
@pytest.mark.parametrize("pixel", [False, True])
@pytest.mark.parametrize("flags", [0, 3])
def test_fast_path_tap_indices_bitexact(shape, pixel, flags):
    """The C8 builder replaces the IEEE-division calls by the reciprocal/Newton/residual sequence;
    its sample positions, integer taps and masks must stay bit-identical to the oracle."""
    from mvs_b200 import ops, _lib as L
    B, D, H, W = shape
    proj = cases.synth.proj_matrices(2, W, seed=5, batch=B)
    p = torch.from_numpy(proj)
    prod = (p[:, 1] @ torch.inverse(p[:, 0])).numpy()
    rot, tr = rt(prod)
    depth = cases.synth.depth_per_pixel(D, H, W, 10.6, B) if pixel else cases.synth.depth_planes(D, B)
    x0, y0, mask, ixy = ops.warp_taps(cu(rot.reshape(-1, 9)), cu(tr), cu(depth), H, W, flags | L.FAST_COORDS)
    ox0, oy0, omask, oixy = O.warp_taps(rot, tr, depth, H, W, flags)
    assert_bitexact(npy(ixy), oixy)
    assert np.array_equal(npy(x0), ox0) and np.array_equal(npy(y0), oy0)
    assert np.array_equal(npy(mask), omask)


@pytest.mark.parametrize("C,pixel,nsrc", [(32, False, 4), (8, True, 2)])
def test_cost_volume_c8_bf16_blend_vs_oracle(C, pixel, nsrc):
    """MVS_BLEND_BF16: the bilinear blend itself runs in packed bf16 (weights and partial sums rounded
    to 8 mantissa bits), sums over views in fp32.  Stated tolerance vs the fp32 oracle on the same
    bf16 features: 2^-5 relative + 0.03 absolute (features ~N(0,1))."""
    from mvs_b200 import ops, _lib as L
    v = cases.volume_case(n_views=nsrc + 1, C=C, H=21, W=40, D=6, seed=30 + C, per_pixel=pixel)
    feats = torch.from_numpy(v["feats"]).bfloat16().float().numpy()
    p = torch.from_numpy(v["proj"])
    prod = torch.stack([p[:, i] @ torch.inverse(p[:, 0]) for i in range(1, nsrc + 1)], 1).numpy()
    rot, tr = rt(prod)
    ref = O.cost_volume(feats[0], feats[1:], rot, tr, v["depth"])
    rots = [cu(rot[:, i].reshape(-1, 9)) for i in range(nsrc)]
    trs = [cu(tr[:, i]) for i in range(nsrc)]
    packed = [ops.pack_c8(cu(f)) for f in feats]
    vol = ops.cost_volume_c8(packed[0], packed[1:], rots, trs, cu(v["depth"]), L.BLEND_BF16)
    out = npy(ops.unpack_c8(vol, C))
    np.testing.assert_allclose(out, ref, rtol=2 ** -5, atol=3e-2)
    print("bf16-blend max abs err", np.abs(out - ref).max(), "mean abs err", np.abs(out - ref).mean())


@pytest.mark.parametrize("C,pixel,nsrc", [(32, False, 4), (16, True, 4), (8, True, 2), (8, False, 6)])
def test_cost_volume_c8h_fp16_features_vs_oracle(C, pixel, nsrc):
    """MVS_FEAT_F16 (the fast path's default feature format): fp16 C8H features, bilinear blend in packed fp16
    (11-bit significands), sums over views / variance in fp32, bf16 volume out.  Against the oracle fed the SAME
    fp16-rounded features the tolerance is the bf16 rounding of the stored result plus the four fp16 roundings of
    the blend: 2^-7 relative + 4e-3 absolute (features ~N(0,1)) -- the same bound as the fp32-blend bf16 path."""
    from mvs_b200 import ops
    v = cases.volume_case(n_views=nsrc + 1, C=C, H=21, W=40, D=6, seed=30 + C, per_pixel=pixel)
    feats = torch.from_numpy(v["feats"]).half().float().numpy()
    p = torch.from_numpy(v["proj"])
    prod = torch.stack([p[:, i] @ torch.inverse(p[:, 0]) for i in range(1, nsrc + 1)], 1).numpy()
    rot, tr = rt(prod)
    ref = O.cost_volume(feats[0], feats[1:], rot, tr, v["depth"])
    rots = [cu(rot[:, i].reshape(-1, 9)) for i in range(nsrc)]
    trs = [cu(tr[:, i]) for i in range(nsrc)]
    packed = [ops.pack_c8(cu(f), torch.float16) for f in feats]
    assert packed[0].dtype == torch.float16
    # the C8H pack is exact on fp16-representable inputs: [B,CB,H,W,8] -> NCHW
    back = packed[0].permute(0, 1, 4, 2, 3).reshape(feats[0].shape[0], -1, *feats[0].shape[2:])[:, :C].float()
    assert np.array_equal(npy(back), feats[0])
    vol = ops.cost_volume_c8(packed[0], packed[1:], rots, trs, cu(v["depth"]))
    out = npy(ops.unpack_c8(vol, C))
    np.testing.assert_allclose(out, ref, rtol=2 ** -7, atol=4e-3)
    print("fp16-feature path max abs err", np.abs(out - ref).max(), "mean abs err", np.abs(out - ref).mean())


def test_pack_c8h_saturates():
    from mvs_b200 import ops
    x = torch.tensor([1e6, -1e6, 3.0, float("nan")], device=DEV).reshape(1, 4, 1, 1)
    p = ops.pack_c8(x, torch.float16).flatten()[:4].float()
    assert p[0] == 65504 and p[1] == -65504 and p[2] == 3 and torch.isnan(p[3])


def test_cvp_pyramid_golden():
    """CVP-MVSNet network.forward (nscale=2, test mode) from feature pyramids: coarse sweep + one refine
    level incl. calDepthHypo's statistical interval, vs the reference's outputs."""
    from mvs_b200 import modules, pyramid
    g = cases.golden("cvp_network")
    H, W = 32, 48
    fine = cases.synth.features(3, 16, H, W, 41, 1)
    coarse = cases.synth.features(3, 16, H // 2, W // 2, 42, 1)
    ref_in, src_in, ref_ex, src_ex = cases.synth.cvp_cameras(2, W, 43, 1)
    reg = _load(modules.CostRegNetCVP(), cases.costreg_state("cvp", seed=15))
    dmin = torch.tensor([425.0], dtype=torch.float64); dmax = torch.tensor([935.0], dtype=torch.float64)
    with torch.no_grad():
        out = pyramid.cvp_hot_path([cu(fine[0]), cu(coarse[0])], [[cu(fine[1]), cu(coarse[1])], [cu(fine[2]), cu(coarse[2])]],
                                   cu(ref_in), cu(src_in), cu(ref_ex), cu(src_ex), dmin.to(DEV), dmax.to(DEV), reg, (H, W), mode="test")
    np.testing.assert_allclose(npy(out["depth_est_list"][1]), g["depth1"], rtol=1e-4, atol=0)     # coarse
    np.testing.assert_allclose(npy(out["depth_est_list"][0]), g["depth0"], rtol=1e-4, atol=0)     # refined
    np.testing.assert_allclose(npy(out["prob_confidence"]), g["conf"], rtol=2e-3, atol=2e-4)


@pytest.mark.parametrize("scale,nd", [(2, 32), (1, 8), (4, 6)])
def test_cas_hypotheses_fused_vs_composition(scale, nd):
    """The fused inter-stage kernel vs the reference's three-op composition (bilinear up-sampling,
    get_depth_range_samples, trilinear resampling: cas_mvsnet.py:129-151) run with torch on the GPU."""
    import torch.nn.functional as F
    from mvs_b200 import ops
    H, W = 64, 96
    prev = cu(np.random.RandomState(3).uniform(450, 900, (2, H // (2 * scale), W // (2 * scale))).astype(np.float32))
    cur = F.interpolate(prev.unsqueeze(1), [H, W], mode="bilinear", align_corners=False).squeeze(1)
    samples = ops.depth_range_samples(cur, nd, 2 * 2.65)
    ref = F.interpolate(samples.unsqueeze(1), [nd, H // scale, W // scale], mode="trilinear", align_corners=False).squeeze(1)
    out = ops.cas_hypotheses(prev, (H, W), (H // scale, W // scale), nd, 2 * 2.65)
    np.testing.assert_allclose(npy(out), npy(ref), rtol=3e-7, atol=0)


# ---- TMA-staged builder (warp_tma.cu) == L1-gather builder (warp_c8.cu), bit for bit ------------------------------
def _c8h_case(n_views, B, C, H, W, D, seed, pixel, interval=10.6, baseline_gain=1.0):
    v = cases.volume_case(n_views=n_views, B=B, C=C, H=H, W=W, D=D, seed=seed, per_pixel=pixel)
    if pixel and interval != 10.6:
        v["depth"] = cases.synth.depth_per_pixel(D, H, W, interval, B)
    p = torch.from_numpy(v["proj"]).double()
    prod = torch.stack([p[:, i] @ torch.inverse(p[:, 0]) for i in range(1, n_views)], 1)
    prod[:, :, :3, 3] *= baseline_gain                       # wider baselines => larger per-depth motion => box overflow
    rot, tr = rt(prod.float().numpy())
    return v, rot, tr


def _both_builders(feats, rot, tr, depth, flags=0):
    from mvs_b200 import ops, _lib as L
    nsrc = len(feats) - 1
    rots = [cu(rot[:, i].reshape(-1, 9)) for i in range(nsrc)]
    trs = [cu(tr[:, i]) for i in range(nsrc)]
    packed = [ops.pack_c8(cu(f), torch.float16) for f in feats]
    d = cu(depth)
    assert nsrc <= 6, "the TMA builder takes up to 6 source views (more: the library falls back to the gather kernel)"
    tma = ops.cost_volume_c8(packed[0], packed[1:], rots, trs, d, flags | L.WARP_TMA)
    gather = ops.cost_volume_c8(packed[0], packed[1:], rots, trs, d, flags | L.WARP_NO_TMA)
    return tma, gather


@pytest.mark.parametrize("C,pixel,nsrc,B,H,W,D", [
    (32, False, 4, 1, 21, 40, 6), (16, True, 4, 2, 37, 50, 12), (8, True, 2, 1, 64, 96, 8), (8, False, 6, 1, 19, 33, 5),
    (8, True, 1, 1, 8, 32, 3), (16, False, 6, 1, 24, 70, 9), (8, True, 4, 1, 2, 2, 2), (24, True, 3, 2, 130, 161, 17),
    (8, False, 5, 1, 40, 64, 4), (16, True, 5, 2, 33, 47, 6)])
def test_tma_builder_equals_gather_builder(C, pixel, nsrc, B, H, W, D):
    """Same taps, same weights, same op order: the staged box only decides WHERE a tap is read from."""
    v, rot, tr = _c8h_case(nsrc + 1, B, C, H, W, D, 70 + C + nsrc, pixel)
    tma, gather = _both_builders(v["feats"], rot, tr, v["depth"])
    assert torch.equal(tma.view(torch.int16), gather.view(torch.int16))
    assert not torch.isnan(tma.float()).any()


@pytest.mark.parametrize("gain,interval,flags", [(6.0, 10.6, 0), (1.0, 400.0, 0), (25.0, 60.0, 0), (1.0, 10.6, 3), (1.0, 10.6, 4)])
def test_tma_builder_box_overflow_and_flags(gain, interval, flags):
    """Footprints far larger than the staged box (wide baselines, huge hypothesis spacing) take the per-lane global
    fallback; align_corners + PL op order (3) and the CVP quirk (4) go through the same kernel.  Still bit-identical."""
    v, rot, tr = _c8h_case(5, 1, 16, 72, 100, 16, 91, True, interval=interval, baseline_gain=gain)
    tma, gather = _both_builders(v["feats"], rot, tr, v["depth"], flags)
    assert torch.equal(tma.view(torch.int16), gather.view(torch.int16))


def test_tma_builder_degenerate_inputs():
    """Non-finite / negative / zero hypotheses and a singular pose: NaN voxels where the reference's CPU path has NaN,
    identical bits to the gather kernel everywhere (NaN payloads included: both write the canonical 0x7fc0)."""
    v, rot, tr = _c8h_case(3, 1, 8, 24, 40, 8, 93, True)
    d = v["depth"].copy()
    d[0, 1, 3:9, 5:17] = np.nan
    d[0, 2, 10:14, :] = 0.0
    d[0, 3, :, 20:30] = -300.0
    d[0, 4, 0, 0] = np.inf
    rot[0, 1] = 0.0
    tma, gather = _both_builders(v["feats"], rot, tr, d)
    assert torch.equal(tma.view(torch.int16), gather.view(torch.int16))
    assert torch.isnan(tma.float()).any()


@pytest.mark.parametrize("stage", [0, 1, 2])
def test_tma_builder_full_size_cfg3(stage):
    """BASELINE cfg3 stage extents (N=5): TMA-staged == gather kernel bit for bit on the full volume, and a row band of
    the result against the CPU oracle on the fp16-rounded features."""
    C, D, H, W = cases.synth.CONFIGS["cfg3"]["stages"][stage]
    v, rot, tr = _c8h_case(5, 1, C, H, W, D, 5 + stage, stage > 0, interval=(2.65 * (4, 2, 1)[stage]))
    tma, gather = _both_builders(v["feats"], rot, tr, v["depth"])
    assert torch.equal(tma.view(torch.int16), gather.view(torch.int16))
    del gather
    from mvs_b200 import ops
    # CPU oracle on a band of reference rows (fp16-rounded features in, same tolerance as the small-size oracle tests)
    y0, y1 = H // 2 - 3, H // 2 + 3
    feats = torch.from_numpy(v["feats"]).half().float().numpy()
    out = npy(ops.unpack_c8(tma, C))[:, :, :, y0:y1]
    del tma
    ref_band = O.cost_volume(feats[0], feats[1:], rot, tr, v["depth"], rows=(y0, y1))
    np.testing.assert_allclose(out, ref_band, rtol=2 ** -7, atol=4e-3)


def test_cas_poses_kernel_vs_float64_algebra():
    """mvs_cas_poses (one launch, fp64 inside) against the reference's K @ E, inverse, product sequence evaluated in float64:
    equal to fp32 rounding; against the fp32 torch sequence the strict path keeps: within a few fp32 ulps."""
    from mvs_b200 import modules
    pm = torch.from_numpy(np.stack([cases.synth.cas_proj_matrices(6, w, seed=3, batch=2) for w in (100, 200, 400)])).to(DEV)   # [S,B,N,2,4,4]
    rot, trans = modules.cas_relative_poses(pm, fused_kernel=True)
    rot32, trans32 = modules.cas_relative_poses(pm, fused_kernel=False)
    p = pm.double()
    fused = p[..., 0, :, :].clone()
    fused[..., :3, :4] = p[..., 1, :3, :3] @ p[..., 0, :3, :4]
    prod = fused[..., 1:, :, :] @ torch.linalg.inv(fused[..., 0, :, :]).unsqueeze(-3)
    assert rot.shape == (3, 2, 5, 9) and trans.shape == (3, 2, 5, 3)
    np.testing.assert_allclose(npy(rot), prod[..., :3, :3].reshape(3, 2, 5, 9).float().cpu().numpy(), rtol=2e-7, atol=1e-9)
    np.testing.assert_allclose(npy(trans), prod[..., :3, 3].float().cpu().numpy(), rtol=2e-7, atol=1e-6)
    np.testing.assert_allclose(npy(rot), npy(rot32), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(npy(trans), npy(trans32), rtol=1e-4, atol=1e-3)


def test_cvp_depth_interval_kernel_vs_reference_op_sequence():
    """calDepthHypo's statistical interval (CVP-MVSNet/models/modules.py:146-209): the one-launch kernel against the
    reference's float64 torch op sequence (pyramid._mean_depth_interval, itself pinned by the cvp_network golden)."""
    from mvs_b200 import pyramid
    B, H, W = 2, 40, 56
    ref_in, src_in, ref_ex, src_ex = [cu(a) for a in cases.synth.cvp_cameras(2, W, seed=6, batch=B)]
    depth = cu(cases.synth.depth_surface(H, W, B) + np.random.RandomState(1).uniform(-3, 3, (B, H, W)).astype(np.float32))
    got = pyramid.mean_depth_interval_device(depth, ref_in, src_in[:, 0], ref_ex, src_ex[:, 0])
    ref = torch.stack([pyramid._mean_depth_interval(depth[b], ref_in[b].double(), src_in[b, 0].double(), ref_ex[b].double(),
                                                    src_ex[b, 0].double()) for b in range(B)])
    assert got.dtype == torch.float64 and got.shape == (B,)
    np.testing.assert_allclose(got.cpu().numpy(), ref.cpu().numpy(), rtol=1e-9)


@pytest.mark.parametrize("pixel", [False, True])
def test_homo_warping_dropin_vs_on_device_reference_ops(pixel):
    """The drop-in with DEVICE-computed projections (torch.inverse / matmul on the GPU, as a patched train.py / eval.py would
    run it) against the reference's own op sequence executed on the same GPU (ATen-CUDA grid_sample, TF32 off).  The
    authoritative oracle for bit-exactness is the reference's CPU path (tests above, fed identical rot / trans bits); ATen-CUDA
    itself deviates from ATen-CPU in the last bits of the sample position (scalar division as multiply-by-reciprocal, FMA
    contraction), so this comparison is to fp32 rounding of the blend, except where a position sits within rounding of an integer
    or a tap within a few ulps of the image border: those pixels are counted, not hidden."""
    from mvs_b200 import ops
    from oracle import torch_port as TP
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        c = cases.warp_pixel_case(B=2, C=8, H=96, W=128, D=12) if pixel else cases.warp_plane_case(B=2, C=8, H=96, W=128, D=12)
        src, sp, rp, d = cu(c["src_fea"]), cu(c["src_proj"]), cu(c["ref_proj"]), cu(c["depth"])
        ours = ops.homo_warping(src, sp, rp, d)
        ref = TP.warp_volume(src, sp, rp, d)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    diff = (ours - ref).abs()
    scale = float(ref.abs().max())
    bad = diff > 1e-4 * scale
    frac = float(bad.float().mean())
    print("on-device drop-in vs ATen-CUDA: max abs diff", float(diff.max()), "fraction beyond 1e-4 of max", frac)
    assert frac < 2e-4, frac
    assert float(diff[~bad].max()) <= 1e-4 * scale
