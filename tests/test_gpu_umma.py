"""GPU tests of the bf16 tcgen05 (UMMA) convolution path against the CPU oracle.
The oracle is fed the SAME bf16-rounded activations and weights, so the only differences are fp32
accumulation order and the single bf16 rounding of the stored output (rel 2^-8)."""
import numpy as np
import pytest
import torch

import cases
from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def bf(a):
    return torch.from_numpy(np.ascontiguousarray(a)).bfloat16().float().numpy()


LAYERS = [
    # cin, cout, stride, transposed, (D, H, W)
    (8, 8, 1, False, (3, 5, 40)),        # conv0 stage 3: Cin=8 tap pairing
    (16, 8, 1, False, (4, 9, 150)),      # conv0 stage 2, ragged W > 128
    (32, 8, 1, False, (2, 4, 130)),      # conv0 stage 1
    (8, 16, 2, False, (5, 7, 61)),       # conv1: stride 2, Cin=8, odd extents
    (16, 16, 1, False, (4, 6, 50)),      # conv2
    (16, 32, 2, False, (4, 6, 300)),     # conv3: stride 2, two M tiles after striding
    (32, 32, 1, False, (3, 5, 37)),      # conv4
    (32, 64, 2, False, (4, 8, 50)),      # conv5: two Cout tiles
    (64, 64, 1, False, (2, 3, 25)),      # conv6
    (64, 32, 2, True, (2, 3, 25)),       # conv7: transposed
    (32, 16, 2, True, (2, 5, 70)),       # conv9 (2*70 = 140 outputs per row)
    (16, 8, 2, True, (3, 4, 130)),       # conv11: two M tiles
    (8, 1, 1, False, (4, 6, 133)),       # prob: fp32 logits out
    (64, 32, 1, True, (2, 4, 20)),       # CVP conv5: transposed stride 1
    # T-merged stride-1 path (nine taps along N, step-axis reduction in the epilogue): several row blocks incl. a
    # ragged one, more steps than TMEM buffers, interior rows with all three row taps
    (32, 8, 1, False, (12, 21, 130)),
    (16, 8, 1, False, (11, 9, 64)),
    (8, 1, 1, False, (11, 19, 140)),
    (16, 16, 1, False, (7, 18, 100)),
    (32, 32, 1, False, (5, 11, 129)),
    (16, 16, 1, True, (6, 10, 70)),      # transposed stride 1 (flipped taps) on the T-merged path
]


@pytest.mark.parametrize("cin,cout,stride,transposed,dhw", LAYERS)
def test_conv3d_c8_layer(cin, cout, stride, transposed, dhw):
    from mvs_b200 import ops
    rng = np.random.RandomState(cin * 1000 + cout * 10 + stride + 2 * transposed)
    D, H, W = dhw
    B = 2
    x = bf(rng.standard_normal((B, cin, D, H, W)).astype(np.float32))
    wshape = (cin, cout, 3, 3, 3) if transposed else (cout, cin, 3, 3, 3)
    w = bf((rng.standard_normal(wshape) / np.sqrt(27 * cin)).astype(np.float32))
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    shift = (0.3 * rng.standard_normal(cout)).astype(np.float32)
    ref = O.conv3d(x, w, None, stride, transposed)
    ref = ref * scale.reshape(1, -1, 1, 1, 1) + shift.reshape(1, -1, 1, 1, 1)
    relu = cout != 1
    if relu:
        ref = np.maximum(ref, 0)
    skip = None
    if cout != 1:
        skip = bf(rng.standard_normal(ref.shape).astype(np.float32))
        ref = skip + ref
    packed = ops.pack_conv_weights(cu(w), stride, transposed)
    y = ops.conv3d_c8(ops.pack_c8(cu(x)), packed, cin, cout, cu(scale), cu(shift),
                      ops.pack_c8(cu(skip)) if skip is not None else None, stride, transposed, relu)
    torch.cuda.synchronize()
    if cout == 1:
        out = y.cpu().numpy()
        np.testing.assert_allclose(out, ref, rtol=1e-4, atol=1e-4)
    else:
        out = ops.unpack_c8(y, cout).cpu().numpy()
        assert out.shape == ref.shape
        np.testing.assert_allclose(out, ref, rtol=2 ** -7, atol=4e-3)


NOSKIP_LAYERS = [
    # cin, cout, (D, H, W): stride-1 layers WITHOUT a skip operand (conv0 / prob as CostRegNet runs them; test_conv3d_c8_layer adds one)
    (8, 8, (3, 5, 40)), (8, 8, (8, 40, 150)), (16, 8, (11, 9, 64)), (16, 8, (32, 13, 130)), (32, 8, (12, 21, 130)),
    (32, 8, (48, 7, 100)), (8, 1, (4, 6, 133)), (8, 1, (21, 19, 140)), (16, 1, (9, 11, 70)), (8, 8, (1, 9, 30)), (8, 8, (2, 3, 129)),
]


@pytest.mark.parametrize("cin,cout,dhw", NOSKIP_LAYERS)
def test_conv3d_c8_stride1_without_skip(cin, cout, dhw):
    """conv0 / prob shapes against the CPU oracle: odd row counts, row blocks that do not divide D, more steps than TMEM buffers,
    `prob`'s one-column packing with up to 15 rows per step."""
    from mvs_b200 import ops
    rng = np.random.RandomState(100 + cin + cout + dhw[0])
    D, H, W = dhw
    x = bf(rng.standard_normal((2, cin, D, H, W)).astype(np.float32))
    w = bf((rng.standard_normal((cout, cin, 3, 3, 3)) / np.sqrt(27 * cin)).astype(np.float32))
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    shift = (0.3 * rng.standard_normal(cout)).astype(np.float32)
    ref = O.conv3d(x, w, None, 1, False) * scale.reshape(1, -1, 1, 1, 1) + shift.reshape(1, -1, 1, 1, 1)
    relu = cout != 1
    if relu:
        ref = np.maximum(ref, 0)
    y = ops.conv3d_c8(ops.pack_c8(cu(x)), ops.pack_conv_weights(cu(w), 1, False), cin, cout, cu(scale), cu(shift), None, 1, False, relu)
    torch.cuda.synchronize()
    if cout == 1:
        np.testing.assert_allclose(y.cpu().numpy(), ref, rtol=1e-4, atol=1e-4)
    else:
        np.testing.assert_allclose(ops.unpack_c8(y, cout).cpu().numpy(), ref, rtol=2 ** -7, atol=4e-3)


def _dw_perm(W):
    """index list p with natural[..., w] == dw[..., p[w]]: column w sits at (w & 1) * ceil(W / 2) + (w >> 1)."""
    w = torch.arange(W)
    return ((w & 1) * ((W + 1) // 2) + (w >> 1)).to(DEV)


DW_LAYERS = [
    # which, cin, cout, stride, transposed, (D, H, W)
    ("y", 8, 8, 1, False, (3, 5, 41)),       # stride-1 layer writes DW (odd W: ceil(W/2) even columns)
    ("y", 32, 16, 1, False, (4, 9, 150)),    # two channel blocks, W > 128
    ("x", 8, 16, 2, False, (5, 7, 61)),      # stride-2 layer reads DW, odd extents
    ("x", 16, 32, 2, False, (4, 6, 300)),    # two M tiles after striding
    ("x", 32, 64, 2, False, (4, 8, 50)),     # two Cout tiles
    ("skip", 16, 8, 2, True, (3, 4, 130)),   # transposed layer adds a DW skip: Cout <= 8 epilogue
    ("skip", 64, 32, 2, True, (2, 3, 25)),   # ... Cout > 8 epilogue (two blocks per TMEM load)
]


@pytest.mark.parametrize("which,cin,cout,stride,transposed,dhw", DW_LAYERS)
def test_conv3d_c8_dw_layout(which, cin, cout, stride, transposed, dhw):
    """MVS_X_DW / MVS_Y_DW / MVS_SKIP_DW: the W-de-interleaved tensors are the natural ones permuted -- the layer's results
    must be bit-identical to the unflagged run (same MMAs, same epilogue arithmetic, different addresses)."""
    from mvs_b200 import ops, _lib as L
    rng = np.random.RandomState(17 + cin + cout)
    D, H, W = dhw
    x = ops.pack_c8(cu(bf(rng.standard_normal((2, cin, D, H, W)).astype(np.float32))))
    wshape = (cin, cout, 3, 3, 3) if transposed else (cout, cin, 3, 3, 3)
    packed = ops.pack_conv_weights(cu(bf((rng.standard_normal(wshape) / np.sqrt(27 * cin)).astype(np.float32))), stride, transposed)
    scale, shift = cu(rng.uniform(0.5, 1.5, cout).astype(np.float32)), cu((0.3 * rng.standard_normal(cout)).astype(np.float32))
    Wo = 2 * W if transposed else ((W - 1) // 2 + 1 if stride == 2 else W)
    skip = None
    if which == "skip":
        Do, Ho = 2 * D, 2 * H
        skip = ops.pack_c8(cu(bf(rng.standard_normal((2, cout, Do, Ho, Wo)).astype(np.float32))))
    ref = ops.conv3d_c8(x, packed, cin, cout, scale, shift, skip, stride, transposed, True)
    if which == "y":
        out = ops.conv3d_c8(x, packed, cin, cout, scale, shift, None, stride, transposed, True, layout=L.Y_DW)
        out = out.index_select(4, _dw_perm(Wo))                      # natural[w] = dw[p[w]]
    elif which == "x":
        x_dw = torch.empty_like(x)
        x_dw.index_copy_(4, _dw_perm(W), x)                          # dw[p[w]] = natural[w]
        out = ops.conv3d_c8(x_dw.contiguous(), packed, cin, cout, scale, shift, None, stride, transposed, True, layout=L.X_DW)
    else:
        s_dw = torch.empty_like(skip)
        s_dw.index_copy_(4, _dw_perm(Wo), skip)
        out = ops.conv3d_c8(x, packed, cin, cout, scale, shift, s_dw.contiguous(), stride, transposed, True, layout=L.SKIP_DW)
    torch.cuda.synchronize()
    assert out.shape == ref.shape
    assert torch.equal(out.view(torch.int16), ref.view(torch.int16))


@pytest.mark.parametrize("cin,cout,dhw", [(8, 8, (5, 9, 150)), (8, 8, (1, 7, 40)), (16, 16, (4, 6, 130)), (32, 8, (5, 5, 64)),
                                          (64, 32, (3, 4, 40))])
def test_conv3d_c8_kd1(cin, cout, dhw):
    """MVS_KD1 (weights zero outside kd = 1: D stacked images, the FeatureNet engine's form): the kernel skips the other depth
    taps' MMAs and the halo rows.  Only exact zeros leave the sums, but the taps pair up into K steps differently, so the fp32
    partial sums round differently: equal to the unflagged run within one fp16 ulp of the output."""
    from mvs_b200 import ops, _lib as L
    rng = np.random.RandomState(5 + cin + cout)
    D, H, W = dhw
    x = torch.from_numpy(rng.standard_normal((1, (cin + 7) // 8, D, H, W, 8)).astype(np.float32)).to(DEV).half()
    w = np.zeros((cout, cin, 3, 3, 3), np.float32)
    w[:, :, 1] = rng.standard_normal((cout, cin, 3, 3)) / np.sqrt(9 * cin)
    packed = ops.pack_conv_weights(cu(w), 1, False, act_f16=True)
    scale, shift = cu(rng.uniform(0.5, 1.5, cout).astype(np.float32)), cu((0.3 * rng.standard_normal(cout)).astype(np.float32))
    ref = ops.conv3d_c8(x, packed, cin, cout, scale, shift, None, 1, False, True, act_f16=True)
    out = ops.conv3d_c8(x, packed, cin, cout, scale, shift, None, 1, False, True, act_f16=True, layout=L.KD1)
    torch.cuda.synchronize()
    tol = 2 ** -10 * ref.float().abs().max().item()
    assert (out.float() - ref.float()).abs().max().item() <= tol
    assert (out != ref).float().mean().item() < 0.02             # and almost everywhere the same bits
    # and images do not mix: every depth slice equals the same layer run on that slice alone
    one = ops.conv3d_c8(x[:, :, :1].contiguous(), packed, cin, cout, scale, shift, None, 1, False, True, act_f16=True, layout=L.KD1)
    assert (one.float() - out[:, :, :1].float()).abs().max().item() <= tol


@pytest.mark.parametrize("cin,cout,nhw", [(8, 8, (3, 37, 150)), (8, 8, (2, 5, 40)), (8, 8, (1, 1, 9)), (8, 16, (2, 31, 129)),
                                          (16, 16, (2, 23, 130)), (32, 16, (3, 16, 64)), (32, 32, (2, 9, 70)), (64, 32, (2, 12, 40)),
                                          (32, 8, (2, 45, 300))])
def test_conv3d_c8_flat2d(cin, cout, nhw):
    """MVS_FLAT2D (plain 2D convolution, image rows tiled by the kernel) against the same layer run as a D = 1 volume: the
    same products summed in a different order (fp32 accumulators, one fp16 rounding of the output)."""
    from mvs_b200 import ops, _lib as L
    rng = np.random.RandomState(9 + cin + cout)
    N, H, W = nhw
    x = torch.from_numpy(rng.standard_normal((N, (cin + 7) // 8, 1, H, W, 8)).astype(np.float32)).to(DEV).half()
    w = np.zeros((cout, cin, 3, 3, 3), np.float32)
    w[:, :, 1] = rng.standard_normal((cout, cin, 3, 3)) / np.sqrt(9 * cin)
    w[:, :, 0] = 7.0                                     # must be ignored: only the centre depth slice is a 2D kernel
    w2 = w.copy(); w2[:, :, 0] = 0
    scale, shift = cu(rng.uniform(0.5, 1.5, cout).astype(np.float32)), cu((0.3 * rng.standard_normal(cout)).astype(np.float32))
    ref = ops.conv3d_c8(x, ops.pack_conv_weights(cu(w2), 1, False, act_f16=True), cin, cout, scale, shift, None, 1, False, True,
                        act_f16=True)
    out = ops.conv3d_c8(x, ops.pack_conv_weights(cu(w), 1, False, act_f16=True, flat2d=True), cin, cout, scale, shift, None, 1,
                        False, True, act_f16=True, layout=L.FLAT2D)
    torch.cuda.synchronize()
    assert out.shape == ref.shape
    err = (out.float() - ref.float()).abs().max().item()
    assert err <= 2 ** -9 * ref.float().abs().max().item() + 1e-3, err
    # and against the fp32 definition
    xr = x.float().permute(0, 1, 5, 2, 3, 4).reshape(N, -1, 1, H, W)[:, :cin, 0]
    y32 = torch.nn.functional.conv2d(xr, cu(w[:, :, 1]).half().float(), padding=1) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    y32 = torch.relu(y32)
    got = out.float().permute(0, 1, 5, 2, 3, 4).reshape(N, -1, 1, H, W)[:, :cout, 0]
    np.testing.assert_allclose(got.cpu().numpy(), y32.cpu().numpy(), rtol=2 ** -9, atol=3e-3)


def test_conv3d_c8_dw_flag_validation():
    from mvs_b200 import ops, _lib as L
    x = torch.zeros(1, 1, 2, 4, 16, 8, dtype=torch.bfloat16, device=DEV)
    w1 = ops.pack_conv_weights(torch.zeros(8, 8, 3, 3, 3, device=DEV), 1, False)
    w2 = ops.pack_conv_weights(torch.zeros(16, 8, 3, 3, 3, device=DEV), 2, False)
    with pytest.raises(L.MvsError):
        ops.conv3d_c8(x, w1, 8, 8, layout=L.X_DW)                    # stride-1 layers read natural order only
    with pytest.raises(L.MvsError):
        ops.conv3d_c8(x, w2, 8, 16, stride=2, layout=L.Y_DW)         # stride-2 layers write natural order only
    with pytest.raises(L.MvsError):
        ops.conv3d_c8(x, w1, 8, 8, layout=L.SKIP_DW)                 # no skip tensor


@pytest.mark.parametrize("family,cin", [("mvsnet", 32), ("cas", 16), ("cas", 8), ("cvp", 16)])
def test_costreg_fast_vs_strict(family, cin):
    """Whole CostRegNet on the tensor-core path vs the strict fp32 path (same weights)."""
    from mvs_b200 import modules, ops
    sd = cases.costreg_state(family, cin=cin, seed=60)
    def make(mode):
        net = {"mvsnet": lambda: modules.CostRegNetMVSNet(mode=mode), "cas": lambda: modules.CostRegNetCas(cin, 8, mode=mode),
               "cvp": lambda: modules.CostRegNetCVP(mode=mode)}[family]()
        net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
        return net.to(DEV).eval()
    x = torch.randn(1, cin, 8, 16, 136, device=DEV).bfloat16().float()
    with torch.no_grad():
        ref = make("strict")(x)
        out = make("fast")(ops.pack_c8(x))
    if family == "cvp":
        assert out.shape == ref.shape
    err = (out.float().reshape(ref.shape) - ref).abs().max().item()
    assert err <= 0.03 * ref.abs().max().item() + 1e-3, (err, ref.abs().max().item())


def test_cas_cascade_fast_vs_golden():
    """Full 3-stage cascade on the fast path (C8 bf16 builder + tcgen05 CostRegNet) vs the reference's
    fp32 forward: bf16 storage => looser tolerance, stated here: depth within 2e-3 relative."""
    from mvs_b200 import modules, cascade
    g = cases.golden("cas_cascade"); k = cases.cascade_case()
    regs = []
    for i, cin in enumerate((32, 16, 8)):
        net = modules.CostRegNet(cin, 8, mode="fast")
        net.load_state_dict({kk: torch.from_numpy(np.asarray(v)) for kk, v in cases.costreg_state("cas", cin=cin, seed=14 + i).items()}, strict=True)
        regs.append(net.to(DEV).eval())
    n_views = k["feats"]["stage1"].shape[0]
    feats = [{s: cu(k["feats"][s][v]) for s in k["feats"]} for v in range(n_views)]
    with torch.no_grad():
        out = cascade.cascade_hot_path(feats, {s: cu(p) for s, p in k["projs"].items()}, cu(k["depth_values"]), regs,
                                       ndepths=k["ndepths"], img_hw=(k["H"], k["W"]))
    for s in ("stage1", "stage2", "stage3"):
        rel = np.abs(out[s]["depth"].cpu().numpy() - g[s + "_depth"]) / g[s + "_depth"]
        assert rel.max() < 2e-3, (s, rel.max())


def test_cfg2_mvsnet_full_size_fast_vs_strict():
    """BASELINE cfg2 shape (MVSNet, features 32x128x160, N=5, D=192, B>1): whole hot path on the fast
    (C8 bf16 + tcgen05) path vs the strict fp32 path on the same bf16-representable inputs."""
    from mvs_b200 import modules
    B, N, D, H, W = 2, 5, 192, 128, 160
    sd = cases.costreg_state("mvsnet", seed=70)
    def make(mode):
        net = modules.CostRegNetMVSNet(mode=mode)
        net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
        return net.to(DEV).eval()
    feats = [cu(f).bfloat16().float() for f in cases.synth.features(N, 32, H, W, 71, B)]
    proj = cu(cases.synth.proj_matrices(N, W, 72, B))
    depth = cu(np.tile((425.0 + 2.65 * np.arange(D)).astype(np.float32), (B, 1)))
    with torch.no_grad():
        ref = modules.mvsnet_hot_path(feats, proj, depth, make("strict"))
        out = modules.mvsnet_hot_path(feats, proj, depth, make("fast"))
    rel = ((out["depth"] - ref["depth"]).abs() / ref["depth"]).max().item()
    assert rel < 5e-3, rel
    assert (out["photometric_confidence"] - ref["photometric_confidence"]).abs().max().item() < 0.05


def test_cfg5_builder_seven_views_full_size():
    """BASELINE cfg5 stage-3 shape (C=8, D=8, 1056x1920, N=7 => 6 source views): C8 builder vs the
    strict builder on bf16-representable features (one bf16 rounding of the result + fp32 reassociation)."""
    from mvs_b200 import ops
    N, C, D, H, W = 7, 8, 8, 1056, 1920
    torch.manual_seed(5)
    feats = [torch.randn(1, C, H, W, device=DEV).bfloat16().float() for _ in range(N)]
    p = torch.from_numpy(cases.synth.proj_matrices(N, W, 73, 1))
    prod = torch.stack([p[:, i] @ torch.inverse(p[:, 0]) for i in range(1, N)], 1)
    rots = [prod[:, i, :3, :3].reshape(1, 9).contiguous().to(DEV) for i in range(N - 1)]
    trs = [prod[:, i, :3, 3].contiguous().to(DEV) for i in range(N - 1)]
    depth = cu(cases.synth.depth_per_pixel(D, H, W, 2.65, 1))
    strict = ops.cost_volume(feats[0], feats[1:], rots, trs, depth)
    fast = ops.unpack_c8(ops.cost_volume_c8(ops.pack_c8(feats[0]), [ops.pack_c8(f) for f in feats[1:]], rots, trs, depth), C)
    assert ((fast - strict).abs() <= strict.abs() * 2 ** -7 + 2e-3).all()


def test_graphed_step_replays_the_cascade():
    """mvs_b200.GraphedStep: the whole 3-stage fast cascade captured in one CUDA graph replays bit-identically to the
    eager launches, also after the static inputs were updated in place (the C-ABI only enqueues on the given stream)."""
    from mvs_b200 import modules, cascade
    from mvs_b200.graph import GraphedStep
    k = cases.cascade_case()
    regs = []
    for i, cin in enumerate((32, 16, 8)):
        net = modules.CostRegNet(cin, 8, mode="fast")
        net.load_state_dict({kk: torch.from_numpy(np.asarray(v)) for kk, v in cases.costreg_state("cas", cin=cin, seed=14 + i).items()}, strict=True)
        regs.append(net.to(DEV).eval())
    n_views = k["feats"]["stage1"].shape[0]
    feats = [{s: cu(k["feats"][s][v]) for s in k["feats"]} for v in range(n_views)]
    projs = {s: cu(p) for s, p in k["projs"].items()}
    dv = cu(k["depth_values"])
    dmin, dmax = float(k["depth_values"][0, 0]), float(k["depth_values"][0, -1])
    run = lambda: cascade.cascade_hot_path(feats, projs, dv, regs, ndepths=k["ndepths"], img_hw=(k["H"], k["W"]),
                                           depth_min=dmin, depth_max=dmax)
    with torch.no_grad():
        eager = {s: run()[s]["depth"].clone() for s in ("stage1", "stage2", "stage3")}
        g = GraphedStep(run)
        out = g()
        torch.cuda.synchronize()
        for s in eager:
            assert torch.equal(out[s]["depth"], eager[s]), s
        # new inputs, in place: replay must follow them
        for f in feats:
            for s in f:
                f[s].mul_(0.5)
        eager2 = run()["stage3"]["depth"].clone()
        out2 = g()
        torch.cuda.synchronize()
        assert torch.equal(out2["stage3"]["depth"], eager2)
        assert not torch.equal(eager2, eager["stage3"])
