"""§8(f) f4 -- geometric-consistency filter: the oracle against the reference (+ real cv2) golden fixture on the CPU,
the CUDA kernels against the oracle and the fixture on the GPU."""
import numpy as np
import pytest

import cases
from oracle import geo_oracle as G


def _views(g, v):
    return g["depth"][0], g["K"][0], g["E"][0], g["depth"][v], g["K"][v], g["E"][v]


def test_oracle_matches_reference_with_cv2():
    """Bit-exact: the NumPy restatement (incl. the cv2.remap emulation) vs the reference's own functions run with cv2."""
    g, gold = cases.geo_case(), cases.golden("geo_filter")
    for v in range(1, g["depth"].shape[0]):
        d_rep, x_rep, y_rep, x_src, y_src = G.reproject_with_depth(*_views(g, v))
        with np.errstate(divide="ignore", invalid="ignore"):
            mask, d_masked, _, _ = G.check_geometric_consistency(*_views(g, v))
        for name, a in (("depth_reprojected", d_rep), ("x_reprojected", x_rep), ("y_reprojected", y_rep), ("x_src", x_src),
                        ("y_src", y_src), ("mask", mask), ("depth_masked", d_masked)):
            assert np.array_equal(a, gold[f"{name}_{v}"], equal_nan=True), (name, v)
        assert 0.3 < mask.mean() < 0.95           # the case exercises both outcomes
    with np.errstate(divide="ignore", invalid="ignore"):
        s, avg, _, gm, fm = G.fuse_ref_view(g["depth"][0], g["conf"], g["K"][0], g["E"][0], g["depth"][1:], g["K"][1:], g["E"][1:])
    assert np.array_equal(s, gold["geo_mask_sum"]) and np.array_equal(avg, gold["depth_est_averaged"], equal_nan=True)
    assert np.array_equal(gm, gold["geo_mask"]) and np.array_equal(fm, gold["final_mask"])


def test_c_oracle_fma_chains_match_reference_bit_exact():
    """The per-pixel C restatement (explicit k-ordered FMA chains = what the CUDA kernel computes, oracle/mvs_oracle.c
    mvso_geo_pair) reproduces the reference + cv2 golden bit for bit: every float output and every mask byte."""
    from oracle import oracle as O
    from mvs_b200.fusion import camera_block
    g, gold = cases.geo_case(), cases.golden("geo_filter")
    for v in range(1, g["depth"].shape[0]):
        with np.errstate(divide="ignore", invalid="ignore"):
            out = O.geo_pair(g["depth"][0], g["depth"][v], camera_block(g["K"][0], g["E"][0], g["K"][v], g["E"][v]))
        for name in ("depth_reprojected", "x_reprojected", "y_reprojected", "x_src", "y_src"):
            assert np.array_equal(out[name], gold[f"{name}_{v}"], equal_nan=True), (name, v)
        assert np.array_equal(out["mask"].astype(bool), gold[f"mask_{v}"]), v


def test_remap_emulation_known_answers():
    src = np.arange(12, dtype=np.float32).reshape(3, 4)
    x = np.array([[0.0, 1.5, 3.0, -1.0, 3.5, 1.0 + 1 / 64]], np.float32)
    y = np.array([[0.0, 0.5, 2.0, 0.0, 2.5, 1.0]], np.float32)
    out = G.remap_bilinear(src, x, y)
    # integer position, centre of four pixels, last pixel, outside, half outside (border 0), 1/64 px rounds to even (=> 0)
    assert out[0, 0] == 0 and out[0, 1] == (1 + 2 + 5 + 6) / 4 and out[0, 2] == 11 and out[0, 3] == 0
    assert out[0, 4] == 11 / 4 and out[0, 5] == 5


@pytest.mark.gpu
def test_gpu_pair_matches_oracle_and_golden():
    import torch
    from mvs_b200 import fusion
    g, gold = cases.geo_case(), cases.golden("geo_filter")
    for v in range(1, g["depth"].shape[0]):
        d_rep, x_rep, y_rep, x_src, y_src = fusion.reproject_with_depth(*_views(g, v))          # NumPy in -> NumPy out
        assert isinstance(d_rep, np.ndarray) and d_rep.dtype == np.float32
        # float64 FMA chains in the dgemm order + explicit non-contracted products: BIT-EXACT vs the reference + cv2
        for name, a in (("depth_reprojected", d_rep), ("x_reprojected", x_rep), ("y_reprojected", y_rep), ("x_src", x_src),
                        ("y_src", y_src)):
            assert np.array_equal(a, gold[f"{name}_{v}"], equal_nan=True), (name, v, int((a != gold[f"{name}_{v}"]).sum()))
        mask, d_masked, xs, ys = fusion.check_geometric_consistency(*_views(g, v))
        assert mask.dtype == np.bool_
        assert np.array_equal(mask, gold[f"mask_{v}"]), int((mask != gold[f"mask_{v}"]).sum())
        assert np.array_equal(d_masked, gold[f"depth_masked_{v}"], equal_nan=True)
        # CUDA tensors in -> CUDA tensors out
        t = [torch.from_numpy(np.ascontiguousarray(a)).cuda() if a.ndim == 2 and a.dtype == np.float32 else a for a in _views(g, v)]
        m2 = fusion.check_geometric_consistency(*t)[0]
        assert m2.is_cuda and np.array_equal(m2.cpu().numpy(), mask)


@pytest.mark.gpu
def test_gpu_fused_view_matches_golden():
    from mvs_b200 import fusion
    g, gold = cases.geo_case(), cases.golden("geo_filter")
    out = fusion.fuse_ref_view(g["depth"][0], g["conf"], g["K"][0], g["E"][0], list(g["depth"][1:]), list(g["K"][1:]),
                               list(g["E"][1:]), per_source=True)
    assert np.array_equal(out["geo_mask_sum"], gold["geo_mask_sum"])                    # byte / integer work: bit-exact
    assert out["depth_est_averaged"].dtype == np.float64
    assert np.array_equal(out["depth_est_averaged"], gold["depth_est_averaged"], equal_nan=True)
    assert np.array_equal(out["final_mask"], gold["final_mask"])
    assert np.array_equal(out["geo_mask"], gold["geo_mask"])
    for v in range(4):
        assert np.array_equal(out["all_srcview_geomask"][v], gold[f"mask_{v + 1}"]), v


@pytest.mark.gpu
def test_gpu_full_size_properties():
    """BASELINE cfg3 image size (1184x1600), 10 source views: identical cameras + identical depth maps => every pixel with
    positive depth is consistent with every view and the fused depth equals the input (a size-independent property)."""
    import torch
    from mvs_b200 import fusion
    g = cases.geo_case(n_src=1, H=1184, W=1600)
    d = torch.from_numpy(g["depth"][0]).cuda()
    conf = torch.ones_like(d)
    out = fusion.fuse_ref_view(d, conf, g["K"][0], g["E"][0], [d] * 10, [g["K"][0]] * 10, [g["E"][0]] * 10)
    ok = d > 0
    assert bool((out["geo_mask_sum"][ok] == 10).all())
    assert torch.allclose(out["depth_est_averaged"][ok], d[ok].double(), rtol=1e-6)
    assert bool((out["final_mask"] == ok).all())


@pytest.mark.gpu
def test_gpu_backproject_matches_oracle():
    from mvs_b200 import fusion
    g, gold = cases.geo_case(), cases.golden("geo_filter")
    ref = G.backproject(gold["depth_est_averaged"], gold["final_mask"], g["K"][0], g["E"][0])
    pts = fusion.backproject(gold["depth_est_averaged"], gold["final_mask"], g["K"][0], g["E"][0])
    assert pts.shape == ref.shape and pts.dtype == np.float32
    np.testing.assert_allclose(pts, ref.astype(np.float32), rtol=2e-7, atol=1e-4)


@pytest.mark.gpu
def test_gpu_full_size_vs_numpy_time(capsys):
    """cfg3 image size, 4 source views: GPU fused filter vs the NumPy restatement -- masks agree up to threshold ties, and the
    two wall times are printed side by side (the reference's filter_depth spends seconds per view here)."""
    import time
    import torch
    from mvs_b200 import fusion
    g = cases.geo_case(n_src=4, H=1184, W=1600)
    t0 = time.perf_counter()
    with np.errstate(divide="ignore", invalid="ignore"):
        s, avg, _, gm, fm = G.fuse_ref_view(g["depth"][0], g["conf"], g["K"][0], g["E"][0], g["depth"][1:], g["K"][1:], g["E"][1:])
    cpu_s = time.perf_counter() - t0
    d = [torch.from_numpy(x).cuda() for x in g["depth"]]
    conf = torch.from_numpy(g["conf"]).cuda()
    fn = lambda: fusion.fuse_ref_view(d[0], conf, g["K"][0], g["E"][0], d[1:], list(g["K"][1:]), list(g["E"][1:]))
    out = fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
    gpu_s = time.perf_counter() - t0
    assert (out["geo_mask_sum"].cpu().numpy() != s).mean() < 1e-4
    assert (out["final_mask"].cpu().numpy() != fm).mean() < 1e-4
    with capsys.disabled():
        print(f"\n[geo filter 1184x1600, 4 sources] NumPy restatement {cpu_s:.2f} s, mvs_geo_fuse {1e3 * gpu_s:.2f} ms")


@pytest.mark.gpu
def test_gpu_filter_depth_driver_matches_reference_loop(tmp_path):
    """`filter_depth` end to end in memory (every view as reference view once, MVSNet/eval.py:212-326 minus the .pfm round
    trip): vertices / colours / PLY against the NumPy restatement of the reference loop, bit for bit."""
    import torch
    from mvs_b200 import fusion, io as mio
    g = cases.geo_case()
    n = g["depth"].shape[0]
    rng = np.random.RandomState(11)
    imgs = [rng.uniform(0, 1, (*g["depth"].shape[1:], 3)).astype(np.float32) for _ in range(n)]
    confs = [np.clip(g["conf"] + 0.05 * v, 0, 1).astype(np.float32) for v in range(n)]
    pairs = [(r, [s for s in range(n) if s != r]) for r in range(n)]
    ply = tmp_path / "fused.ply"
    # CUDA tensors in (as cascade_hot_path hands them over): nothing touches the disk before the PLY
    v, c, per_view = fusion.filter_depth([torch.from_numpy(d).cuda() for d in g["depth"]], [torch.from_numpy(x).cuda() for x in confs],
                                         list(g["K"]), list(g["E"]), pairs, images=imgs, plyfilename=str(ply))
    ref_v, ref_c = [], []
    for r, srcs in pairs:
        with np.errstate(divide="ignore", invalid="ignore"):
            _, avg, _, _, fm = G.fuse_ref_view(g["depth"][r], confs[r], g["K"][r], g["E"][r], [g["depth"][s] for s in srcs],
                                               [g["K"][s] for s in srcs], [g["E"][s] for s in srcs])
        ref_v.append(G.backproject(avg, fm, g["K"][r], g["E"][r]).astype(np.float32))
        ref_c.append((imgs[r][fm] * 255).astype(np.uint8))
    ref_v, ref_c = np.concatenate(ref_v), np.concatenate(ref_c)
    assert v.shape == ref_v.shape and len(v) > 1000
    assert np.array_equal(c, ref_c)
    np.testing.assert_allclose(v, ref_v, rtol=3e-7, atol=1e-4)          # float64 chains, float32 store (dgemm order on the host side)
    pv, pc = mio.read_ply(str(ply))
    assert np.array_equal(pv, v) and np.array_equal(pc, c)
