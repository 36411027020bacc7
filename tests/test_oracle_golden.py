"""Pins the CPU oracle (oracle/mvs_oracle.c) against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  Warp / cost volume: bit-exact.  Conv / softmax: rounding-level."""
import numpy as np
import pytest
import torch

import cases
from oracle import oracle as O


def rt(proj):
    """rot [.,3,3], trans [.,3] from the stored torch-computed src@inv(ref) product."""
    return np.ascontiguousarray(proj[..., :3, :3]), np.ascontiguousarray(proj[..., :3, 3])


def compose(proj):  # MVSNet-style proj [B,V,4,4] -> products for views 1.. relative to view 0
    p = torch.from_numpy(proj)
    return torch.stack([p[:, i] @ torch.inverse(p[:, 0]) for i in range(1, p.shape[1])], 1).numpy()


def assert_bitexact(a, b):
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape
    same = (a.view(np.uint32) == b.view(np.uint32)) | (a == b)
    assert same.all(), f"{(~same).sum()} of {same.size} differ, max abs {np.abs(a - b).max()}"


def test_warp_plane_bitexact():
    g = cases.golden("mvsnet_warp_plane"); c = cases.warp_plane_case()
    rot, tr = rt(g["proj"])
    assert_bitexact(O.homo_warp(c["src_fea"], rot, tr, c["depth"]), g["out"])
    # the stored product is what torch recomputes here (same build) -- keeps GPU-box tests honest
    assert np.array_equal(compose(np.stack([c["ref_proj"], c["src_proj"]], 1))[:, 0], g["proj"])


def test_warp_pixel_bitexact():
    g = cases.golden("cas_warp_pixel"); c = cases.warp_pixel_case()
    rot, tr = rt(g["proj"])
    assert_bitexact(O.homo_warp(c["src_fea"], rot, tr, c["depth"]), g["out"])


def test_warp_pl_bitexact():
    g = cases.golden("pl_homo_warp"); c = cases.warp_plane_case(seed=9)
    rot, tr = rt(g["proj"])
    assert_bitexact(O.homo_warp(c["src_fea"], rot, tr, c["depth"], O.ALIGN_CORNERS | O.PL_ORDER), g["out"])


def cvp_products(c):
    """CVP composes K@E[:3] + [0,0,0,1] itself (CVP-MVSNet/models/modules.py:90-97)."""
    ref_in, src_in, ref_ex, src_ex = (torch.from_numpy(c[k]) for k in ("ref_in", "src_in", "ref_ex", "src_ex"))
    B = ref_in.shape[0]
    last = torch.tensor([[[0, 0, 0, 1.0]]]).repeat(B, 1, 1)
    ref_proj = torch.cat((torch.matmul(ref_in, ref_ex[:, 0:3, :]), last), 1)
    out = []
    for s in range(src_in.shape[1]):
        src_proj = torch.cat((torch.matmul(src_in[:, s], src_ex[:, s, 0:3, :]), last), 1)
        out.append(torch.matmul(src_proj, torch.inverse(ref_proj)))
    return torch.stack(out, 1).numpy()


def test_warp_cvp_bitexact():
    g = cases.golden("cvp_warp"); c = cases.cvp_case(per_pixel=False)
    rot, tr = rt(cvp_products(c)[:, 0])
    assert_bitexact(O.homo_warp(c["feats"][1], rot, tr, c["depth"]), g["out"])


def test_cost_volume_mvsnet_bitexact():
    g = cases.golden("mvsnet_forward"); v = cases.volume_case(n_views=4, C=32, H=16, W=24, D=8, seed=3)
    rot, tr = rt(g["proj"])
    assert_bitexact(O.cost_volume(v["feats"][0], v["feats"][1:], rot, tr, v["depth"]), g["var"])


@pytest.mark.parametrize("tag,cin,pp,seed", [("s2", 16, True, 4), ("s1", 32, False, 8)])
def test_cost_volume_cas_bitexact(tag, cin, pp, seed):
    g = cases.golden(f"cas_depthnet_{tag}"); c = cases.cas_case(C=cin, per_pixel=pp, seed=seed)
    rot, tr = rt(g["proj"])
    assert_bitexact(O.cost_volume(c["feats"][0], c["feats"][1:], rot, tr, c["depth"]), g["var"])


def test_cost_volume_cvp_quirk_bitexact():
    g = cases.golden("cvp_proj_cost"); c = cases.cvp_case(per_pixel=True)
    rot, tr = rt(cvp_products(c))
    out = O.cost_volume(c["feats"][0], c["feats"][1:], rot, tr, c["depth"], O.REF_SUM_SQUARED)
    assert_bitexact(out, g["out"])
    plain = O.cost_volume(c["feats"][0], c["feats"][1:], rot, tr, c["depth"])
    assert not np.array_equal(plain, g["out"])      # the aliasing quirk is observable


def close(a, b, rtol, atol):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def test_costreg_mvsnet():
    x = np.random.RandomState(21).standard_normal((1, 32, 8, 16, 24)).astype(np.float32)
    out = O.costreg_mvsnet(x, cases.costreg_state("mvsnet", seed=11))
    close(out, cases.golden("mvsnet_costreg")["out"], 1e-4, 2e-5)


@pytest.mark.parametrize("cin", [32, 16, 8])
def test_costreg_cas(cin):
    x = np.random.RandomState(22 + cin).standard_normal((1, cin, 8, 16, 24)).astype(np.float32)
    out = O.costreg_cas(x, cases.costreg_state("cas", cin=cin, seed=13))
    close(out, cases.golden(f"cas_costreg_c{cin}")["out"], 1e-4, 2e-5)


def test_costreg_cvp():
    x = np.random.RandomState(23).standard_normal((1, 16, 8, 16, 24)).astype(np.float32)
    out = O.costreg_cvp(x, cases.costreg_state("cvp", seed=15))
    close(out, cases.golden("cvp_costreg")["out"], 1e-4, 2e-5)


def test_regress_conf():
    g = cases.golden("mvsnet_regress"); c = cases.logits_case()
    r = O.softargmin_conf(c["logits"], c["depth"], clamp_index=False, want_prob=True)
    close(r["prob"], g["prob"], 1e-5, 1e-7)
    close(r["depth"], g["depth"], 1e-6, 0)
    close(r["expect_idx"], g["expect_idx"], 1e-5, 1e-5)
    stable = np.abs(g["expect_idx"] - np.round(g["expect_idx"])) > 1e-3
    assert (r["index"] == g["index"])[stable].all()
    close(r["conf"][stable], g["conf"][stable], 1e-5, 1e-6)


def test_mvsnet_forward_chain():
    """variance -> CostRegNet -> softmax/regression chained in the oracle vs MVSNet.forward."""
    g = cases.golden("mvsnet_forward"); v = cases.volume_case(n_views=4, C=32, H=16, W=24, D=8, seed=3)
    rot, tr = rt(g["proj"])
    var = O.cost_volume(v["feats"][0], v["feats"][1:], rot, tr, v["depth"])
    logits = O.costreg_mvsnet(var, cases.costreg_state("mvsnet", seed=11))
    close(logits, g["logits"], 1e-4, 2e-5)
    r = O.softargmin_conf(logits[:, 0], v["depth"])
    close(r["depth"], g["depth"], 1e-5, 0)
    close(r["conf"], g["conf"], 1e-4, 1e-5)


@pytest.mark.parametrize("tag,cin,pp,seed", [("s2", 16, True, 4), ("s1", 32, False, 8)])
def test_cas_depthnet_chain(tag, cin, pp, seed):
    g = cases.golden(f"cas_depthnet_{tag}"); c = cases.cas_case(C=cin, per_pixel=pp, seed=seed)
    rot, tr = rt(g["proj"])
    var = O.cost_volume(c["feats"][0], c["feats"][1:], rot, tr, c["depth"])
    logits = O.costreg_cas(var, cases.costreg_state("cas", cin=cin, seed=12))
    close(logits, g["logits"], 1e-4, 2e-5)
    r = O.softargmin_conf(logits[:, 0], c["depth"], clamp_index=True)
    close(r["depth"], g["depth"], 1e-5, 0)
    close(r["conf"], g["conf"], 1e-4, 1e-5)


def test_range_samples_bitexact():
    cur = np.random.RandomState(31).uniform(500, 800, (2, 12, 20)).astype(np.float32)
    assert_bitexact(O.depth_range_samples(cur, 5.3, 8), cases.golden("cas_range_samples")["out"])


def test_known_answers():
    """Synthetic known-answer cases (SURVEY.md §8(c) 'what pins it for us')."""
    H, W, C, D = 12, 16, 3, 4
    fea = np.random.RandomState(5).standard_normal((1, C, H, W)).astype(np.float32)
    eye, zero = np.eye(3, dtype=np.float32)[None], np.zeros((1, 3), np.float32)
    depth = cases.synth.depth_planes(D)
    # identity pose + align_corners=True => the warp is the identity for every plane
    out = O.homo_warp(fea, eye, zero, depth, O.ALIGN_CORNERS)
    for d in range(D):
        np.testing.assert_allclose(out[:, :, d], fea, rtol=0, atol=2e-5)
    # all views equal + identity => variance == 0 (up to rounding)
    var = O.cost_volume(fea, np.stack([fea, fea]), np.tile(eye, (1, 2, 1, 1)), np.tile(zero, (1, 2, 1)), depth,
                        O.ALIGN_CORNERS)
    assert np.abs(var).max() < 1e-5
    # one-hot logits => depth_regression returns that plane, confidence 1
    logits = np.full((1, D, 2, 2), -1e4, np.float32); logits[:, 2] = 0
    r = O.softargmin_conf(logits, depth)
    assert np.allclose(r["depth"], depth[0, 2]) and np.allclose(r["conf"], 1.0) and (r["index"] == 2).all()
    # non-finite sample positions propagate NaN like the ATen CPU kernel (P_z == 0)
    bad = O.homo_warp(fea, np.zeros((1, 3, 3), np.float32), zero, depth)
    assert np.isnan(bad).all()
