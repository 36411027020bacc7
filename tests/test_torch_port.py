"""The torch-CPU port (oracle/torch_port.py, the CPU baseline of bench.py) reproduces the golden
fixtures of the unmodified reference: bit-exact for warp / variance (identical op sequence)."""
import numpy as np
import torch

import cases
from oracle import torch_port as TP


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def sdt(sd):
    return {k: t(v) for k, v in sd.items()}


def test_port_warp_bitexact():
    g = cases.golden("mvsnet_warp_plane"); c = cases.warp_plane_case()
    out = TP.warp_volume(t(c["src_fea"]), t(c["src_proj"]), t(c["ref_proj"]), t(c["depth"]))
    assert np.array_equal(out.numpy(), g["out"])


def test_port_mvsnet_forward():
    g = cases.golden("mvsnet_forward"); v = cases.volume_case(n_views=4, C=32, H=16, W=24, D=8, seed=3)
    feats = [t(f) for f in v["feats"]]
    projs = torch.unbind(t(v["proj"]), 1)
    with torch.no_grad():
        var = TP.variance_volume(feats[0], feats[1:], projs[0], projs[1:], t(v["depth"]))
        assert np.array_equal(var.numpy(), g["var"])
        logits = TP.costreg(var, sdt(cases.costreg_state("mvsnet", seed=11)), "mvsnet")
        depth, conf = TP.regress(logits, t(v["depth"]), False)
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(depth.numpy(), g["depth"], rtol=1e-6)
    np.testing.assert_allclose(conf.numpy(), g["conf"], rtol=1e-5, atol=1e-6)


def test_port_cascade():
    g = cases.golden("cas_cascade"); k = cases.cascade_case()
    sds = [sdt(cases.costreg_state("cas", cin=cin, seed=14 + i)) for i, cin in enumerate((32, 16, 8))]
    n_views = k["feats"]["stage1"].shape[0]
    feats = [{s: t(k["feats"][s][v]) for s in k["feats"]} for v in range(n_views)]
    with torch.no_grad():
        out = TP.cas_cascade(feats, {s: t(p) for s, p in k["projs"].items()}, t(k["depth_values"]), sds,
                             ndepths=k["ndepths"], img_hw=(k["H"], k["W"]))
    for s in ("stage1", "stage2", "stage3"):
        np.testing.assert_allclose(out[s]["depth"].numpy(), g[s + "_depth"], rtol=1e-5)
        np.testing.assert_allclose(out[s]["photometric_confidence"].numpy(), g[s + "_conf"], rtol=1e-4, atol=1e-5)


def test_costreg_state_dict_keys_match_reference_shapes():
    """mvs_b200's CostRegNet mirrors load reference-keyed state dicts with strict=True (no GPU needed)."""
    from mvs_b200 import modules
    for fam, net in (("mvsnet", modules.CostRegNet()), ("cas", modules.CostRegNet(16, 8)), ("cvp", modules.CostRegNetCVP())):
        want = cases.costreg_shapes(fam, 16, 8)
        have = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        assert have == {k: tuple(s) for k, s in want.items()}, fam


def test_featurenet_port_and_mirror_match_reference():
    """The reference's own FeatureNet output (tests/golden/cas_featurenet.npz) vs (a) the torch port the CPU arm times
    and (b) the repo's state-dict-compatible mirror in strict mode (both the same ATen op sequence: bit-exact)."""
    from mvs_b200.featurenet import FeatureNet
    gold = cases.golden("cas_featurenet")
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in cases.featurenet_state(33).items()}
    img = torch.from_numpy(cases.synth.images_u8(2, 48, 80, seed=21)[0])
    with torch.no_grad():
        port = TP.featurenet(img.float() / 255.0, sd)
        net = FeatureNet(mode="strict").eval()
        net.load_state_dict(sd, strict=True)                     # the reference's exact key set
        mirror = net(img)                                        # uint8 in: normalised like the loader
    for k in ("stage1", "stage2", "stage3"):
        assert np.array_equal(port[k].numpy(), gold[k]), k
        assert np.array_equal(mirror[k].numpy(), gold[k]), k


def test_full_model_port_matches_reference():
    """CascadeMVSNet.forward from images, nothing stubbed (tests/golden/cas_full_model.npz)."""
    gold = cases.golden("cas_full_model")
    k = cases.full_model_case()
    sd = {n: torch.from_numpy(np.asarray(v)) for n, v in cases.full_model_state().items()}
    with torch.no_grad():
        out = TP.cas_model(torch.from_numpy(k["imgs_u8"]).float() / 255.0, {s: torch.from_numpy(p) for s, p in k["projs"].items()},
                           torch.from_numpy(k["depth_values"]), sd, ndepths=k["ndepths"])
    for s in ("stage1", "stage2", "stage3"):
        np.testing.assert_allclose(out[s]["depth"].numpy(), gold[s + "_depth"], rtol=1e-6)
        np.testing.assert_allclose(out[s]["photometric_confidence"].numpy(), gold[s + "_conf"], rtol=1e-5, atol=1e-6)


def test_mvsnet_model_port_and_mirror_keys():
    """Whole MVSNet.forward from images (tests/golden/mvsnet_full_model.npz): the torch port the CPU arm would time, and the
    mirror's state-dict key set / strict-mode extractor (same ATen sequence: bit-exact)."""
    from mvs_b200.mvsnet import MVSNet
    gold = cases.golden("mvsnet_full_model")
    k = cases.mvsnet_model_case()
    sd = sdt(cases.mvsnet_model_state())
    imgs = t(k["imgs_u8"]).float() / 255.0
    with torch.no_grad():
        out = TP.mvsnet_model(imgs, t(k["proj"]), t(k["depth"]), sd)
        m = MVSNet(mode="strict").eval()
        m.load_state_dict(sd, strict=True)
        feat = m.feature(t(k["imgs_u8"])[:, 1])                     # uint8 in: normalised like the loader
    np.testing.assert_allclose(out["depth"].numpy(), gold["depth"], rtol=1e-6)
    np.testing.assert_allclose(out["photometric_confidence"].numpy(), gold["conf"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(feat.numpy(), gold["feature_view1"])
