#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/*.npz by running the UNMODIFIED reference modules
(imported from /root/reference) on the seeded inputs of tests/cases.py.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py            # all projects, one subprocess each
    python tests/golden/make_golden.py mvsnet     # one project

Harness-level shims (none touches the hot-path code under test):
  * each project is imported in its own interpreter: all of them call their package `models`;
  * the 2D FeatureNet / FeaturePyramid (not on the hot path) is swapped for a stub that returns
    the seeded synthetic feature maps, so fixtures do not have to store feature tensors;
  * CVP-MVSNet: `torch.Tensor.cuda` -> identity and `pdb.set_trace` -> no-op (hard-coded .cuda()
    at modules.py:78,92,137,157,237 and a stray breakpoint at net.py:157); depth_min/max are passed
    as float64 tensors as DataLoader collation produces;
  * MVSNet_pl: `kornia.utils.create_meshgrid` and `inplace_abn.InPlaceABN` are absent from this
    image; stand-ins with the documented semantics are injected into sys.modules so that the
    reference's own `homo_warp` (MVSNet_pl/models/modules.py:25-62) runs unmodified.
Fixtures store outputs (+ the torch-computed projection products the oracle consumes).
"""
from __future__ import annotations

import os
import subprocess
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"

warnings.filterwarnings("ignore")


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print(f"  wrote {name}.npz  ({os.path.getsize(path) / 1024:.0f} KiB)")


def t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a))


def load_sd(module, sd_np):
    import torch
    module.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()}, strict=True)


def proj_product(src_proj, ref_proj):
    """proj = src @ inverse(ref) exactly as the reference computes it (module.py:63), on CPU."""
    import torch
    return torch.matmul(t(src_proj), torch.inverse(t(ref_proj))).numpy()


class Replay:
    """Stand-in for the 2D feature extractor: returns preset outputs call by call."""

    def __init__(self, outs):
        self.outs = list(outs)
        self.i = 0

    def __call__(self, *a, **k):
        o = self.outs[self.i % len(self.outs)]
        self.i += 1
        return o


# ------------------------------------------------------------------------------------------------
def gen_mvsnet():
    import torch
    import cases
    sys.path.insert(0, os.path.join(REF, "MVSNet"))
    from models.module import homo_warping, depth_regression
    from models.mvsnet import MVSNet, CostRegNet

    torch.set_grad_enabled(False)
    # (1) homo_warping, fixed planes  -- MVSNet/models/module.py:46-87
    c = cases.warp_plane_case()
    out = homo_warping(t(c["src_fea"]), t(c["src_proj"]), t(c["ref_proj"]), t(c["depth"]))
    save("mvsnet_warp_plane", out=out.numpy(), proj=proj_product(c["src_proj"], c["ref_proj"]))

    # (2) CostRegNet alone -- mvsnet.py:48-93
    sd = cases.costreg_state("mvsnet", seed=11)
    net = CostRegNet().eval()
    load_sd(net, sd)
    x = np.random.RandomState(21).standard_normal((1, 32, 8, 16, 24)).astype(np.float32)
    save("mvsnet_costreg", out=net(t(x)).numpy())

    # (3) full hot path through MVSNet.forward (eval) with the FeatureNet stubbed out
    v = cases.volume_case(n_views=4, C=32, H=16, W=24, D=8, seed=3)
    model = MVSNet(refine=False).eval()
    load_sd(model.cost_regularization, sd)
    model._modules.pop("feature")
    model.feature = Replay([t(f) for f in v["feats"]])
    cap = {}
    model.cost_regularization.register_forward_pre_hook(lambda m, inp: cap.__setitem__("var", inp[0].clone()))
    model.cost_regularization.register_forward_hook(lambda m, inp, o: cap.__setitem__("logits", o.clone()))
    imgs = torch.zeros(1, 4, 3, 64, 96)
    res = model(imgs, t(v["proj"]), t(v["depth"]))
    projs = np.stack([proj_product(v["proj"][:, i], v["proj"][:, 0]) for i in range(1, 4)], 1)
    save("mvsnet_forward", var=cap["var"].numpy(), logits=cap["logits"].numpy(), depth=res["depth"].numpy(),
         conf=res["photometric_confidence"].numpy(), proj=projs)

    # (4) softmax + regression + confidence on raw logits (mvsnet.py:183-191), no clamp
    lc = cases.logits_case()
    import torch.nn.functional as F
    p = F.softmax(t(lc["logits"]), dim=1)
    depth = depth_regression(p, t(lc["depth"]))
    D = p.shape[1]
    sum4 = 4 * F.avg_pool3d(F.pad(p.unsqueeze(1), pad=(0, 0, 0, 0, 1, 2)), (4, 1, 1), stride=1, padding=0).squeeze(1)
    eidx = depth_regression(p, torch.arange(D, dtype=torch.float))
    idx = eidx.long()
    conf = torch.gather(sum4, 1, idx.unsqueeze(1)).squeeze(1)
    save("mvsnet_regress", prob=p.numpy(), depth=depth.numpy(), expect_idx=eidx.numpy(), index=idx.numpy(),
         conf=conf.numpy())


def gen_cas():
    import torch
    import torch.nn.functional as F
    import cases
    sys.path.insert(0, os.path.join(REF, "CasMVSNet"))
    from models.module import homo_warping, CostRegNet, get_depth_range_samples
    from models.cas_mvsnet import DepthNet, CascadeMVSNet

    torch.set_grad_enabled(False)
    # (1) homo_warping with per-pixel hypotheses -- CasMVSNet/models/module.py:245-280
    c = cases.warp_pixel_case()
    out = homo_warping(t(c["src_fea"]), t(c["src_proj"]), t(c["ref_proj"]), t(c["depth"]))
    save("cas_warp_pixel", out=out.numpy(), proj=proj_product(c["src_proj"], c["ref_proj"]))

    # (2) DepthNet.forward (one cascade stage) -- cas_mvsnet.py:12-66
    for tag, cin, per_pixel in (("s2", 16, True), ("s1", 32, False)):
        cc = cases.cas_case(C=cin, per_pixel=per_pixel, seed=4 if per_pixel else 8)
        sd = cases.costreg_state("cas", cin=cin, base=8, seed=12)
        reg = CostRegNet(cin, 8).eval()
        load_sd(reg, sd)
        cap = {}
        reg.register_forward_pre_hook(lambda m, inp: cap.__setitem__("var", inp[0].clone()))
        reg.register_forward_hook(lambda m, inp, o: cap.__setitem__("logits", o.clone()))
        dn = DepthNet().eval()
        depth = t(cc["depth"])
        if not per_pixel:   # stage 1 hands [B,D,H,W] too (cas_mvsnet.py:150); planes are uniform
            depth = depth[:, :, None, None].repeat(1, 1, 16, 24)
        res = dn([t(f) for f in cc["feats"]], t(cc["proj"]), depth, depth.shape[1], reg)
        pm = t(cc["proj"])
        comp = pm[:, :, 0].clone()
        comp[:, :, :3, :4] = torch.matmul(pm[:, :, 1, :3, :3], pm[:, :, 0, :3, :4])
        projs = torch.stack([torch.matmul(comp[:, i], torch.inverse(comp[:, 0])) for i in range(1, comp.shape[1])], 1)
        save(f"cas_depthnet_{tag}", var=cap["var"].numpy(), logits=cap["logits"].numpy(), depth=res["depth"].numpy(),
             conf=res["photometric_confidence"].numpy(), proj=projs.numpy())

    # (3) CostRegNet alone, all three stage widths
    for cin in (32, 16, 8):
        sd = cases.costreg_state("cas", cin=cin, base=8, seed=13)
        reg = CostRegNet(cin, 8).eval()
        load_sd(reg, sd)
        x = np.random.RandomState(22 + cin).standard_normal((1, cin, 8, 16, 24)).astype(np.float32)
        save(f"cas_costreg_c{cin}", out=reg(t(x)).numpy())

    # (4) full 3-stage cascade -- cas_mvsnet.py:109-165 -- FeatureNet stubbed
    k = cases.cascade_case()
    model = CascadeMVSNet(ndepths=k["ndepths"]).eval()
    for i, cin in enumerate((32, 16, 8)):
        load_sd(model.cost_regularization[i], cases.costreg_state("cas", cin=cin, base=8, seed=14 + i))
    n_views = k["feats"]["stage1"].shape[0]
    outs = [{s: t(k["feats"][s][v]) for s in k["feats"]} for v in range(n_views)]
    model._modules.pop("feature")
    model.feature = Replay(outs)
    imgs = torch.zeros(1, n_views, 3, k["H"], k["W"])
    res = model(imgs, {s: t(p) for s, p in k["projs"].items()}, t(k["depth_values"]))
    arrays = {}
    for s in ("stage1", "stage2", "stage3"):
        arrays[s + "_depth"] = res[s]["depth"].numpy()
        arrays[s + "_conf"] = res[s]["photometric_confidence"].numpy()
    save("cas_cascade", **arrays)

    # (5) hypothesis generation alone -- module.py:485-524
    cur = t(np.random.RandomState(31).uniform(500, 800, (2, 12, 20)).astype(np.float32))
    smp = get_depth_range_samples(cur_depth=cur, ndepth=8, depth_inteval_pixel=5.3, dtype=cur.dtype,
                                  device=cur.device, shape=[2, 12, 20], max_depth=935.0, min_depth=425.0)
    save("cas_range_samples", out=smp.numpy())


def gen_cvp():
    import torch
    import pdb
    import cases
    pdb.set_trace = lambda *a, **k: None
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.empty_cache = lambda: None
    sys.path.insert(0, os.path.join(REF, "CVP-MVSNet"))
    from models.modules import homo_warping, proj_cost
    from models.net import CostRegNet, network

    np.seterr(all="warn")
    torch.set_grad_enabled(False)
    cc = cases.cvp_case(per_pixel=False)
    # (1) homo_warping(K, E separately) -- CVP-MVSNet/models/modules.py:81-119
    out = homo_warping(t(cc["feats"][1]), t(cc["ref_in"]), t(cc["src_in"][:, 0]), t(cc["ref_ex"]),
                       t(cc["src_ex"][:, 0]), t(cc["depth"]))
    save("cvp_warp", out=out.numpy())

    # (2) proj_cost (per-pixel hypotheses, aliasing quirk) -- modules.py:221-275
    cp = cases.cvp_case(per_pixel=True)
    settings = types.SimpleNamespace(nsrc=2, mode="train")
    srcs = [[t(cp["feats"][1 + i])] for i in range(2)]
    vol = proj_cost(settings, t(cp["feats"][0]).clone(), srcs, 0, t(cp["ref_in"]), t(cp["src_in"]),
                    t(cp["ref_ex"]), t(cp["src_ex"]), t(cp["depth"]))
    save("cvp_proj_cost", out=vol.numpy())

    # (3) CostRegNet -- net.py:52-89
    sd = cases.costreg_state("cvp", seed=15)
    reg = CostRegNet().eval()
    load_sd(reg, sd)
    x = np.random.RandomState(23).standard_normal((1, 16, 8, 16, 24)).astype(np.float32)
    save("cvp_costreg", out=reg(t(x)).numpy())

    # (4) whole network.forward, nscale=2, test mode, FeaturePyramid stubbed -- net.py:99-207
    args = types.SimpleNamespace(nsrc=2, nscale=2, mode="test")
    net = network(args).eval()
    load_sd(net.cost_reg_refine, sd)
    H, W = 32, 48
    from mvs_b200 import synth
    fine = synth.features(3, 16, H, W, 41, 1)
    coarse = synth.features(3, 16, H // 2, W // 2, 42, 1)
    net._modules.pop("featurePyramid")
    net.featurePyramid = Replay([[t(fine[v]), t(coarse[v])] for v in range(3)])
    ref_in, src_in, ref_ex, src_ex = synth.cvp_cameras(2, W, 43, 1)
    dmin = torch.tensor([425.0], dtype=torch.float64)
    dmax = torch.tensor([935.0], dtype=torch.float64)
    res = net(torch.zeros(1, 3, H, W), torch.zeros(1, 2, 3, H, W), t(ref_in), t(src_in), t(ref_ex), t(src_ex), dmin, dmax)
    save("cvp_network", depth0=res["depth_est_list"][0].numpy(), depth1=res["depth_est_list"][1].numpy(),
         conf=res["prob_confidence"].numpy())


def gen_pl():
    import torch
    import cases
    kornia = types.ModuleType("kornia")
    kutils = types.ModuleType("kornia.utils")

    def create_meshgrid(height, width, normalized_coordinates=True, device=None, dtype=torch.float32):
        # kornia.utils.create_meshgrid semantics: [1,H,W,2], last dim (x,y), pixel units when not normalised
        xs = torch.linspace(0, width - 1, width, dtype=dtype)
        ys = torch.linspace(0, height - 1, height, dtype=dtype)
        assert not normalized_coordinates
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        return torch.stack((gx, gy), -1)[None]

    kutils.create_meshgrid = create_meshgrid
    kornia.utils = kutils
    sys.modules["kornia"] = kornia
    sys.modules["kornia.utils"] = kutils
    iabn = types.ModuleType("inplace_abn")
    iabn.InPlaceABN = torch.nn.BatchNorm3d
    sys.modules["inplace_abn"] = iabn
    sys.path.insert(0, os.path.join(REF, "MVSNet_pl"))
    from models.modules import homo_warp

    torch.set_grad_enabled(False)
    c = cases.warp_plane_case(seed=9)
    ref_inv = torch.inverse(t(c["ref_proj"]))
    out = homo_warp(t(c["src_fea"]), t(c["src_proj"]), ref_inv, t(c["depth"]))
    save("pl_homo_warp", out=out.numpy(), proj=(t(c["src_proj"]) @ ref_inv).numpy())


def gen_casfeat():
    """The caller side: the reference's own FeatureNet (CasMVSNet/models/module.py:304-405) and the
    WHOLE CascadeMVSNet.forward from images (cas_mvsnet.py:109-165), nothing stubbed."""
    import torch
    import cases
    sys.path.insert(0, os.path.join(REF, "CasMVSNet"))
    from models.module import FeatureNet
    from models.cas_mvsnet import CascadeMVSNet

    torch.set_grad_enabled(False)
    fn = FeatureNet(base_channels=8, stride=4, num_stage=3, arch_mode="fpn").eval()
    load_sd(fn, cases.featurenet_state(33))
    img = t(cases.synth.images_u8(2, 48, 80, seed=21)[0]).float() / 255.0          # general_eval.py:81-86
    out = fn(img)
    save("cas_featurenet", **{k: v.numpy() for k, v in out.items()})

    k = cases.full_model_case()
    model = CascadeMVSNet(ndepths=k["ndepths"]).eval()
    load_sd(model, cases.full_model_state())
    res = model(t(k["imgs_u8"]).float() / 255.0, {s: t(p) for s, p in k["projs"].items()}, t(k["depth_values"]))
    arrays = {}
    for s in ("stage1", "stage2", "stage3"):
        arrays[s + "_depth"] = res[s]["depth"].numpy()
        arrays[s + "_conf"] = res[s]["photometric_confidence"].numpy()
    save("cas_full_model", **arrays)


def gen_mvsnetfeat():
    """MVSNet's own FeatureNet + the WHOLE MVSNet.forward from images (MVSNet/models/mvsnet.py:8-45,136-194), nothing stubbed."""
    import torch
    import cases
    sys.path.insert(0, os.path.join(REF, "MVSNet"))
    from models.mvsnet import MVSNet
    torch.set_grad_enabled(False)
    k = cases.mvsnet_model_case()
    model = MVSNet(refine=False).eval()
    load_sd(model, cases.mvsnet_model_state())
    imgs = t(k["imgs_u8"]).float() / 255.0
    feat = model.feature(imgs[:, 1])
    res = model(imgs, t(k["proj"]), t(k["depth"]))
    save("mvsnet_full_model", feature_view1=feat.numpy(), depth=res["depth"].numpy(), conf=res["photometric_confidence"].numpy())


GEN = {"mvsnet": gen_mvsnet, "cas": gen_cas, "cvp": gen_cvp, "pl": gen_pl, "casfeat": gen_casfeat, "mvsnetfeat": gen_mvsnetfeat}

if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("make_golden.py needs /root/reference (build container only)")
    if len(sys.argv) > 1:
        import torch
        torch.manual_seed(0)
        torch.set_num_threads(8)
        print(f"[{sys.argv[1]}]")
        GEN[sys.argv[1]]()
    else:
        for name in GEN:
            subprocess.check_call([sys.executable, os.path.abspath(__file__), name])
