#!/usr/bin/env python
"""One warm-up + ONE whole-model step of the cfg3 workload (uint8 images -> FeatureNet mirror -> 3-stage hot path), eager
launches: the command `tools/launch_list.sh` wraps in ncu to list every kernel of a step with its device time."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import bench
from mvs_b200.featurenet import CascadeMVSNet

wl = bench.WORKLOADS[os.environ.get("MVS_CFG", "cfg3")]
hi = bench.host_inputs(wl)
dev = "cuda:0"
m = CascadeMVSNet(ndepths=wl["ndepths"], mode="fast")
m.load_state_dict({k: torch.from_numpy(np.asarray(a)) for k, a in bench.model_state(wl).items()}, strict=True)
m = m.to(dev).eval()
imgs = torch.from_numpy(hi["imgs"]).to(dev)
projs = {k: torch.from_numpy(a).to(dev) for k, a in hi["projs"].items()}
dv = torch.from_numpy(hi["depth_values"]).to(dev)
dmin, dmax = float(hi["depth_values"][0, 0]), float(hi["depth_values"][0, -1])
with torch.no_grad():
    for i in range(2):
        if i == 1:
            torch.cuda.synchronize()
            torch.cuda.nvtx.range_push("step")
        out = m(imgs, projs, dv, depth_min=dmin, depth_max=dmax)
    torch.cuda.synchronize()
print("done", float(out["depth"].mean()))
