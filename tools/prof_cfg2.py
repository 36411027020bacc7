#!/usr/bin/env python
"""Secondary data point: BASELINE.json configs[1] (MVSNet 640x512 image => features 32x128x160, N=5, D=192, batch 4) through
`mvsnet_hot_path` in fast mode (builder + CostRegNet + softargmin), CUDA events, eager and CUDA-graph replay.

    python tools/prof_cfg2.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import cases
from mvs_b200 import modules, synth
from mvs_b200.graph import GraphedStep


def main():
    dev = "cuda:0"
    B, N, D, H, W = 4, 5, 192, 128, 160
    sd = cases.costreg_state("mvsnet", seed=70)
    net = modules.CostRegNetMVSNet(mode="fast")
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    net = net.to(dev).eval()
    feats = [torch.from_numpy(f).to(dev).bfloat16() for f in synth.features(N, 32, H, W, 71, B)]
    proj = torch.from_numpy(synth.proj_matrices(N, W, 72, B)).to(dev)
    depth = torch.from_numpy(np.tile((425.0 + 2.65 * np.arange(D)).astype(np.float32), (B, 1))).to(dev)
    run = lambda: modules.mvsnet_hot_path(feats, proj, depth, net)

    def timed(fn, k=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(k):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / k

    with torch.no_grad():
        ms_eager = timed(run)
        g = GraphedStep(run)
        ms_graph = timed(g)
    print(json.dumps(dict(config="cfg2: MVSNet 640x512, N=5, D=192, batch 4, fast mode", ms_per_batch_eager=round(ms_eager, 3),
                          ms_per_batch_graph=round(ms_graph, 3), depth_maps_per_s=round(B / ms_graph * 1e3, 1))))


if __name__ == "__main__":
    main()
