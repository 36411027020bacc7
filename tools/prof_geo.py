#!/usr/bin/env python
"""Times the fused geometric-consistency filter (mvs_geo_fuse) at the cfg3 image size against its HBM floor and counts
(the side-by-side time of the NumPy restatement is printed by tests/test_geo_filter.py::test_gpu_full_size_vs_numpy_time:
only tests/ may import oracle/).

    python tools/prof_geo.py [--nsrc 10]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import cases
from mvs_b200 import fusion


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nsrc", type=int, default=10)
    a = ap.parse_args()
    H, W = 1184, 1600
    g = cases.geo_case(n_src=a.nsrc, H=H, W=W)
    d = [torch.from_numpy(x).cuda() for x in g["depth"]]
    conf = torch.from_numpy(g["conf"]).cuda()
    fn = lambda: fusion.fuse_ref_view(d[0], conf, g["K"][0], g["E"][0], d[1:], list(g["K"][1:]), list(g["E"][1:]))
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    # algorithmic bytes: every map read once + geo_sum (4) + depth_avg (8) + final_mask (1) written once
    nbytes = H * W * (4 * (a.nsrc + 2) + 13)
    print(json.dumps(dict(H=H, W=W, nsrc=a.nsrc, ms=round(ms, 4), alg_MB=round(nbytes / 1e6, 1), GBps=round(nbytes / ms / 1e6, 1),
                          final_mask_frac=float(out["final_mask"].float().mean()))))


if __name__ == "__main__":
    main()
