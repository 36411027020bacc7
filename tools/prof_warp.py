#!/usr/bin/env python
"""Times the fused warp+variance kernels alone at BASELINE.json shapes (CUDA events, L2 flushed
between repetitions) and prints achieved algorithmic GB/s vs the measured HBM peak.
Also the command ncu wraps for the --set full capture of this kernel (profiles/).

    python tools/prof_warp.py [--cfg cfg3] [--reps 10] [--mode c8|strict|both]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from mvs_b200 import synth, ops


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="cfg3")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--mode", default="both")
    ap.add_argument("--stages", default="")
    a = ap.parse_args()
    cfg = synth.CONFIGS[a.cfg]
    dev = "cuda:0"
    peak = 6545.6
    pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pp):
        peak = float(json.load(open(pp))["hbm_gbs"])
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    n, B = cfg["n_views"], cfg["batch"]
    results = []
    for si, (c, d, h, w) in enumerate(cfg["stages"]):
        if a.stages and str(si + 1) not in a.stages.split(","):
            continue
        per_pixel = cfg["family"] == "cas" and si > 0
        proj = torch.from_numpy(synth.proj_matrices(n, w, 0, B))
        prod = torch.stack([proj[:, i] @ torch.inverse(proj[:, 0]) for i in range(1, n)], 1)
        rots = [prod[:, i, :3, :3].reshape(B, 9).contiguous().to(dev) for i in range(n - 1)]
        trs = [prod[:, i, :3, 3].contiguous().to(dev) for i in range(n - 1)]
        if per_pixel:
            depth = torch.from_numpy(synth.depth_per_pixel(d, h, w, 2.65 * (2 if si == 1 else 1), B)).to(dev)
        else:
            depth = torch.from_numpy(synth.depth_planes(d, B)).to(dev)
        g = torch.Generator(device=dev).manual_seed(si)
        feats = [torch.randn(B, c, h, w, device=dev, generator=g) for _ in range(n)]
        # in-bounds statistics of this rig (changes gather traffic)
        _, _, mask, _ = ops.warp_taps(rots[0], trs[0], depth, h, w, want_ixy=False)
        inb = float((mask == 15).float().mean())
        del mask
        modes = ["c8h", "c8", "c8b", "strict"] if a.mode == "both" else a.mode.split(",")
        for mode in modes:
            if mode in ("c8", "c8b", "c8h", "c8g"):     # c8h: fp16 maps, TMA-staged kernel; c8g: fp16 maps, L1-gather kernel
                packed = [ops.pack_c8(f, torch.float16 if mode in ("c8h", "c8g") else torch.bfloat16) for f in feats]
                fl = 64 if mode == "c8b" else (512 if mode == "c8g" else (1024 if mode == "c8h" else 0))        # MVS_BLEND_BF16 / MVS_WARP_NO_TMA
                fn = lambda: ops.cost_volume_c8(packed[0], packed[1:], rots, trs, depth, fl)
                s = 2
            else:
                fn = lambda: ops.cost_volume(feats[0], feats[1:], rots, trs, depth)
                s = 4
            nbytes = synth.warp_variance_bytes(n, B, c, d, h, w, s, s, per_pixel)
            for _ in range(3):
                out = fn()
            del out
            times = []
            for _ in range(a.reps):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); out = fn(); e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
                del out
            ms = float(np.median(times))
            gbs = nbytes / ms / 1e6
            r = dict(cfg=a.cfg, stage=si + 1, mode=mode, C=c, D=d, H=h, W=w, B=B, nsrc=n - 1, per_pixel=per_pixel,
                     fully_inbounds_frac=round(inb, 3), alg_MB=round(nbytes / 1e6, 1), ms_median=round(ms, 4),
                     ms_min=round(min(times), 4), GBps=round(gbs, 1), frac_of_measured_peak=round(gbs / peak, 3))
            results.append(r)
            print(json.dumps(r), flush=True)
    return results


if __name__ == "__main__":
    main()
