#!/usr/bin/env python
"""Times every CostRegNet layer of BASELINE.json cfg3 on the tcgen05 path (CUDA events, L2 flushed)
and prints time, algorithmic HBM bytes -> GB/s, FLOPs -> TFLOP/s per layer and per stage.

    python tools/prof_conv.py [--stages 1,2,3] [--reps 5]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from mvs_b200 import synth, ops, _lib as L

LAYERS = [  # name, cin(x base), cout, stride, transposed, input level, skip
    ("conv0", None, 1, 1, False, 0, False), ("conv1", 1, 2, 2, False, 0, False), ("conv2", 2, 2, 1, False, 1, False),
    ("conv3", 2, 4, 2, False, 1, False), ("conv4", 4, 4, 1, False, 2, False), ("conv5", 4, 8, 2, False, 2, False),
    ("conv6", 8, 8, 1, False, 3, False), ("conv7", 8, 4, 2, True, 3, True), ("conv9", 4, 2, 2, True, 2, True),
    ("conv11", 2, 1, 2, True, 1, True), ("prob", 1, None, 1, False, 0, False)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stages", default="1,2,3")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--layers", default="")
    a = ap.parse_args()
    dev = "cuda:0"
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    cfg = synth.CONFIGS["cfg3"]
    for si in [int(s) - 1 for s in a.stages.split(",")]:
        c, d, h, w = cfg["stages"][si]
        tot_ms = tot_b = tot_f = 0.0
        for name, cin_m, cout_m, stride, tr, lvl, has_skip in LAYERS:
            if a.layers and name not in a.layers.split(","):
                continue
            cin = c if cin_m is None else 8 * cin_m
            cout = 1 if cout_m is None else 8 * cout_m
            D, H, W = d >> lvl, h >> lvl, w >> lvl
            x = torch.randn(1, (cin + 7) // 8, D, H, W, 8, device=dev).bfloat16()
            wt = torch.randn((cin, cout, 3, 3, 3) if tr else (cout, cin, 3, 3, 3), device=dev) / (27 * cin) ** 0.5
            pk = ops.pack_conv_weights(wt, stride, tr)
            scale = torch.ones(cout, device=dev); shift = torch.zeros(cout, device=dev)
            # layout flags as CostRegNet's fast path sets them (W-de-interleaved skip tensors)
            lay0 = (L.X_DW if (stride == 2 and not tr) else 0) | (L.Y_DW if name in ("conv0", "conv2", "conv4") else 0)
            fn = lambda skip=None: ops.conv3d_c8(x, pk, cin, cout, scale, shift, skip, stride, tr, cout != 1,
                                                 layout=lay0 | (L.SKIP_DW if skip is not None else 0))
            y = fn()
            skip = torch.zeros_like(y) if has_skip else None
            for _ in range(2):
                fn(skip)
            ts = []
            for _ in range(a.reps):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(skip); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            nbytes = x.numel() * 2 + y.numel() * y.element_size() + (y.numel() * 2 if has_skip else 0)
            vout = y.shape[2] * y.shape[3] * y.shape[4]
            flops = 2 * 27 * cin * cout * (D * H * W if tr and stride == 2 else vout)
            print(json.dumps(dict(stage=si + 1, layer=name, cin=cin, cout=cout, stride=stride, transposed=tr, D=D, H=H, W=W,
                                  ms=round(ms, 4), GBps=round(nbytes / ms / 1e6, 1), TFLOPs=round(flops / ms / 1e9, 2))), flush=True)
            tot_ms += ms; tot_b += nbytes; tot_f += flops
            del x, y, skip
        print(json.dumps(dict(stage=si + 1, layer="TOTAL", ms=round(tot_ms, 3), GBps=round(tot_b / tot_ms / 1e6, 1),
                              TFLOPs=round(tot_f / tot_ms / 1e9, 2))), flush=True)


if __name__ == "__main__":
    main()
