#!/usr/bin/env python
"""Per-role timers of the UMMA conv kernels (mvs_conv3d_c8_set_trace): where a CTA's time goes.

    python tools/prof_conv_trace.py [--stages 1,2,3] [--layers conv0,prob]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from mvs_b200 import synth, ops, _lib
from prof_conv import LAYERS

NAMES = ["cta_total", "prod_wait_empty", "prod_stage", "iss_wait_full", "iss_wait_tempty", "iss_issue", "epi_wait_tfull",
         "epi_work", "-", "prologue"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stages", default="1,2,3")
    ap.add_argument("--layers", default="conv0,conv2,conv4,prob")
    ap.add_argument("--ctas", type=int, default=4096)
    ap.add_argument("--dw", type=int, default=1, help="W-de-interleaved layout flags as CostRegNet's fast path sets them")
    a = ap.parse_args()
    dev = "cuda:0"
    cfg = synth.CONFIGS["cfg3"]
    lib = _lib.lib()
    for si in [int(s) - 1 for s in a.stages.split(",")]:
        c, d, h, w = cfg["stages"][si]
        for name, cin_m, cout_m, stride, tr, lvl, has_skip in LAYERS:
            if name not in a.layers.split(","):
                continue
            cin = c if cin_m is None else 8 * cin_m
            cout = 1 if cout_m is None else 8 * cout_m
            D, H, W = d >> lvl, h >> lvl, w >> lvl
            x = torch.randn(1, (cin + 7) // 8, D, H, W, 8, device=dev).bfloat16()
            wt = torch.randn((cin, cout, 3, 3, 3) if tr else (cout, cin, 3, 3, 3), device=dev) / (27 * cin) ** 0.5
            pk = ops.pack_conv_weights(wt, stride, tr)
            y0 = ops.conv3d_c8(x, pk, cin, cout, None, None, None, stride, tr, cout != 1)
            sk = torch.zeros_like(y0) if has_skip else None
            lay = 0
            if a.dw:
                lay = (_lib.X_DW if (stride == 2 and not tr) else 0) | (_lib.SKIP_DW if has_skip else 0) | \
                      (_lib.Y_DW if name in ("conv0", "conv2", "conv4") else 0)
            fn = lambda: ops.conv3d_c8(x, pk, cin, cout, None, None, sk, stride, tr, cout != 1, layout=lay)
            fn(); fn()
            buf = torch.zeros(a.ctas * 16, dtype=torch.int64, device=dev)
            lib.mvs_conv3d_c8_set_trace(buf.data_ptr(), a.ctas)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            lib.mvs_conv3d_c8_set_trace(None, 0)
            torch.cuda.synchronize()
            t = buf.view(a.ctas, 16).cpu().double()
            t = t[t[:, 0] > 0]
            steps = t[:, 10].mean().item()
            print(f"stage {si + 1} {name} cin={cin} cout={cout} {D}x{H}x{W}: {e0.elapsed_time(e1):.3f} ms, {len(t)} CTAs traced, "
                  f"{steps:.1f} steps/CTA")
            if len(t) == 0:
                continue
            for k, nm in enumerate(NAMES):
                if nm == "-":
                    continue
                print(f"    {nm:18s} {t[:, k].mean().item():10.0f} clk/CTA   {t[:, k].mean().item() / steps:8.0f} clk/step")


if __name__ == "__main__":
    main()
