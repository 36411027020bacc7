import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from mvs_b200 import ops, train
dev = "cuda:0"
for cin, cout, D, H, W in [(16, 16, 48, 432, 576), (32, 32, 24, 216, 288), (64, 64, 24, 216, 288), (16, 1, 48, 432, 576)]:
    x = torch.randn(1, cin, D, H, W, device=dev)
    w = torch.randn(cout, cin, 3, 3, 3, device=dev) / (27 * cin) ** 0.5
    gy = torch.randn(1, cout, D, H, W, device=dev)
    fl = 2 * 27 * cin * cout * D * H * W
    for name, fn in (("fwd", lambda: ops.conv3d(x, w, None, None, None, 1, False, False)),
                     ("wgrad", lambda: train.conv3d_wgrad(x, gy, 1, False))):
        fn(); fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3): fn()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        print(f"{cin}->{cout} {D}x{H}x{W} {name}: {ms:.2f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
