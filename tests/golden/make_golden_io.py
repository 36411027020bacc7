#!/usr/bin/env python
"""PFM fixtures written by the UNMODIFIED reference writer (MVSNet/datasets/data_io.py:44-70), build container only:
    python tests/golden/make_golden_io.py
The PLY writer has no fixture: the reference writes it through the third-party `plyfile` package, which is absent from this
image (parity of mvs_b200.io.write_ply is therefore pinned to the published PLY format only -- see its docstring)."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_data_io", "/root/reference/MVSNet/datasets/data_io.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


def arrays():
    rng = np.random.RandomState(77)
    return rng.uniform(400, 900, (7, 11)).astype(np.float32), rng.uniform(0, 1, (5, 6, 3)).astype(np.float32)


if __name__ == "__main__":
    g, c = arrays()
    ref.save_pfm(os.path.join(HERE, "pfm_gray.pfm"), g)
    ref.save_pfm(os.path.join(HERE, "pfm_color.pfm"), c, scale=2)
    for n in ("pfm_gray.pfm", "pfm_color.pfm"):
        d, s = ref.read_pfm(os.path.join(HERE, n))
        print(n, os.path.getsize(os.path.join(HERE, n)), d.shape, s)
