#!/usr/bin/env python
"""tests/golden/geo_filter.npz: the reference's OWN reproject_with_depth / check_geometric_consistency
(MVSNet/eval.py:138-208, identical in CasMVSNet/test.py:237-294) executed with the real cv2 on tests/cases.geo_case().

eval.py cannot be imported (module-level `from plyfile import ...`, argparse, CUDA): the two function definitions are
extracted from the unmodified file with `ast` and executed in a namespace that holds only numpy and cv2.
Run in the build container only:   python tests/golden/make_golden_geo.py
"""
import ast
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE)))
import cases

SRC = "/root/reference/MVSNet/eval.py"
tree = ast.parse(open(SRC).read())
wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("reproject_with_depth", "check_geometric_consistency")]
assert len(wanted) == 2
ns = {"np": np, "cv2": cv2}
exec(compile(ast.Module(body=wanted, type_ignores=[]), SRC, "exec"), ns)

g = cases.geo_case()
out = {}
geo_sum = 0
reproj_all = []
for v in range(1, g["depth"].shape[0]):
    d_rep, x_rep, y_rep, x_src, y_src = ns["reproject_with_depth"](g["depth"][0], g["K"][0], g["E"][0], g["depth"][v], g["K"][v], g["E"][v])
    mask, d_masked, xs, ys = ns["check_geometric_consistency"](g["depth"][0], g["K"][0], g["E"][0], g["depth"][v], g["K"][v], g["E"][v])
    out[f"depth_reprojected_{v}"] = d_rep; out[f"x_reprojected_{v}"] = x_rep; out[f"y_reprojected_{v}"] = y_rep
    out[f"x_src_{v}"] = x_src; out[f"y_src_{v}"] = y_src; out[f"mask_{v}"] = mask; out[f"depth_masked_{v}"] = d_masked
    geo_sum = geo_sum + mask.astype(np.int32); reproj_all.append(d_masked)
# filter_depth loop body, MVSNet/eval.py:256-263 verbatim
depth_est_averaged = (sum(reproj_all) + g["depth"][0]) / (geo_sum + 1)
geo_mask = geo_sum >= 3
final_mask = np.logical_and(g["conf"] > 0.8, geo_mask)
out.update(geo_mask_sum=geo_sum, depth_est_averaged=depth_est_averaged, geo_mask=geo_mask, final_mask=final_mask)
np.savez_compressed(os.path.join(HERE, "geo_filter.npz"), **out)
print("wrote geo_filter.npz; mask fractions:", [float(out[f"mask_{v}"].mean()) for v in range(1, 5)], "final", float(final_mask.mean()))
