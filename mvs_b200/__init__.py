"""mvs_b200 -- B200-native (sm_100a) implementation of the MVSNet-family cost-volume hot path:
homography warp -> variance cost volume -> 3D-UNet CostRegNet -> softmax / depth regression,
behind the reference's own Python call signatures (doubleZ0108/MVS).

Importing the package does not need a GPU; calling any operator does (there is no CPU fallback).
"""
from . import synth  # noqa: F401  (NumPy only)

__version__ = "0.1.0"

_LAZY = {
    "ops": "ops", "modules": "modules", "cascade": "cascade", "patch": "patch", "dist": "dist",
    "homo_warping": "ops", "homo_warping_cvp": "ops", "homo_warp": "ops", "depth_regression": "ops",
    "depth_regression_refine": "ops", "cost_volume": "ops", "cost_volume_c8": "ops", "pack_c8": "ops",
    "unpack_c8": "ops", "conv3d": "ops", "softargmin_conf": "ops", "warp_taps": "ops", "relative_pose": "ops",
    "depth_range_samples": "ops",
    "CostRegNet": "modules", "CostRegNetMVSNet": "modules", "CostRegNetCas": "modules", "CostRegNetCVP": "modules",
    "DepthNet": "modules", "build_cost_volume": "modules", "proj_cost": "modules", "mvsnet_hot_path": "modules",
    "ConvBnReLU3D": "modules", "Conv3d": "modules", "Deconv3d": "modules",
    "cascade_hot_path": "cascade", "patch_reference": "patch", "graph": "graph", "GraphedStep": "graph",
    "train": "train", "pyramid": "pyramid", "cvp_hot_path": "pyramid", "FeaturePyramid": "pyramid",
    "featurenet": "featurenet", "FeatureNet": "featurenet", "CascadeMVSNet": "featurenet", "io": "io", "read_pfm": "io",
    "mvsnet": "mvsnet", "MVSNet": "mvsnet",
    "save_pfm": "io", "write_ply": "io", "filter_depth": "fusion",
    "fusion": "fusion", "reproject_with_depth": "fusion", "check_geometric_consistency": "fusion", "fuse_ref_view": "fusion", "backproject": "fusion",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        mod = importlib.import_module("." + _LAZY[name], __name__)
        return mod if name == _LAZY[name] else getattr(mod, name)
    raise AttributeError(name)
