"""patch_reference(): rebind the reference's hot-path names to the sm_100a implementations.

The reference has no plugin registry; its model files do ``from .module import *``
(MVSNet/models/mvsnet.py:4, CasMVSNet/models/cas_mvsnet.py:4, CVP-MVSNet/models/net.py:12), so the
names must be replaced in the CONSUMER module's globals as well as in the defining module.
train.py / eval.py / test.py then run unchanged.
"""
from __future__ import annotations

import sys
import types

from . import featurenet, modules, mvsnet, ops, pyramid

_COMMON = {
    "depth_regression": ops.depth_regression,
}

_BY_FAMILY = {
    # MVSNet/models/{module,mvsnet}.py
    "mvsnet": {"homo_warping": ops.homo_warping, "CostRegNet": modules.CostRegNetMVSNet,
               "ConvBnReLU3D": modules.ConvBnReLU3D, "FeatureNet": mvsnet.FeatureNet, "MVSNet": mvsnet.MVSNet},
    # CasMVSNet/models/{module,cas_mvsnet}.py
    "cas": {"homo_warping": ops.homo_warping, "CostRegNet": modules.CostRegNetCas, "DepthNet": modules.DepthNet,
            "Conv3d": modules.Conv3d, "Deconv3d": modules.Deconv3d, "FeatureNet": featurenet.FeatureNet,
            "CascadeMVSNet": featurenet.CascadeMVSNet},
    # CVP-MVSNet/models/{modules,net}.py
    "cvp": {"homo_warping": ops.homo_warping_cvp, "proj_cost": modules.proj_cost,
            "depth_regression_refine": ops.depth_regression_refine, "CostRegNet": modules.CostRegNetCVP,
            "FeaturePyramid": pyramid.FeaturePyramid},
    # MVSNet_pl/models/{modules,mvsnet}.py
    "pl": {"homo_warp": ops.homo_warp},
}


def patch_fusion(*mods) -> dict:
    """Rebind `reproject_with_depth` / `check_geometric_consistency` (MVSNet/eval.py:138-208, CasMVSNet/test.py:237-294)
    in the given modules -- or, with no arguments, in every loaded module that defines both (eval.py / test.py run as
    scripts define them in `__main__`).  The replacements take and return NumPy arrays like the originals."""
    from . import fusion
    if not mods:
        mods = tuple(m for m in list(sys.modules.values())
                     if m is not None and callable(getattr(m, "reproject_with_depth", None)) and
                     callable(getattr(m, "check_geometric_consistency", None)) and m is not fusion
                     and not getattr(m, "__name__", "").startswith("mvs_b200"))
    done = {}
    for m in mods:
        m.reproject_with_depth = fusion.reproject_with_depth
        m.check_geometric_consistency = fusion.check_geometric_consistency
        done[m.__name__] = ["reproject_with_depth", "check_geometric_consistency"]
    return done


def _family_of(mod: types.ModuleType) -> str:
    names = set(vars(mod))
    if "proj_cost" in names or "depth_regression_refine" in names or "network" in names:
        return "cvp"
    if "homo_warp" in names:
        return "pl"
    if "DepthNet" in names or "CascadeMVSNet" in names or "get_depth_range_samples" in names:
        return "cas"
    return "mvsnet"


def patch_reference(*mods, family: str | None = None) -> dict:
    """Rebind hot-path names inside already-imported reference modules.

    ``patch_reference()`` with no arguments patches every loaded module named ``models.*``.
    Returns {module_name: [patched names]}.
    """
    if not mods:
        mods = tuple(m for n, m in list(sys.modules.items()) if m is not None and (n == "models" or n.startswith("models.")))
    done = {}
    for m in mods:
        fam = family or _family_of(m)
        table = dict(_COMMON)
        table.update(_BY_FAMILY[fam])
        hit = []
        for name, repl in table.items():
            if name in vars(m):
                setattr(m, name, repl)
                hit.append(name)
        done[m.__name__] = hit
    return done
