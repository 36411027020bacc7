"""The C-ABI library loads without a GPU and exports exactly what include/mvs_b200.h declares."""
import os
import re

import cases
from mvs_b200 import _lib


def header_functions():
    src = open(os.path.join(cases.ROOT, "include", "mvs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(mvs_[a-z0-9_]+)\s*\(", src))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    declared = header_functions()
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mvs_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_metadata_calls_need_no_gpu():
    lib = _lib.lib()
    assert lib.mvs_version() >= 100
    assert lib.mvs_sm() == 100
    assert lib.mvs_last_error() is not None
    assert lib.mvs_launch_count() >= 0


def test_flags_match_header():
    src = open(os.path.join(cases.ROOT, "include", "mvs_b200.h")).read()
    defs = dict(re.findall(r"#define\s+MVS_([A-Z0-9_]+)\s+\(?(-?\d+)\)?", src))
    for py, c in [("ALIGN_CORNERS", "ALIGN_CORNERS"), ("PL_ORDER", "PL_ORDER"), ("REF_SUM_SQUARED", "REF_SUM_SQUARED"),
                  ("RELU", "RELU"), ("CLAMP_INDEX", "CLAMP_INDEX"), ("INPUT_IS_PROB", "INPUT_IS_PROB"),
                  ("DEPTH_PLANE", "DEPTH_PLANE"), ("DEPTH_PIXEL", "DEPTH_PIXEL"), ("F32", "F32"), ("BF16", "BF16"),
                  ("MAX_SRC", "MAX_SRC")]:
        assert getattr(_lib, py) == int(defs[c]), py


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from mvs_b200 import ops
    with pytest.raises(_lib.MvsError):
        ops.homo_warping(torch.zeros(1, 4, 8, 8), torch.eye(4)[None], torch.eye(4)[None], torch.ones(1, 2))
