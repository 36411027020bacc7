"""PFM / PLY writers: byte-identical to the reference's own PFM writer (fixtures written by
MVSNet/datasets/data_io.py, tests/golden/make_golden_io.py); PLY against the published format (plyfile is not in this image)."""
import os

import numpy as np

import cases
from mvs_b200 import io as mio


def _arrays():
    rng = np.random.RandomState(77)
    return rng.uniform(400, 900, (7, 11)).astype(np.float32), rng.uniform(0, 1, (5, 6, 3)).astype(np.float32)


def test_save_pfm_bytes_equal_reference_writer(tmp_path):
    g, c = _arrays()
    for name, arr, scale in (("pfm_gray.pfm", g, 1), ("pfm_color.pfm", c, 2)):
        out = tmp_path / name
        mio.save_pfm(str(out), arr, scale)
        assert out.read_bytes() == open(os.path.join(cases.GOLDEN_DIR, name), "rb").read(), name
        back, s = mio.read_pfm(os.path.join(cases.GOLDEN_DIR, name))
        assert np.array_equal(back, arr) and s == scale
    # [H,W,1] is written as greyscale, other dtypes / ranks raise like the reference
    mio.save_pfm(str(tmp_path / "one.pfm"), g[..., None])
    assert (tmp_path / "one.pfm").read_bytes() == open(os.path.join(cases.GOLDEN_DIR, "pfm_gray.pfm"), "rb").read()
    for bad in (g.astype(np.float64), np.zeros((2, 3, 4), np.float32)):
        try:
            mio.save_pfm(str(tmp_path / "bad.pfm"), bad)
        except Exception:
            continue
        raise AssertionError("expected an exception")


def test_write_ply_layout_and_round_trip(tmp_path):
    rng = np.random.RandomState(3)
    v = rng.standard_normal((13, 3)).astype(np.float32) * 100
    c = rng.randint(0, 256, (13, 3)).astype(np.uint8)
    p = tmp_path / "m.ply"
    mio.write_ply(str(p), v, c)
    raw = p.read_bytes()
    head, body = raw.split(b"end_header\n", 1)
    assert head.decode().splitlines() == ["ply", "format binary_little_endian 1.0", "element vertex 13", "property float x",
                                          "property float y", "property float z", "property uchar red",
                                          "property uchar green", "property uchar blue"]
    assert len(body) == 13 * 15
    v2, c2 = mio.read_ply(str(p))
    assert np.array_equal(v2, v) and np.array_equal(c2, c)
    mio.write_ply(str(p), np.zeros((0, 3)), np.zeros((0, 3)))           # empty cloud
    assert mio.read_ply(str(p))[0].shape == (0, 3)
