"""File formats on either side of the hot path (row f4 "writers").

Byte-compatible with what the reference's callers read and write, so eval.py / test.py outputs stay interchangeable:
  * PFM  (`read_pfm`, `save_pfm`): MVSNet/datasets/data_io.py:5-70 == CasMVSNet/datasets/data_io.py -- header
    ``Pf\\n W H\\n -1.000000\\n`` ("PF" for 3 channels), little-endian float32 rows stored BOTTOM-UP;
  * PLY  (`write_ply`): what ``PlyData([PlyElement.describe(vertex_all, 'vertex')]).write(plyfilename)`` emits in
    MVSNet/eval.py:318-326 (third-party `plyfile`, absent from this image: binary_little_endian 1.0, one `vertex` element
    with float x, y, z + uchar red, green, blue, 15 bytes per vertex, no comments).
Plain host code: these are file formats, not kernels.
"""
from __future__ import annotations

import re
import sys

import numpy as np


def read_pfm(filename):
    """-> (data float32 [H,W] or [H,W,3], top row first; scale) -- data_io.py:5-41."""
    with open(filename, "rb") as f:
        header = f.readline().decode("utf-8").rstrip()
        if header == "PF":
            color = True
        elif header == "Pf":
            color = False
        else:
            raise Exception("Not a PFM file.")
        m = re.match(r"^(\d+)\s(\d+)\s$", f.readline().decode("utf-8"))
        if not m:
            raise Exception("Malformed PFM header.")
        width, height = map(int, m.groups())
        scale = float(f.readline().rstrip())
        endian = "<" if scale < 0 else ">"
        scale = abs(scale)
        data = np.fromfile(f, endian + "f")
    shape = (height, width, 3) if color else (height, width)
    return np.flipud(np.reshape(data, shape)), scale


def save_pfm(filename, image, scale=1):
    """data_io.py:44-70: float32 [H,W], [H,W,1] or [H,W,3]; rows are written bottom-up."""
    image = np.asarray(image)
    if image.dtype.name != "float32":
        raise Exception("Image dtype must be float32.")
    if len(image.shape) == 3 and image.shape[2] == 3:
        color = True
    elif len(image.shape) == 2 or (len(image.shape) == 3 and image.shape[2] == 1):
        color = False
    else:
        raise Exception("Image must have H x W x 3, H x W x 1 or H x W dimensions.")
    image = np.flipud(image)
    endian = image.dtype.byteorder
    if endian == "<" or (endian == "=" and sys.byteorder == "little"):
        scale = -scale
    with open(filename, "wb") as f:
        f.write(("PF\n" if color else "Pf\n").encode("utf-8"))
        f.write("{} {}\n".format(image.shape[1], image.shape[0]).encode("utf-8"))
        f.write(("%f\n" % scale).encode("utf-8"))
        np.ascontiguousarray(image).tofile(f)


PLY_VERTEX_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("red", "u1"), ("green", "u1"), ("blue", "u1")])


def write_ply(plyfilename, vertexs, vertex_colors):
    """vertexs [N,3] float, vertex_colors [N,3] uint8 -> the binary PLY of MVSNet/eval.py:311-326."""
    vertexs = np.asarray(vertexs, dtype=np.float32).reshape(-1, 3)
    vertex_colors = np.asarray(vertex_colors, dtype=np.uint8).reshape(-1, 3)
    if len(vertexs) != len(vertex_colors):
        raise ValueError("one colour per vertex")
    rec = np.empty(len(vertexs), PLY_VERTEX_DTYPE)
    rec["x"], rec["y"], rec["z"] = vertexs[:, 0], vertexs[:, 1], vertexs[:, 2]
    rec["red"], rec["green"], rec["blue"] = vertex_colors[:, 0], vertex_colors[:, 1], vertex_colors[:, 2]
    header = ("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
              "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n" % len(rec))
    with open(plyfilename, "wb") as f:
        f.write(header.encode("ascii"))
        rec.tofile(f)


def read_ply(plyfilename):
    """Inverse of write_ply (tests / round trips): -> (vertexs float32 [N,3], vertex_colors uint8 [N,3])."""
    with open(plyfilename, "rb") as f:
        n = None
        while True:
            line = f.readline().decode("ascii").strip()
            if line.startswith("element vertex"):
                n = int(line.split()[-1])
            if line == "end_header":
                break
        rec = np.fromfile(f, PLY_VERTEX_DTYPE, count=n)
    return np.stack([rec["x"], rec["y"], rec["z"]], 1), np.stack([rec["red"], rec["green"], rec["blue"]], 1)
