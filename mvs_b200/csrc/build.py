"""Builds mvs_b200/libmvs_b200.so from the .cu sources in this directory with nvcc for sm_100a.

    python -m mvs_b200.csrc.build [--force] [--verbose]

The library is built IN-TREE (it travels to the GPU box with the repo snapshot) and links cudart
statically, so loading it needs no CUDA driver: the symbol test runs in the GPU-less container.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libmvs_b200.so")
OBJ_DIR = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-DMVS_TARGET_SM=100",
          "--expt-relaxed-constexpr"]


def sources():
    return sorted(f for f in os.listdir(HERE) if f.endswith(".cu"))


def headers_mtime():
    hs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".cuh")]
    hs.append(os.path.join(os.path.dirname(PKG), "include", "mvs_b200.h"))
    return max(os.path.getmtime(h) for h in hs)


def compile_one(src, force, verbose):
    obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
    spath = os.path.join(HERE, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(spath), headers_mtime()):
        return obj, ""
    cmd = [NVCC, *ARCH, *CFLAGS, "-c", spath, "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{p.stdout}\n{p.stderr}")
    with open(obj + ".ptxas.log", "w") as f:
        f.write(p.stderr)
    return obj, p.stderr if verbose else ""


def build_variant(name, defines, only=("warp_c8.cu",)):
    """An experiment build of the same C-ABI (kernel ablations, tuning): recompiles `only` with extra -D flags, links it with
    the regular objects of everything else into mvs_b200/libmvs_b200_<name>.so (select with MVS_B200_LIB)."""
    build()
    vdir = os.path.join(OBJ_DIR, name)
    os.makedirs(vdir, exist_ok=True)
    objs = []
    for src in sources():
        if src in only:
            obj = os.path.join(vdir, src[:-3] + ".o")
            cmd = [NVCC, *ARCH, *CFLAGS, *[f"-D{d}" for d in defines], "-c", os.path.join(HERE, src), "-o", obj]
            p = subprocess.run(cmd, capture_output=True, text=True)
            if p.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{p.stdout}\n{p.stderr}")
        else:
            obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        objs.append(obj)
    out = os.path.join(PKG, f"libmvs_b200_{name}.so")
    p = subprocess.run([NVCC, *ARCH, "-shared", "-o", out, *objs, "-cudart", "static", "-Xcompiler", "-fPIC"], capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
    return out


def build(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: compile_one(s, force, verbose), srcs))
    objs = [r[0] for r in res]
    if verbose:
        for _, log in res:
            sys.stderr.write(log)
    if force or not os.path.exists(OUT) or any(os.path.getmtime(o) > os.path.getmtime(OUT) for o in objs):
        cmd = [NVCC, *ARCH, "-shared", "-o", OUT, *objs, "-cudart", "static", "-Xcompiler", "-fPIC"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
    return OUT


if __name__ == "__main__":
    if "--variant" in sys.argv:          # python -m mvs_b200.csrc.build --variant abl1 MVS_C8_ABLATE=1 [--only conv3d_umma.cu]
        argv = list(sys.argv)
        only = ("warp_c8.cu",)
        if "--only" in argv:
            j = argv.index("--only")
            only = tuple(argv[j + 1].split(","))
            del argv[j:j + 2]
        i = argv.index("--variant")
        print(build_variant(argv[i + 1], argv[i + 2:], only))
    else:
        print(build("--force" in sys.argv, "--verbose" in sys.argv))
