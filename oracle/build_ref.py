#!/usr/bin/env python
"""Build the UNMODIFIED reference extensions into oracle/_ref/ (test infrastructure only).

This is the recipe the task calls "oracle/_ref": the reference's own three torch
extensions (`pointnet2_cuda`, `roipool3d_cuda`, `iou3d_cuda`; reference `setup.py:53-83`)
compiled from the sources WHERE THEY LIE under /root/reference, with outputs only into
`oracle/_ref/` (git-ignored, but shipped to the GPU box by gpurun).  We do not run the
reference's own build system (setup.py); we invoke nvcc / g++ directly with the same
flags `torch.utils.cpp_extension.CUDAExtension` would use for TORCH_CUDA_ARCH_LIST=10.0.

No reference source is copied into this repository.

Only tests/, bench.py's cpu_baseline / --impl reference legs and __graft_entry__.smoke()
may load what this script produces; the product path (jmodt_b200/) never does.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("JMODT_REFERENCE", "/root/reference")

EXTS = {
    "iou3d_cuda": ("jmodt/ops/iou3d/src", ["iou3d.cpp", "iou3d_kernel.cu"]),
    "pointnet2_cuda": (
        "jmodt/ops/pointnet2/src",
        [
            "pointnet2_api.cpp", "ball_query.cpp", "ball_query_gpu.cu", "group_points.cpp",
            "group_points_gpu.cu", "interpolate.cpp", "interpolate_gpu.cu", "sampling.cpp",
            "sampling_gpu.cu",
        ],
    ),
    "roipool3d_cuda": ("jmodt/ops/roipool3d/src", ["roipool3d.cpp", "roipool3d_kernel.cu"]),
}


def _flags():
    import torch
    from torch.utils import cpp_extension as ce

    inc = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}"]
    abi = f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"
    common = ["-DTORCH_API_INCLUDE_EXTENSION_H", abi, "-std=c++17", "-O3"]
    libdirs = ce.library_paths("cuda")
    return inc, common, libdirs


def available() -> bool:
    return all(os.path.exists(os.path.join(OUT, n + ".so")) for n in EXTS)


def build(force: bool = False, jobs: int = 8) -> bool:
    """Returns True if oracle/_ref holds all three extensions after the call.  Also refreshes the mirror of the
    reference's python files under oracle/_ref/py (oracle/refpy.py) that the GPU-box golden generator and the drop-in
    tests import."""
    try:
        from oracle import refpy
    except ImportError:
        import refpy
    refpy.stage(force=force)
    if available() and not force:
        return True
    if not os.path.isdir(REF):
        return available()
    os.makedirs(OUT, exist_ok=True)
    objdir = os.path.join(OUT, "obj")
    os.makedirs(objdir, exist_ok=True)
    inc, common, libdirs = _flags()
    cmds = []
    for name, (sub, srcs) in EXTS.items():
        for s in srcs:
            src = os.path.join(REF, sub, s)
            obj = os.path.join(objdir, f"{name}__{s}.o")
            define = f"-DTORCH_EXTENSION_NAME={name}"
            if s.endswith(".cu"):
                cmd = ["nvcc", "-c", src, "-o", obj, define, *common, *inc,
                       "-gencode", "arch=compute_100,code=sm_100",
                       "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
                       "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
                       "-D__CUDA_NO_BFLOAT16_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__"]
            else:
                cmd = ["g++", "-c", src, "-o", obj, define, *common, *inc, "-fPIC", "-w"]
            cmds.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("reference build failed: " + " ".join(cmd) + "\n" + r.stderr[-4000:])

    with ThreadPoolExecutor(jobs) as ex:
        list(ex.map(run, cmds))
    for name, (sub, srcs) in EXTS.items():
        objs = [os.path.join(objdir, f"{name}__{s}.o") for s in srcs]
        so = os.path.join(OUT, name + ".so")
        cmd = ["g++", "-shared", *objs, "-o", so, *[f"-L{d}" for d in libdirs],
               *[f"-Wl,-rpath,{d}" for d in libdirs],
               "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
               "-lcudart"]
        run(cmd)
    return available()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "built" if ok else "UNAVAILABLE (no /root/reference and no prebuilt files)")
    sys.exit(0 if ok else 1)
