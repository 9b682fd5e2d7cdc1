"""In-tree build of librestir_b200.so (sm_100a only) with nvcc.

    python -m restir_b200.build [--force] [--verbose]

The library is compiled with -fmad=false (device) and -ffp-contract=off (host) because bit-level parity with
the reference's arithmetic order is part of the contract; -lineinfo keeps ncu's source
page usable.  The .so lands next to this file so that it travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# RSTR_LIBNAME / RSTR_DEFINES: side-by-side experimental builds for A/B timing (scripts/gpu_ab.py); the default is the product
LIB = os.path.join(HERE, os.environ.get("RSTR_LIBNAME", "librestir_b200.so"))
EXTRA_DEFINES = [d for d in os.environ.get("RSTR_DEFINES", "").split() if d]
SOURCES = ["capi.cu", "strip_group.cu", "bvh_gpu.cu", "denoise.cu", "gi.cu", "kernels.cu", "scene_host.cpp", "scene_file.cpp", "bvh_fast.cpp", "image_io.cpp", "image_jpeg.cpp"]
HEADERS = ["kernels.h", "gi_kernels.inl", "camera_dev.h", "capi_internal.h", "device_types.h", "scene_host.h", "vecmath.h", os.path.join("..", "..", "include", "restir_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = "/usr/bin/g++"


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_id() -> str:
    """SHA-1 over the library's sources and headers (rstr_build_id)."""
    import hashlib

    h = hashlib.sha1()
    for s in SOURCES + HEADERS:
        with open(os.path.join(CSRC, s), "rb") as f:
            h.update(f.read())
    h.update(" ".join(EXTRA_DEFINES).encode() + os.environ.get("RSTR_FMAD", "false").encode())
    return h.hexdigest()[:12]


def nvcc_command(verbose: bool = False) -> list[str]:
    cmd = [
        NVCC, "-std=c++17", "-O3", "-shared",
        "-ccbin", HOST_CXX,
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-lineinfo", "-fmad=" + os.environ.get("RSTR_FMAD", "false"),       # RSTR_FMAD=true: the contraction experiment of scripts/ref_cuda_compare.py, never the product
        "-Xcompiler", "-fPIC,-fopenmp,-ffp-contract=off,-fno-fast-math,-O2",
        "-I", os.path.join(HERE, "..", "include"),
        "-o", LIB,
    ]
    cmd += EXTRA_DEFINES
    cmd += ['-DRSTR_BUILD_ID="%s"' % build_id()]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    cmd += ["-lgomp"]
    return cmd


def build(force: bool = False, verbose: bool = False, loading: bool = False) -> str:
    # a side-by-side experimental build (RSTR_LIBNAME) is made explicitly with its own RSTR_DEFINES; merely LOADING it must never
    # rebuild it with whatever defines the loading process happens to have
    if loading and "RSTR_LIBNAME" in os.environ and os.path.exists(LIB):
        return LIB
    if force or _stale():
        cmd = nvcc_command(verbose)
        if verbose:
            print(" ".join(cmd))
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed building librestir_b200.so")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv)
    print(LIB)
