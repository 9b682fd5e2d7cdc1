"""TEST INFRASTRUCTURE ONLY: builds tests/emu/_build/libkernels_emu.so -- the device code of restir_b200/csrc/kernels.cu compiled by g++
(see cuda_host_shim.h) -- for tests/test_device_code_on_host.py.  Same floating-point flags as the oracle (no contraction, no fast-math)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "restir_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libkernels_emu.so")
SOURCES = [os.path.join(HERE, "kernels_emu.cpp"), os.path.join(CSRC, "scene_host.cpp"), os.path.join(CSRC, "bvh_fast.cpp")]
DEPS = SOURCES + [os.path.join(HERE, "cuda_host_shim.h")] + [os.path.join(CSRC, f) for f in
                                                             ("kernels.cu", "gi_kernels.inl", "denoise.cu", "capi_internal.h", "kernels.h", "device_types.h", "vecmath.h", "camera_dev.h", "scene_host.h")]


def build(force: bool = False) -> str:
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in DEPS):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-w", "-shared",
           "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-I", "/usr/local/cuda/include", "-x", "c++"] + SOURCES + ["-o", OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed building libkernels_emu.so:\n" + r.stderr[-4000:])
    return OUT


if __name__ == "__main__":
    print(build(force=True))
