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
                                                             ("kernels.cu", "gi_kernels.inl", "denoise.cu", "bvh_gpu.cu", "capi_internal.h", "kernels.h", "device_types.h", "vecmath.h", "camera_dev.h", "scene_host.h")]


def build(force: bool = False, sanitize: str = "", small_stack: bool = False) -> str:
    """sanitize = "address" / "thread": the same library instrumented by ASan / TSan (tests/test_kernels_under_sanitizers.py runs it in a
    child process with the sanitizer runtime preloaded): the memcheck / racecheck of the kernels' code that needs no GPU."""
    out = OUT if not sanitize else OUT.replace(".so", "_%s.so" % sanitize)
    if small_stack:                              # RS_SMEM_STACK = 4: the per-lane stacks spill to their local arrays after four entries
        out = out.replace(".so", "_stack4.so")
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in DEPS):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    opt = ["-O2"] if not sanitize else ["-O1", "-g", "-fno-omit-frame-pointer", "-fsanitize=" + sanitize]
    if small_stack:
        opt.append("-DRS_SMEM_STACK=4")
    if sanitize == "thread":
        opt.append("-DEMU_STD_THREADS")          # libgomp's barriers are invisible to TSan: the warps run on std::threads instead
    cmd = ["/usr/bin/g++", "-std=c++17"] + opt + ["-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-w", "-shared",
           "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-I", "/usr/local/cuda/include", "-x", "c++"] + SOURCES + ["-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed building %s:\n" % os.path.basename(out) + r.stderr[-4000:])
    return out


def sanitizer_runtime(sanitize: str) -> str:
    """Path of libasan.so / libtsan.so for LD_PRELOAD (the interpreter itself is not instrumented)."""
    name = {"address": "libasan.so", "thread": "libtsan.so", "undefined": "libubsan.so"}[sanitize]
    return subprocess.run(["/usr/bin/gcc", "-print-file-name=" + name], capture_output=True, text=True).stdout.strip()


if __name__ == "__main__":
    print(build(force=True))
