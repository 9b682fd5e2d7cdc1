"""CPU: the library's kernels under AddressSanitizer, ThreadSanitizer and UndefinedBehaviorSanitizer.

compute-sanitizer cannot attach on this GPU pool (profiles/r02_sanitizer_closed_on_pool.txt), so the memcheck / racecheck of SURVEY section 5 is
run on the kernels' CODE instead: tests/emu compiles kernels.cu + gi_kernels.inl + denoise.cu with g++ -fsanitize=address / thread and launches
every kernel family as grids of 32-lane warps (tests/emu/sanitizer_child.py).  Every global-memory access of a kernel is then an instrumented
access to a heap buffer of exactly the plane's size (ASan: out-of-bounds reads / writes of any plane, queue, list or table), and warps on
different OS threads are checked against each other (TSan: two warps touching the same word without an atomic; the lanes of one warp run one
after the other and are ordered at every warp intrinsic; UBSan: signed overflow, shifts, misaligned vector loads, float -> int casts out of
range).  ASan and TSan each also get a negative control -- a plane one pixel short, and the
reference's own in-place spatial reuse (restir.cu:192-196) -- that it must report, so a clean run means something.

The sanitizer runtime has to be loaded before the interpreter starts, hence the child processes.  Test infrastructure only."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
import build as emu_build  # noqa: E402

CHILD = os.path.join(ROOT, "tests", "emu", "sanitizer_child.py")


def run_child(tool, what, fault=None):
    if not os.path.isabs(emu_build.sanitizer_runtime(tool)):
        pytest.skip("gcc has no %s sanitizer runtime here" % tool)
    emu_build.build(sanitize=tool)
    from restir_b200 import build as product_build

    product_build.build()                     # the child uses the library's host-side camera functions: never let it run nvcc under a sanitizer preload
    env = dict(os.environ)
    env["LD_PRELOAD"] = emu_build.sanitizer_runtime(tool)
    env["ASAN_OPTIONS"] = "detect_leaks=0:abort_on_error=0"        # the interpreter's own allocations are not the subject
    env["TSAN_OPTIONS"] = "report_signal_unsafe=0:exitcode=66"
    env["UBSAN_OPTIONS"] = "print_stacktrace=1"
    env["OMP_NUM_THREADS"] = "1" if tool == "thread" else env.get("OMP_NUM_THREADS", "4")   # TSan: the warps run on std::threads, nothing else is parallel
    env.pop("EMU_INJECT_FAULT", None)
    if fault:
        env["EMU_INJECT_FAULT"] = fault
    cmd = [sys.executable, CHILD, tool, what]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=1500)
    # a sanitizer runtime that cannot place its shadow memory (address-space randomisation with many random bits) is an environment problem,
    # not a finding: once more without randomisation, else skip
    env_trouble = ("Shadow memory range interleaves", "unexpected memory mapping", "FATAL: ThreadSanitizer", "ReserveShadowMemoryRange failed")
    if any(t in r.stderr for t in env_trouble):
        import shutil

        if shutil.which("setarch"):
            r = subprocess.run(["setarch", "x86_64", "-R"] + cmd, capture_output=True, text=True, env=env, timeout=1500)
        if any(t in r.stderr for t in env_trouble):
            pytest.skip("the %s sanitizer runtime cannot map its shadow memory in this environment" % tool)
    return r


@pytest.mark.parametrize("tool,marker", [("address", "ERROR: AddressSanitizer"), ("thread", "WARNING: ThreadSanitizer"), ("undefined", "runtime error")])
def test_every_kernel_family_is_clean(tool, marker):
    r = run_child(tool, "all")
    assert marker not in r.stderr, r.stderr[-3000:]
    assert r.returncode == 0, r.stderr[-3000:]
    assert "all kernel families ran" in r.stdout


def test_asan_reports_a_plane_one_pixel_short():
    r = run_child("address", "fault", fault="short_plane")
    assert r.returncode != 0
    assert "heap-buffer-overflow" in r.stderr and "kernels.cu" in r.stderr, r.stderr[-3000:]


def test_tsan_reports_the_references_in_place_spatial_reuse():
    r = run_child("thread", "fault", fault="in_place_spatial")
    assert "WARNING: ThreadSanitizer: data race" in r.stderr and "k_restir_b" in r.stderr, r.stderr[-3000:]
    assert r.returncode != 0
