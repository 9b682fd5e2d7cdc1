"""TEST INFRASTRUCTURE ONLY: run by tests/test_kernels_under_sanitizers.py in a child process with the sanitizer runtime preloaded.
Drives every kernel family of the library, as grids of 32-lane warps on the CPU, through the ASan / TSan build of the emulation library."""
import dataclasses
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), HERE):
    sys.path.insert(0, p)

import helpers  # noqa: E402
from emu import Emu  # noqa: E402
from restir_b200 import scenes  # noqa: E402


def main():
    e = Emu(sanitize=sys.argv[1])
    what = sys.argv[2] if len(sys.argv) > 2 else "all"
    gen = scenes.procedural(3, 2000, 100, (64, 48))
    tex = dataclasses.replace(helpers.textured_scenes()["gen2000_tex"], resolution=(48, 32))
    glass = dataclasses.replace(helpers.gi_scenes()["cornell_glass"], resolution=(48, 32))
    if what == "fault":
        e.run_di(gen, 1, 3, passes=2)            # two spatial passes: the first one publishes reservoirs (the last one only shades)
        return
    e.run_di(gen, 2, 3, passes=2, drain=True, light_index=True)          # staged: k_primary, k_candidates, k_shadow<drain>, k_temporal, fix-up, k_restir_b x 2, exports
    e.run_di(tex, 2, 1, drain=False)                                      # textures / environment map, k_shadow<no drain>, temporal only
    e.run_di(gen, 2, 3, pipeline=1)                                       # fused: k_gbuffer_restir_a
    e.run_di(gen, 2, 2, pipeline=2)                                       # split: k_gbuffer, k_restir_a
    e.run_di(glass, 2, 3, pipeline=3)                                     # reference-order kernels
    e.run_di(gen, 2, 3, unbiased=True, passes=2)                          # k_temporal_unb, k_restir_b_unb (two passes: a pixel's RNG word is rewritten between them while neighbours read its wo)
    e.run_di(gen, 2, 0, ptdirect=True, want=("radiance",))               # k_ptdirect
    for mode in (3, 4, 5):                                                # GI: ray queues, staged, one kernel
        e.run_gi(gen, 2, 3, 1, accumulate=True, staged=mode)
    e.run_gi(glass, 2, 4, 1, staged=3)
    import test_device_code_on_host as t                                  # ragged resolution, single triangle, no lights, nothing in view
    for name, (sd, orbit) in t.edge_scenes().items():
        e.run_di(sd, 2, 3, orbit=orbit, light_index=True)
        e.run_di(sd, 2, 3, orbit=orbit, pipeline=1)
        e.run_gi(sd, 2, 3, 1, orbit=orbit, staged=3)
        e.run_gi(sd, 2, 3, 1, orbit=orbit, staged=5)
    e.run_di_strips(scenes.procedural(3, 2000, 100, (64, 56)), 2, (0, 13, 32, 56), halo=10, reuse=3, radius=5.0, passes=2)   # strips: plane indices relative to the resident rows
    e.run_denoiser(t.edge_scenes()["ragged"][0], 2, "svgf")
    for mode in (0, 1):                                                   # bvh_gpu.cu: the device-side builders' kernels as 256-thread blocks, then frames on their trees
        e.traced_build = mode
        e.run_di(gen, 1, 3)
        e.run_gi(gen, 1, 3, 1, staged=3)
    e.traced_build = None
    e.run_denoiser(gen, 2, "eaw")
    e.run_denoiser(gen, 2, "svgf", modulate=True)
    print("sanitizer child: all kernel families ran")


if __name__ == "__main__":
    main()
