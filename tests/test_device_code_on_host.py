"""CPU: the DEVICE code of the library's kernels (restir_b200/csrc/kernels.cu + gi_kernels.inl), compiled by g++ through
tests/emu/cuda_host_shim.h, against the oracle -- bit for bit, because here both sides use the same libm.

Two levels.  (1) Per pixel: the bodies of the GI kernels run one "thread" per call (transcription errors, RNG ledger, reservoir logic, the
per-lane walk of the traced tree, the fix-up path, the export layout).  (2) Per kernel ("as warps"): the __global__ functions themselves
launched as grids of 32-lane warps -- the lanes of a warp are fibers on one OS thread, warp intrinsics exchange the lanes' values at each
synchronisation point, __shared__ is one copy per warp, atomics are real across the OpenMP threads running the warps -- so the packet
walk, k_shadow's lane refill / leaf wait / cooperative drain, warp-aggregated queue appends and the persistent GI walkers run as
written: every launch form of the direct path (staged, fused, split, reference-order, unbiased), PTDirect and the three GI forms.

This needs no GPU; it is test infrastructure, not a CPU path of the product: the emulation library is built into tests/emu/_build and
nothing under restir_b200/ can load it.  The GPU run of the same code (tests/test_gpu_parity.py, tests/test_restir_gi.py, -m gpu)
differs from this one only by libdevice's sinf / cosf / expf.
"""
import dataclasses
import os
import sys

import numpy as np
import pytest

import helpers
from restir_b200 import scenes

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))


@pytest.fixture(scope="module")
def emu():
    from emu import Emu

    return Emu()


@pytest.mark.parametrize("name", ["cornell", "cornell_metal", "cornell_glass", "gen2000", "cornell_tex"])
def test_gi_device_code_matches_golden(emu, name):
    """k_gbuffer_exact + k_restir_indirect_exact per pixel against the fixture of the reference's own code (tests/golden/gi.npz)."""
    sd = helpers.gi_scenes()[name]
    g = np.load(os.path.join(helpers.GOLDEN, "gi.npz"))
    got, _ = emu.run_gi(sd, 3, 3, 1)
    for f in range(3):
        assert helpers.mismatches(got[f]["indirect"], g["%s_f%d_indirect" % (name, f)]) == 0, f
    assert helpers.mismatches(got[2]["reservoir"], g["%s_f2_reservoir" % name]) == 0
    acc, _ = emu.run_gi(sd, 2, 2, 0, accumulate=True, orbit=False)
    assert helpers.mismatches(acc[-1]["indirect"], g["%s_acc_indirect" % name]) == 0


@pytest.mark.parametrize("name", ["gen20k", "cornell_glass", "gen2000_tex"])
def test_gi_device_code_traced_tree_matches_oracle(emu, port_oracle, name):
    """Bounce rays through traceClosestFast, shadow rays through traceOccludedFast, undecided pixels through the fix-up path: same
    bits as the oracle (and therefore as the reference-order walk), depth 5, four accumulated orbit frames with temporal reuse."""
    sd = {"gen20k": lambda: scenes.procedural(3, 20000, 1000, (320, 180)),
          "cornell_glass": lambda: dataclasses.replace(helpers.gi_scenes()["cornell_glass"], resolution=(192, 144)),
          "gen2000_tex": lambda: dataclasses.replace(helpers.textured_scenes()["gen2000_tex"], resolution=(192, 144))}[name]()
    want = helpers.run_oracle_gi(port_oracle, sd, 4, 5, 1, accumulate=True)
    exact, _ = emu.run_gi(sd, 4, 5, 1, accumulate=True)
    traced, undecided = emu.run_gi(sd, 4, 5, 1, accumulate=True, traced_tree=True)
    staged, undecided_staged = emu.run_gi(sd, 4, 5, 1, accumulate=True, staged=True)
    helpers.assert_frames_equal(exact, want, "device code, reference-order walk, vs oracle")
    helpers.assert_frames_equal(traced, want, "device code, traced tree, vs oracle")
    helpers.assert_frames_equal(staged, want, "device code, staged pipeline (k_gi_primary / k_gi_bounce / k_gi_resolve bodies), vs oracle")
    queued, undecided_queued = emu.run_gi(sd, 4, 5, 1, accumulate=True, staged=2)
    helpers.assert_frames_equal(queued, want, "device code, ray-queue pipeline (k_gi_head / walks / k_gi_tail bodies), vs oracle")
    assert undecided_queued <= undecided
    if name == "gen20k":
        assert undecided > 0            # the fix-up path was exercised
        assert undecided_staged == undecided
    assert (want[-1]["indirect"].sum(1) > 0).mean() > 0.2


@pytest.mark.parametrize("staged", [True, 2])
@pytest.mark.parametrize("depth,reuse", [(0, 1), (1, 0), (2, 1)])
def test_gi_staged_device_code_shallow_depths(emu, port_oracle, depth, reuse, staged):
    """The staged bodies at depth 0 (no path leaves k_gi_primary), 1 (paths end in their first bounce) and 2, with and without history."""
    sd = dataclasses.replace(helpers.gi_scenes()["cornell_metal"], resolution=(96, 72))
    want = helpers.run_oracle_gi(port_oracle, sd, 3, depth, reuse, accumulate=True)
    got, _ = emu.run_gi(sd, 3, depth, reuse, accumulate=True, staged=staged)
    helpers.assert_frames_equal(got, want, "staged device code, depth %d" % depth)


@pytest.mark.parametrize("name", ["gen20k", "cornell_glass"])
def test_gi_queued_kernels_as_warps_match_oracle(emu, port_oracle, name):
    """The GI KERNELS of all three launch forms (k_gi_primary with its packet walk, k_gi_head, k_gi_walk_shadow, k_gi_walk_closest, k_gi_tail,
    k_gi_bounce, k_gi_resolve, k_restir_indirect, k_restir_indirect_fix) run as grids of 32-lane warps on the CPU (tests/emu/cuda_host_shim.h: lanes are fibers, warp intrinsics exchange their values): persistent warps
    with lane refill from the ray lists, lanes waiting for each other at leaves, warp-aggregated appends -- same bits as the oracle."""
    sd = {"gen20k": lambda: scenes.procedural(3, 20000, 1000, (160, 90)),
          "cornell_glass": lambda: dataclasses.replace(helpers.gi_scenes()["cornell_glass"], resolution=(96, 72))}[name]()
    want = helpers.run_oracle_gi(port_oracle, sd, 3, 4, 1, accumulate=True)
    got, undecided = emu.run_gi(sd, 3, 4, 1, accumulate=True, staged=3)
    helpers.assert_frames_equal(got, want, "ray-queue kernels as warps vs oracle")
    staged, undecided_staged = emu.run_gi(sd, 3, 4, 1, accumulate=True, staged=4)
    helpers.assert_frames_equal(staged, want, "staged kernels (k_gi_bounce) as warps vs oracle")
    fused, undecided_fused = emu.run_gi(sd, 3, 4, 1, accumulate=True, staged=5)
    helpers.assert_frames_equal(fused, want, "k_restir_indirect as warps vs oracle")
    assert undecided_fused == undecided_staged
    if name == "gen20k":
        assert 0 < undecided <= undecided_staged          # the fix-up kernel ran on the queue the kernels wrote
    assert (want[-1]["indirect"].sum(1) > 0).mean() > 0.2


@pytest.mark.parametrize("name,res,reuse,passes,drain", [("cornell", (64, 48), 3, 1, True), ("gen2000", (160, 90), 3, 2, False), ("gen20k", (160, 90), 1, 1, True),
                                                         ("gen2000_tex", (96, 72), 3, 1, True), ("cornell_metal", (64, 48), 0, 1, False), ("gen2000", (96, 72), 2, 3, True)])
def test_di_kernels_as_warps_match_oracle(emu, port_oracle, name, res, reuse, passes, drain):
    """The direct path's KERNELS -- k_primary (packet walk of the centre + jittered rays of every 8x4 tile: G-buffer and the shaded-pixel
    queue), k_candidates, k_shadow (persistent warps, lane refill, leaf wait, cooperative drain), k_temporal, the fix-up kernel, k_restir_b per
    spatial pass and the export kernels -- launched as grids of 32-lane warps on the CPU in the order launchPhaseAStaged / launchRestirB use:
    every buffer of every frame bit-identical to the oracle (RIS only, temporal, spatial, spatiotemporal; 1-3 passes)."""
    sd = {"gen2000": lambda: scenes.procedural(3, 2000, 100, res), "gen20k": lambda: scenes.procedural(3, 20000, 1000, res),
          "gen2000_tex": lambda: dataclasses.replace(helpers.textured_scenes()["gen2000_tex"], resolution=res)}.get(
              name, lambda: dataclasses.replace(helpers.gi_scenes()[name], resolution=res))()
    want = helpers.run_oracle(port_oracle, sd, 3, reuse, passes=passes, light_index=True)
    got, _ = emu.run_di(sd, 3, reuse, passes=passes, drain=drain, light_index=True)
    helpers.assert_frames_equal(got, want, "direct-path kernels as warps vs oracle")
    assert (want[-1]["radiance"].sum(1) > 0).mean() > 0.2


@pytest.mark.parametrize("passes", [1, 2, 3])
def test_strips_of_kernels_as_warps_equal_the_full_frame(emu, port_oracle, passes):
    """The multi-GPU decomposition with the kernels as warps: four uneven strips that render only their own rows, their G-buffer and
    reservoir halo rows copied where the library pushes them to the neighbour (after phase A, between spatial passes, the history after the
    frame) == the oracle's single full frame, bit for bit, with no read outside the resident rows -- strip-local plane indices, rowResident,
    the global pixel / RNG indices of restir.cu:126-127."""
    sd = scenes.procedural(3, 2000, 100, (96, 80))
    want = helpers.run_oracle(port_oracle, sd, 3, 3, radius=6.0, passes=passes, light_index=True)
    got, miss = emu.run_di_strips(sd, 3, (0, 21, 40, 64, 80), halo=12, reuse=3, radius=6.0, passes=passes)
    for f in range(3):
        for n in got[f]:
            assert helpers.mismatches(got[f][n], want[f][n]) == 0, (f, n)
    assert miss == [0, 0, 0, 0]


@pytest.mark.parametrize("mode", [0, 1])
def test_device_side_tree_build_as_blocks_gives_the_same_frames(emu, port_oracle, mode):
    """rstr_scene_build_traced_gpu's kernels (bvh_gpu.cu: Morton codes, PLOC clustering rounds or the binary radix tree with its bottom-up
    pass, the leaf rule, node emission, triangle re-ordering) run as 256-thread blocks on the CPU (__syncthreads = the block's
    synchronisation point, __shared__ per block, real atomics); the direct path's and the GI kernels then trace THAT tree: the frames are
    the oracle's bit for bit, because what a ray reports does not depend on the tree it walks."""
    sd = scenes.procedural(3, 2000, 100, (96, 72))
    want = helpers.run_oracle(port_oracle, sd, 2, 3, light_index=True)
    want_gi = helpers.run_oracle_gi(port_oracle, sd, 2, 3, 1)
    emu.traced_build = mode
    try:
        got, _ = emu.run_di(sd, 2, 3, light_index=True)
        got_gi, _ = emu.run_gi(sd, 2, 3, 1, staged=3)
    finally:
        emu.traced_build = None
    helpers.assert_frames_equal(got, want, "direct path on the device-built tree (mode %d)" % mode)
    helpers.assert_frames_equal(got_gi, want_gi, "GI ray queues on the device-built tree (mode %d)" % mode)


@pytest.mark.parametrize("bands", [2, 3, 8])
def test_di_staged_bands_as_warps_match_oracle(emu, port_oracle, bands):
    """rstr_frame_set_bands: the staged pipeline over horizontal bands of whole 8-row tiles, each with its own shaded-pixel queue and cursor
    (launchPhaseAStaged), incl. a last band that is shorter and a band count the image cannot fill."""
    sd = scenes.procedural(3, 2000, 100, (96, 52))
    want = helpers.run_oracle(port_oracle, sd, 3, 3, light_index=True)
    got, _ = emu.run_di(sd, 3, 3, bands=bands, light_index=True)
    helpers.assert_frames_equal(got, want, "staged pipeline in %d bands" % bands)


def test_fuzzed_scenes_as_warps_match_oracle(emu, port_oracle):
    """scripts/fuzz_kernels_on_cpu.py over a bounded range of seeds: random triangle soups with shared edges, coplanar stacks (near ties
    beyond what a ray keeps: the fix-up kernels), slivers, metal / glass / emitters, random cameras and ragged resolutions, through the direct
    path (all reuse modes, 1-3 passes, staged / fused) and the three GI forms, every buffer bit for bit."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import fuzz_kernels_on_cpu as fz

    ran, bad = fz.run(emu, port_oracle, 0, 400, verbose=False)
    assert ran > 300 and bad == 0


def test_spilling_stacks_as_warps_match_oracle(port_oracle):
    """The same kernels built with RS_SMEM_STACK = 4 instead of 24: every per-lane walk (reference-order walk, traceClosestFast /
    traceOccludedFast, k_shadow and its cooperative drain, the GI walkers) runs past the shared-memory part of its stack into the
    local-array part -- on the bench scenes only the deepest rays do -- and must report the same hits."""
    from emu import Emu

    e = Emu(small_stack=True)
    sd = scenes.procedural(3, 20000, 1000, (96, 54))
    want = helpers.run_oracle(port_oracle, sd, 2, 3, light_index=True)
    for pipeline, drain in ((0, True), (0, False), (3, True)):
        got, _ = e.run_di(sd, 2, 3, pipeline=pipeline, drain=drain, light_index=True)
        helpers.assert_frames_equal(got, want, "RS_SMEM_STACK = 4, pipeline %d" % pipeline)
    want_gi = helpers.run_oracle_gi(port_oracle, sd, 2, 4, 1)
    for mode in (3, 4, 5):
        got, _ = e.run_gi(sd, 2, 4, 1, staged=mode)
        helpers.assert_frames_equal(got, want_gi, "RS_SMEM_STACK = 4, GI form %d" % mode)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import fuzz_kernels_on_cpu as fz

    ran, bad = fz.run(e, port_oracle, 1000, 1150, verbose=False)
    assert ran > 100 and bad == 0


def edge_scenes():
    """tests/test_gpu_parity.py::test_edge_cases: ragged resolution, a single (emissive) triangle, no lights, a camera that sees nothing,
    an exactly axis-aligned centre ray, a camera looking straight down."""
    v = np.asarray([(-1, 0, -2), (1, 0, -2), (0, 1.5, -2)], np.float32)
    one = scenes.SceneData("one", v, scenes._face_normals(v), np.zeros((3, 2), np.float32), np.zeros(1, np.int32),
                           scenes.make_materials([(scenes.LIGHT, (5, 5, 5), 0.0, 1.0)]), ["l"], eye=(0, 0.5, 2), rotation=(-90, 0, 0), fovy=30.0, resolution=(40, 32))
    away = scenes.cornell_box((48, 36))
    away.rotation = (90.0, 0.0, 0.0)
    down = scenes.cornell_box((33, 33))
    down.eye, down.rotation = (0.05, 1.9, 0.02), (-90.0, -89.99, 0.0)
    return {"ragged": (scenes.cornell_box((37, 23), metal_tall_box=True), True), "one_triangle": (one, True), "no_lights": (helpers.five_triangles(), True),
            "all_miss": (away, True), "axis_aligned": (scenes.cornell_box((33, 33)), False), "looking_down": (down, False)}


@pytest.mark.parametrize("name", ["ragged", "one_triangle", "no_lights", "all_miss", "axis_aligned", "looking_down"])
@pytest.mark.parametrize("pipeline", [0, 1])
def test_di_edge_cases_as_warps_match_oracle(emu, port_oracle, name, pipeline):
    """The edge cases of the GPU suite through the staged and the fused kernels as warps (partial tiles, a leaf as the BVH root, an empty
    light table, empty queues for the persistent kernels, the special cases of bvh.h:91-123)."""
    sd, orbit = edge_scenes()[name]
    want = helpers.run_oracle(port_oracle, sd, 2, 3, orbit=orbit, light_index=True)
    got, _ = emu.run_di(sd, 2, 3, orbit=orbit, pipeline=pipeline, light_index=True)
    helpers.assert_frames_equal(got, want, "edge case %s, pipeline %d" % (name, pipeline))


@pytest.mark.parametrize("name", ["ragged", "one_triangle", "no_lights", "all_miss"])
def test_gi_edge_cases_as_warps_match_oracle(emu, port_oracle, name):
    """... and through the three GI forms' kernels."""
    sd, orbit = edge_scenes()[name]
    want = helpers.run_oracle_gi(port_oracle, sd, 2, 3, 1, orbit=orbit)
    for mode, what in ((3, "ray queues"), (4, "staged"), (5, "one kernel")):
        got, _ = emu.run_gi(sd, 2, 3, 1, orbit=orbit, staged=mode)
        helpers.assert_frames_equal(got, want, "GI edge case %s, %s" % (name, what))


@pytest.mark.parametrize("pipeline", [1, 2, 3])
@pytest.mark.parametrize("name,reuse", [("gen2000", 3), ("cornell_glass", 1)])
def test_di_other_pipelines_as_warps_match_oracle(emu, port_oracle, name, reuse, pipeline):
    """The other launch forms of the direct path as warps: 1 the fused kernel (k_gbuffer_restir_a, what tiny scenes use), 2 G-buffer and phase A
    as separate launches (k_gbuffer, k_restir_a: strips), 3 the validation mode (k_gbuffer_exact, k_restir_a_exact), each with its fix-up
    kernel and k_restir_b."""
    sd = scenes.procedural(3, 2000, 100, (96, 72)) if name == "gen2000" else dataclasses.replace(helpers.gi_scenes()[name], resolution=(64, 48))
    want = helpers.run_oracle(port_oracle, sd, 3, reuse, light_index=True)
    got, _ = emu.run_di(sd, 3, reuse, pipeline=pipeline, light_index=True)
    helpers.assert_frames_equal(got, want, "direct-path kernels as warps, pipeline %d, vs oracle" % pipeline)


@pytest.mark.parametrize("reuse,passes", [(3, 1), (2, 2), (1, 1)])
def test_di_unbiased_kernels_as_warps_match_oracle(emu, port_oracle, reuse, passes):
    """RstrParams::unbiased through the staged pipeline's kernels (k_temporal_unb, k_restir_b_unb; no k_shadow) as warps."""
    sd = scenes.procedural(3, 2000, 100, (96, 72))
    want = helpers.run_oracle(port_oracle, sd, 3, reuse, passes=passes, unbiased=True, light_index=True)
    got, _ = emu.run_di(sd, 3, reuse, passes=passes, unbiased=True, light_index=True)
    helpers.assert_frames_equal(got, want, "unbiased kernels as warps vs oracle")
    assert (want[-1]["radiance"].sum(1) > 0).mean() > 0.2


def test_ptdirect_kernels_as_warps_match_golden(emu):
    """k_ptdirect + k_ptdirect_fix as warps against the fixture of the reference's own pathTraceDirect (tests/golden/ptdirect.npz)."""
    g = np.load(os.path.join(helpers.GOLDEN, "ptdirect.npz"))
    for name, sd in helpers.test_scenes().items():
        got, _ = emu.run_di(sd, 2, 0, accumulate=True, orbit=False, ptdirect=True, looper0=100, want=("radiance",))
        assert helpers.mismatches(got[-1]["radiance"], g[name]) == 0, name


@pytest.mark.parametrize("kind,modulate", [("eaw", False), ("svgf", False), ("svgf", True)])
def test_denoiser_kernels_as_warps_match_oracle(emu, port_oracle, kind, modulate):
    """denoise.cu's kernels (k_eaw, k_temporal_accumulate, k_estimate_variance, k_filter_variance, k_eaw_svgf, k_modulate) in the launch
    order of rstr_denoiser_filter, on frames the emulated direct-path kernels rendered: with glibc's expf / powf on both sides the filtered
    colour and the variance are the oracle's bit for bit (on the GPU, libdevice: 2e-4 relative, tests/test_denoiser.py)."""
    import test_denoiser

    sd = scenes.procedural(3, 2000, 100, (96, 72))
    want = test_denoiser.oracle_frames(port_oracle, sd, 4, kind, modulate=modulate)
    got = emu.run_denoiser(sd, 4, kind, modulate=modulate)
    for f in range(4):
        assert helpers.mismatches(got[f]["radiance"], want[f]["radiance"]) == 0, f
        assert helpers.mismatches(got[f]["rgb"], want[f]["rgb"]) == 0, (f, "filtered colour")
        if kind == "svgf":
            assert helpers.mismatches(got[f]["var"], want[f]["var"]) == 0, (f, "variance")
    assert float(np.abs(got[-1]["rgb"] - got[-1]["radiance"]).max()) > 0      # the filter did something


def test_gbuffer_device_code_matches_oracle(emu, port_oracle):
    """The G-buffer the GI kernel reads (k_gbuffer_exact's per-pixel body) against the oracle's, over an orbit."""
    sd = dataclasses.replace(helpers.textured_scenes()["gen2000_tex"], resolution=(96, 72))
    got, _ = emu.run_gi(sd, 3, 1, 1, gbuffer=True)
    want = helpers.run_oracle(port_oracle, sd, 3, 0, want=helpers.GBUF)
    for f in range(3):
        for n in helpers.GBUF:
            assert helpers.mismatches(got[f][n], want[f][n]) == 0, (f, n)
