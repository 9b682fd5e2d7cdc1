"""Fuzz the library's kernels against the oracle without a GPU: random triangle soups (clustered, grid-snapped vertices: shared edges, coplanar
faces, hits at equal distance = near ties; slivers; Lambertian / metal / glass / emitters; random cameras and ragged resolutions) through
the kernels run as 32-lane warps on the CPU (tests/emu), every buffer of two frames compared bit for bit with the oracle: the direct path
(RIS / temporal / spatial / spatiotemporal, 1-3 passes, staged and fused pipelines) and ReSTIR GI (ray queues / staged / one kernel).

    python scripts/fuzz_kernels_on_cpu.py <first seed> <last seed> [scale [address|thread|undefined|stack4]]       # ~100 scenes per second

tests/test_device_code_on_host.py::test_fuzzed_scenes_as_warps_match_oracle runs a bounded range of seeds."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emu")):
    sys.path.insert(0, _p)
import numpy as np
import helpers
from emu import Emu
from oracle import oracle as orc_mod
from restir_b200 import scenes

SCALE = 1.0


def soup(seed):
    """random triangle soup: clustered + grid-snapped vertices (shared edges, coplanar and coincident-distance hits = near ties), slivers,
    random emitters / metals / glass, random camera"""
    r = np.random.default_rng(seed)
    T = int(r.integers(9, 120))
    snap = r.choice([0.0, 0.25, 0.5])
    c = r.uniform(-1.5, 1.5, (T, 1, 3)).astype(np.float32)
    v = (c + r.normal(0, r.choice([0.05, 0.4, 1.0]), (T, 3, 3))).astype(np.float32)
    if snap:
        v = (np.round(v / snap) * snap).astype(np.float32)
    if r.random() < 0.35:
        # a stack of overlapping coplanar triangles (shifted in their plane): a ray through the overlap hits them all at the same distance,
        # more near ties than a ray keeps -> undecided -> the fix-up kernels (reference-order walk)
        k = int(r.integers(3, 9))
        base = r.uniform(-1, 1, (3, 3)).astype(np.float32) * 1.5
        e1, e2 = base[1] - base[0], base[2] - base[0]
        stack = np.stack([base + (e1 * r.uniform(-0.2, 0.2) + e2 * r.uniform(-0.2, 0.2)).astype(np.float32) for _ in range(k)])
        v = np.concatenate([v, stack.astype(np.float32)])
        T = len(v)
    # drop degenerate triangles (the reference's builder needs distinct centroids; zero-area triangles are culled by the determinant test anyway)
    e1, e2 = v[:, 1] - v[:, 0], v[:, 2] - v[:, 0]
    area = np.linalg.norm(np.cross(e1, e2), axis=1)
    cen = v.mean(1)
    _, first = np.unique(np.round(cen, 5), axis=0, return_index=True)
    keep = np.zeros(T, bool); keep[first] = True
    keep &= area > 1e-4
    v = v[keep]
    T = len(v)
    if T < 9:
        return None
    verts = v.reshape(-1, 3)
    scale = np.float32(SCALE)                    # --scale: the whole scene (and the camera) far from the unit scale the epsilons were chosen at
    verts = (verts * scale).astype(np.float32)
    mats = scenes.make_materials([(scenes.LAMBERTIAN, (0.7, 0.6, 0.5), 0.0, 1.0), (scenes.METALLIC_WORKFLOW, (0.9, 0.8, 0.7), float(r.uniform(0, 1)), float(r.uniform(0.05, 1))),
                                  (scenes.DIELECTRIC, (0.95, 0.95, 1.0), 0.0, 0.0), (scenes.LIGHT, (8, 7, 6), 0.0, 1.0), (scenes.LIGHT, (2, 9, 3), 0.0, 1.0)])
    ids = r.choice(5, T, p=[0.45, 0.2, 0.1, 0.15, 0.1]).astype(np.int32)
    eye = tuple(float(x) * float(scale) for x in r.uniform(-3, 3, 3))
    rot = (float(r.uniform(-180, 180)), float(r.uniform(-60, 60)), 0.0)
    W, H = int(r.integers(17, 49)), int(r.integers(9, 33))
    sd = scenes.SceneData("soup%d" % seed, verts, scenes._face_normals(verts), np.zeros((3 * T, 2), np.float32), ids, mats, ["a", "b", "c", "d", "e"],
                            eye=eye, rotation=rot, fovy=float(r.uniform(15, 40)), resolution=(W, H))
    if seed % 5 == 4:
        sd = scenes.with_textures(sd, env=bool(seed % 2))      # base-colour / metallic / roughness / normal maps, procedural pattern, environment-map light
    return sd

def run(e, orc, lo, hi, verbose=True):
    bad = 0; ran = 0; fixups = 0; lit = 0; t0 = time.time()
    for seed in range(lo, hi):
        sd = soup(seed)
        if sd is None: continue
        ran += 1
        reuse = seed % 4; passes = 1 + seed % 3 if reuse & 2 else 1
        try:
            want = helpers.run_oracle(orc, sd, 2, reuse, passes=passes, light_index=True)
        except Exception as ex:
            print(seed, "oracle rejected:", str(ex)[:80])
            continue
        e.traced_build = (None, None, 0, 1)[seed % 4] if sd.num_tris >= 8 else None    # half of the scenes on a tree built by the device-side builders' kernels
        try:
            got, fix = e.run_di(sd, 2, reuse, passes=passes, light_index=True, pipeline=seed % 2, bands=1 + (seed // 2) % 3, drain=bool(seed % 3))
            m = {n: helpers.mismatches(got[f][n], want[f][n]) for f in range(2) for n in want[f] if helpers.mismatches(got[f][n], want[f][n])}
            if seed % 7 == 0 and sd.env_map < 0:                 # the unbiased mode (triangle lights only)
                wu = helpers.run_oracle(orc, sd, 2, reuse, passes=passes, unbiased=True, light_index=True)
                gu, _ = e.run_di(sd, 2, reuse, passes=passes, unbiased=True, light_index=True)
                m.update({"unbiased " + n: helpers.mismatches(gu[f][n], wu[f][n]) for f in range(2) for n in wu[f] if helpers.mismatches(gu[f][n], wu[f][n])})
            if seed % 6 == 0 and reuse == 3 and sd.resolution[1] >= 16:    # strips with random cuts (halo: radius 5 px + the motion of one orbit step)
                H = sd.resolution[1]
                cut = sorted(set([0, H] + [int(x) for x in np.random.default_rng(seed).integers(1, H, 2)]))
                gs, miss = e.run_di_strips(sd, 2, tuple(cut), halo=H, reuse=3, radius=5.0, passes=passes)
                m.update({"strips " + n: helpers.mismatches(gs[f][n], want[f][n]) for f in range(2) for n in gs[f] if helpers.mismatches(gs[f][n], want[f][n])})
                if any(miss):
                    m["halo_miss"] = miss
        finally:
            pass
        gi_depth, gi_reuse = (3, 1) if seed % 3 else ((seed // 3) % 7, (seed // 21) % 2)     # a third of the scenes: trace depth 0..6, with / without history
        wg = helpers.run_oracle_gi(orc, sd, 2, gi_depth, gi_reuse, accumulate=bool(seed % 2))
        try:
            gg, _ = e.run_gi(sd, 2, gi_depth, gi_reuse, accumulate=bool(seed % 2), staged=3 + seed % 3)
        finally:
            e.traced_build = None
        mg = {n: helpers.mismatches(gg[f][n], wg[f][n]) for f in range(2) for n in wg[f] if helpers.mismatches(gg[f][n], wg[f][n])}
        if seed % 8 == 1:                                        # the image-space filters on the frames of this scene
            import test_denoiser

            kind = "svgf" if seed % 16 == 1 else "eaw"
            wd = test_denoiser.oracle_frames(orc, sd, 3, kind, modulate=bool(seed % 3))
            gd = e.run_denoiser(sd, 3, kind, modulate=bool(seed % 3))
            for f in range(3):
                for n in ("rgb", "var"):
                    if wd[f][n] is not None and helpers.mismatches(gd[f][n], wd[f][n]):
                        mg["denoiser %s %s f%d" % (kind, n, f)] = helpers.mismatches(gd[f][n], wd[f][n])
        fixups += fix > 0
        lit += bool((want[-1]["radiance"].sum(1) > 0).any())
        if m or mg:
            bad += 1
            print("MISMATCH seed", seed, sd.num_tris, sd.resolution, reuse, passes, m, mg, "fix", fix, flush=True)
    if verbose:
        print("scenes %d, mismatching %d, with fix-up pixels %d, with lit pixels %d, %.1f s" % (ran, bad, fixups, lit, time.time() - t0))
    return ran, bad


if __name__ == "__main__":
    if len(sys.argv) > 3:
        SCALE = float(sys.argv[3])
    variant = sys.argv[4] if len(sys.argv) > 4 else ""           # address / thread / undefined (run with the runtime in LD_PRELOAD, see
    emu = Emu(small_stack=True) if variant == "stack4" else Emu(sanitize=variant)     # tests/test_kernels_under_sanitizers.py), or stack4
    run(emu, orc_mod.Oracle("port"), int(sys.argv[1]), int(sys.argv[2]))
