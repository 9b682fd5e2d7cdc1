"""Shared drivers for the parity tests: the same frame loop (main.cpp:146-185 with a fixed animation clock)
run through the oracle (CPU) and through the C ABI (GPU)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from oracle import oracle as orc_mod
from restir_b200 import scenes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GBUF = ("albedo", "normal", "matid", "depth", "motion")
ALL_BUFS = GBUF + ("radiance", "reservoir", "reservoir_temp")


def five_triangles() -> scenes.SceneData:
    """SURVEY.md App. D probe (i): triangles i=0..4, x = 1.5 i: (x,0,0) (x+1,0,0) (x,1,i)."""
    tris = []
    for i in range(5):
        x = 1.5 * i
        tris += [(x, 0, 0), (x + 1, 0, 0), (x, 1, i)]
    v = np.asarray(tris, np.float32)
    return scenes.SceneData("five", v, scenes._face_normals(v), np.zeros((15, 2), np.float32), np.zeros(5, np.int32),
                            scenes.make_materials([(scenes.LAMBERTIAN, (0.9, 0.9, 0.9), 0.0, 1.0)]), ["m0"],
                            eye=(3.0, 0.5, 8.0), rotation=(-90.0, 0.0, 0.0), fovy=30.0, resolution=(32, 24))


def test_scenes():
    """name -> SceneData for the small parity cases."""
    return {
        "cornell": scenes.cornell_box((48, 36)),
        "cornell_metal": scenes.cornell_box((48, 36), metal_tall_box=True),
        "gen2000": scenes.procedural(1, 2000, 100, (48, 36)),
        "five": five_triangles(),
    }


def textured_scenes():
    """Small cases with every texture branch of scene.h:78-99 and an environment map (scene.cpp:136-152)."""
    return {
        "cornell_tex": scenes.with_textures(scenes.cornell_box((48, 36)), env=True),
        "gen2000_tex": scenes.with_textures(scenes.procedural(1, 2000, 100, (48, 36)), env=True),
        "cornell_tex_noenv": scenes.with_textures(scenes.cornell_box((48, 36), metal_tall_box=True), env=False),
    }


def run_oracle(orc, sd, frames, reuse, radius=5.0, k=5, cap=20, candidates=32, accumulate=False, orbit=True, want=ALL_BUFS, light_index=False, passes=1, unbiased=False):
    """Returns a list (per frame) of {buffer name: array}."""
    W, H = sd.resolution
    so = orc.scene(sd)
    fo = so.frame(W, H)
    base = orc_mod.make_camera(sd)
    orc.lib.orc_camera_update(C.byref(base))
    prm = orc_mod.default_params(reuse=reuse, radius=radius, k=k, cap=cap, candidates=candidates, passes=passes, unbiased=unbiased)
    out = []
    for f in range(frames):
        cam = orc_mod.orbit_camera(orc, base, f) if orbit else base
        fo.gbuffer_render(cam)
        fo.restir_direct(cam, prm, f, f if accumulate else 0)
        names = list(want) + (["light_index"] if light_index else [])
        out.append({n: fo.buffer(n) for n in names})
        fo.gbuffer_update(cam)
    fo.close()
    so.close()
    return out


def run_gpu(rb, sd, frames, reuse, radius=5.0, k=5, cap=20, candidates=32, accumulate=False, orbit=True, want=ALL_BUFS, light_index=False,
            rows=None, halo=0, scene=None, exact=False, passes=1, fuse=True, staged=True, unbiased=False):   # staged=None: the library's choice by scene size
    W, H = sd.resolution
    sc = scene or rb.Scene.from_arrays(sd)
    sc.set_traversal(exact)
    fr = sc.frame(W, H, rows=rows, halo=halo)
    fr.set_fusion(fuse)
    fr.set_pipeline(staged)
    base = rb.Camera.from_scene(sd)
    prm = rb.default_params(reuse=reuse, radius=radius, k=k, cap=cap, candidates=candidates, passes=passes, unbiased=unbiased)
    out = []
    for f in range(frames):
        cam = base.orbit(f) if orbit else base
        fr.gbuffer_render(cam)
        fr.restir_direct(cam, prm, f, f if accumulate else 0)
        names = list(want) + (["light_index"] if light_index else [])
        out.append({n: fr.read(n) for n in names})
        fr.gbuffer_update(cam)
    miss = fr.halo_miss()
    fr.close()
    if scene is None:
        sc.close()
    return out, miss


def mismatches(a: np.ndarray, b: np.ndarray) -> int:
    """Number of pixels whose bytes differ."""
    av = np.ascontiguousarray(a).view(np.uint8).reshape(a.shape[0], -1)
    bv = np.ascontiguousarray(b).view(np.uint8).reshape(b.shape[0], -1)
    assert av.shape == bv.shape, (a.shape, a.dtype, b.shape, b.dtype)
    return int((av != bv).any(1).sum())


def assert_frames_equal(got, want, what=""):
    assert len(got) == len(want)
    for f, (g, w) in enumerate(zip(got, want)):
        for n in w:
            if n not in g:
                continue
            bad = mismatches(g[n], w[n])
            assert bad == 0, "%s frame %d buffer %s: %d pixels differ" % (what, f, n, bad)


# ---------------------------------------------------------------------------------------------- ReSTIR GI
def gi_scenes():
    """Small cases for ReSTIRIndirect: every Material::sample branch (material.h:242-256) and the environment-map miss."""
    import dataclasses
    glass = scenes.cornell_box((48, 36), metal_tall_box=True)
    mats = glass.materials.copy()
    mats[0]["type"] = scenes.LAMBERTIAN
    glass = dataclasses.replace(glass, name="cornell_glass", materials=np.concatenate([mats, scenes.make_materials([(scenes.DIELECTRIC, (0.95, 0.97, 1.0), 0.0, 0.0)])]),
                                material_names=list(glass.material_names) + ["glass"])
    ids = glass.material_ids.copy()
    ids[12:24] = 5                      # the short box (12 triangles after the 6 quads) becomes glass, ior 1.5
    glass = dataclasses.replace(glass, material_ids=ids)
    return {
        "cornell": scenes.cornell_box((48, 36)),
        "cornell_metal": scenes.cornell_box((48, 36), metal_tall_box=True),
        "cornell_glass": glass,
        "gen2000": scenes.procedural(1, 2000, 100, (48, 36)),
        "cornell_tex": scenes.with_textures(scenes.cornell_box((48, 36)), env=True),
    }


def run_oracle_gi(orc, sd, frames, max_depth=3, reuse=1, accumulate=False, orbit=True):
    """Per frame {indirect (P,3), reservoir (P,17)} of ReSTIRIndirect (restir.cu:448-476) in the frame loop of main.cpp:146-185."""
    W, H = sd.resolution
    so = orc.scene(sd)
    fo = so.frame(W, H)
    gi = orc_mod.OracleGI(fo)
    base = orc_mod.make_camera(sd)
    orc.lib.orc_camera_update(C.byref(base))
    out = []
    for f in range(frames):
        cam = orc_mod.orbit_camera(orc, base, f) if orbit else base
        fo.gbuffer_render(cam)
        gi.restir_indirect(cam, f, f if accumulate else 0, max_depth, reuse)
        out.append({"indirect": gi.indirect().copy(), "reservoir": gi.reservoirs().copy()})
        fo.gbuffer_update(cam)
    gi.close()
    fo.close()
    so.close()
    return out


def run_gpu_gi(rb, sd, frames, max_depth=3, reuse=1, accumulate=False, orbit=True, exact=False, bounce_exact=False, scene=None, staged=None):
    """The same loop through the C ABI (rstr_gi_*); reservoirs come back in the reference's 68-byte layout and are returned as
    (P, 17) float32 with numSamples converted to float, like the oracle's export.  Also returns the number of fix-up pixels."""
    W, H = sd.resolution
    sc = scene or rb.Scene.from_arrays(sd)
    sc.set_traversal(exact)
    fr = sc.frame(W, H)
    gi = rb.ReSTIRIndirect(fr)
    gi.set_bounce_walk(bounce_exact)
    if staged is not None:
        gi.set_pipeline(staged)
    base = rb.Camera.from_scene(sd)
    out = []
    for f in range(frames):
        cam = base.orbit(f) if orbit else base
        fr.gbuffer_render(cam)
        gi.restir_indirect(cam, f, f if accumulate else 0, max_depth, reuse)
        r = gi.read_reservoirs()
        r17 = np.concatenate([r["Lo"], r["xv"], r["nv"], r["xs"], r["ns"], r["numSamples"].astype(np.float32)[:, None], r["weight"][:, None]], axis=1)
        out.append({"indirect": gi.read(), "reservoir": np.ascontiguousarray(r17, np.float32)})
        fr.gbuffer_update(cam)
    fallback = gi.fallback_pixels()
    gi.close()
    fr.close()
    if scene is None:
        sc.close()
    return out, fallback


def gi_agreement(got, want, tol=1e-3):
    """Per-frame agreement statistics between two GI runs whose BSDF sampling went through different libm's (sinf / cosf)."""
    ia, ib = got["indirect"], want["indirect"]
    scale = np.maximum(np.abs(ia), np.abs(ib)).max(1)
    err = np.abs(ia - ib).max(1)
    ok = err <= tol * scale + 1e-7
    ra, rb_ = got["reservoir"], want["reservoir"]
    lit = scale > 0
    return {
        "bit_identical": float((ia.view(np.uint32) == ib.view(np.uint32)).all(1).mean()),
        "within_tol": float(ok.mean()),
        "within_tol_lit": float(ok[lit].mean()) if lit.any() else 1.0,
        "lit": float(lit.mean()),
        "mean_got": float(ia.mean()), "mean_want": float(ib.mean()),
        "M_equal": float((ra[:, 15] == rb_[:, 15]).mean()),
        "xv_nv_bit_identical": float((ra[:, 3:9].view(np.uint32) == rb_[:, 3:9].view(np.uint32)).all(1).mean()),
        "xs_close": float((np.abs(ra[:, 9:12] - rb_[:, 9:12]).max(1) <= 1e-4 * (1.0 + np.abs(rb_[:, 9:12]).max(1))).mean()),
        "weight_close": float((np.abs(ra[:, 16] - rb_[:, 16]) <= tol * np.maximum(np.abs(ra[:, 16]), np.abs(rb_[:, 16])) + 1e-7).mean()),
    }
