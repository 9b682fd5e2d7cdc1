"""CPU: the oracle restatement against the committed golden fixtures (generated from the reference's own code by
tests/golden/make_golden.py) and against the known-answer vectors of SURVEY.md 8c / App. D."""
import os

import numpy as np
import pytest

import helpers

G = helpers.GOLDEN


def test_rng_known_answers(port_oracle):
    # nvcc constant-folded PTX for seed (iter=7, index=12345)
    want = np.array([0x3E9C4ABF, 0x3DC90DD9, 0x3F4F3A55, 0x3F182607, 0x3F6A7184], np.uint32)
    assert np.array_equal(port_oracle.rng_draws(7, 12345, 5).view(np.uint32), want)
    ka = np.load(os.path.join(G, "known_answers.npz"))
    for key in ka.files:
        if key.startswith("rng_"):
            l, i = key[4:].split("_")
            got = port_oracle.rng_draws(int(l[1:]), int(i[1:]), 8)
            assert np.array_equal(got.view(np.uint32), ka[key].view(np.uint32)), key


def test_alias_known_answers(port_oracle):
    # powers {1,2,3,10} -> probs {.25,.5,.75,1}, failId 3, sum 16
    t, s = port_oracle.alias_build([1, 2, 3, 10])
    assert s == 16.0 and np.array_equal(t["prob"], np.array([0.25, 0.5, 0.75, 1.0], np.float32)) and np.all(t["failId"] == 3)
    ka = np.load(os.path.join(G, "known_answers.npz"))
    t2, s2 = port_oracle.alias_build(np.linspace(0.1, 7.3, 37) ** 2)
    assert np.array_equal(t2["prob"].view(np.uint32), ka["alias2_prob"].view(np.uint32))
    assert np.array_equal(t2["failId"], ka["alias2_fail"]) and s2 == float(ka["alias2_sum"])
    # every alias table is a distribution: total mass n
    mass = np.zeros(37)
    for i in range(37):
        mass[i] += t2["prob"][i]
        mass[t2["failId"][i]] += 1.0 - t2["prob"][i]
    w = np.linspace(0.1, 7.3, 37) ** 2
    assert np.allclose(mass / 37.0, w / w.sum(), atol=1e-5)


def test_five_triangle_bvh_probe(port_oracle):
    # BVHSize 9 and ordering-0 node list (prim, box, miss)
    so = port_oracle.scene(helpers.five_triangles())
    assert so.bvh_size == 9
    want = np.array([(-1, 0, 9), (4, 8, 2), (-1, 1, 9), (-1, 3, 8), (-1, 5, 7), (3, 7, 6), (2, 6, 7), (1, 4, 8), (0, 2, 9)], np.int32)
    assert np.array_equal(so.mtbvh(0), want)


@pytest.mark.parametrize("name", ["cornell", "cornell_metal", "gen2000", "five"])
def test_host_build_matches_golden(port_oracle, name):
    sd = helpers.test_scenes()[name]
    g = np.load(os.path.join(G, "host_%s.npz" % name))
    so = port_oracle.scene(sd)
    assert np.array_equal(so.boxes().view(np.uint32), g["boxes"].view(np.uint32))
    for i in range(6):
        assert np.array_equal(so.mtbvh(i), g["mtbvh%d" % i]), i
    assert np.array_equal(so.light_prim_ids(), g["light_prim_ids"])
    assert np.array_equal(so.light_radiance(), g["light_radiance"])
    assert np.array_equal(so.alias_table()["prob"].view(np.uint32), g["alias_prob"].view(np.uint32))
    assert np.array_equal(so.alias_table()["failId"], g["alias_fail"])
    assert so.sum_light_power() == float(g["sum_power"])


MODES = (("ris", 0, 5.0), ("temporal", 1, 5.0), ("spatial", 2, 5.0), ("st", 3, 5.0), ("st_r30", 3, 30.0))


@pytest.mark.parametrize("name", ["cornell", "cornell_metal", "gen2000", "five"])
def test_frames_match_golden(port_oracle, name):
    sd = helpers.test_scenes()[name]
    g = np.load(os.path.join(G, "frames_%s.npz" % name))
    for mode, reuse, radius in MODES:
        frames = helpers.run_oracle(port_oracle, sd, 3, reuse, radius=radius)
        for f, bufs in enumerate(frames):
            for n, a in bufs.items():
                key = "%s_f%d_%s" % (mode, f, n)
                if key not in g.files:
                    continue
                assert helpers.mismatches(a, g[key]) == 0, key


def test_ptdirect_matches_golden(port_oracle):
    import ctypes as C

    from oracle.oracle import make_camera

    g = np.load(os.path.join(G, "ptdirect.npz"))
    for name, sd in helpers.test_scenes().items():
        so = port_oracle.scene(sd)
        fo = so.frame(*sd.resolution)
        cam = make_camera(sd)
        port_oracle.lib.orc_camera_update(C.byref(cam))
        for it in range(2):
            fo.pathtrace_direct(cam, 100 + it, it)
        assert helpers.mismatches(fo.buffer("radiance"), g[name]) == 0, name


def test_restir_converges_to_ptdirect(port_oracle):
    """Statistical pin: accumulated RIS image vs accumulated one-sample NEE image.
    The reference's RIS estimator is biased by design (its pdf, App. C2, and 1/M weights), so only agreement of
    the overall energy within a loose factor and of the lit/unlit structure is asserted."""
    import ctypes as C

    from oracle.oracle import default_params, make_camera

    sd = helpers.test_scenes()["cornell"]
    so = port_oracle.scene(sd)
    cam = make_camera(sd)
    port_oracle.lib.orc_camera_update(C.byref(cam))
    fa, fb = so.frame(*sd.resolution), so.frame(*sd.resolution)
    prm = default_params(reuse=0)
    for it in range(24):
        fa.gbuffer_render(cam)
        fa.restir_direct(cam, prm, it, it)
        fb.gbuffer_render(cam)
        fb.pathtrace_direct(cam, 1000 + it, it)
    a, b = fa.buffer("radiance"), fb.buffer("radiance")
    assert np.isfinite(a).all() and np.isfinite(b).all()
    lit_a, lit_b = a.sum(1) > 1e-3, b.sum(1) > 1e-3
    assert (lit_a == lit_b).mean() > 0.9


@pytest.mark.parametrize("name", ["cornell_tex", "gen2000_tex", "cornell_tex_noenv"])
def test_textured_frames_match_golden(port_oracle, name):
    """Textures, procedural pattern, normal / metallic / roughness maps and the environment map (scene.h:68-99,
    364-375, 400-403; scene.cpp:136-152): the restatement against the reference's own code, bit for bit."""
    sd = helpers.textured_scenes()[name]
    g = np.load(os.path.join(G, "frames_textured.npz"))
    so = port_oracle.scene(sd)
    assert np.array_equal(np.frombuffer(so.alias_table().tobytes(), np.uint8), g[name + "_alias"])
    assert np.array_equal(np.frombuffer(so.env_alias()[0].tobytes(), np.uint8), g[name + "_env_alias"])
    assert so.sum_light_power() == float(g[name + "_sum_power"])
    so.close()
    for mode, reuse, radius in (("ris", 0, 5.0), ("st_r30", 3, 30.0)):
        frames = helpers.run_oracle(port_oracle, sd, 3, reuse, radius=radius)
        for f, bufs in enumerate(frames):
            for n, a in bufs.items():
                key = "%s_%s_f%d_%s" % (name, mode, f, n)
                if key in g.files:
                    assert helpers.mismatches(a, g[key]) == 0, key
    # PTDirect with textured base colour and the environment map as a light (pathtrace.cu:295-302, scene.h:377-392)
    import ctypes as C
    from oracle.oracle import make_camera
    so = port_oracle.scene(sd)
    fo = so.frame(*sd.resolution)
    cam = make_camera(sd)
    port_oracle.lib.orc_camera_update(C.byref(cam))
    for it in range(2):
        fo.pathtrace_direct(cam, 100 + it, it)
    assert helpers.mismatches(fo.buffer("radiance"), g[name + "_ptdirect"]) == 0
    # the cases are not degenerate: textured albedo varies, and the environment map is visible / sampled where open
    alb = frames[0]["albedo"]
    assert len(np.unique(alb.round(3), axis=0)) > 50
    if name == "gen2000_tex":
        L = so.num_lights
        li = helpers.run_oracle(port_oracle, sd, 1, 0, light_index=True)[0]["light_index"]
        assert (li >= L - 1).sum() > 20          # reservoirs holding an environment-map texel
        assert (frames[0]["matid"] == -1).sum() > 100 and frames[0]["radiance"][frames[0]["matid"] == -1].max() > 0


@pytest.mark.parametrize("name", ["cornell_metal", "gen2000"])
def test_multi_pass_frames_match_golden(port_oracle, name):
    """Two and three spatial passes (BASELINE config 5; restir.cu:201-209 with preClampedMerge<4>): restatement against
    the fixture the reference harness produced."""
    sd = helpers.test_scenes()[name]
    g = np.load(os.path.join(G, "frames_multipass.npz"))
    for passes in (2, 3):
        frames = helpers.run_oracle(port_oracle, sd, 3, 3, radius=12.0, passes=passes, want=("radiance", "reservoir", "reservoir_temp"))
        for f, bufs in enumerate(frames):
            for n, a in bufs.items():
                assert helpers.mismatches(a, g["%s_p%d_f%d_%s" % (name, passes, f, n)]) == 0, (passes, f, n)


@pytest.mark.parametrize("name", ["cornell", "cornell_metal", "cornell_glass", "gen2000", "cornell_tex"])
def test_restir_indirect_matches_golden(port_oracle, name):
    """ReSTIRIndirect (restir.cu:242-416, 448-476): the restatement against the fixture the reference's own
    Material::sample / pdf / BSDF, DevScene::intersect / sampleDirectLight and Reservoir<IndirectLiSample> produced
    (oracle/ref_harness.cpp), bit for bit: lambertian, metallic-workflow (GTR2 visible normals) and dielectric bounces, the
    environment-map miss, temporal reuse over an orbit, and the progressive mean."""
    sd = helpers.gi_scenes()[name]
    g = np.load(os.path.join(G, "gi.npz"))
    frames = helpers.run_oracle_gi(port_oracle, sd, 3, max_depth=3, reuse=1)
    for f, bufs in enumerate(frames):
        assert helpers.mismatches(bufs["indirect"], g["%s_f%d_indirect" % (name, f)]) == 0, f
    assert helpers.mismatches(frames[-1]["reservoir"], g["%s_f2_reservoir" % name]) == 0
    acc = helpers.run_oracle_gi(port_oracle, sd, 2, max_depth=2, reuse=0, accumulate=True, orbit=False)
    assert helpers.mismatches(acc[-1]["indirect"], g["%s_acc_indirect" % name]) == 0
    if name != "gen2000":
        assert (frames[-1]["indirect"].sum(1) > 0).mean() > 0.2          # the fixture is not trivially black
