"""CPU, build container only: the oracle restatement vs the reference's own code (oracle/_ref, compiled from
/root/reference by oracle/Makefile) on inputs beyond the committed fixtures.  Skipped where _ref is absent."""
import numpy as np
import pytest

import helpers
from restir_b200 import scenes


@pytest.mark.parametrize("seed,T,L", [(3, 1500, 60), (9, 6000, 700)])
def test_host_build_random_scenes(port_oracle, ref_oracle, seed, T, L):
    sd = scenes.procedural(seed, T, L, (40, 30))
    a, b = port_oracle.scene(sd), ref_oracle.scene(sd)
    assert np.array_equal(a.boxes().view(np.uint32), b.boxes().view(np.uint32))
    for i in range(6):
        assert np.array_equal(a.mtbvh(i), b.mtbvh(i))
    assert np.array_equal(a.alias_table().view(np.uint8), b.alias_table().view(np.uint8))
    assert a.sum_light_power() == b.sum_light_power()


@pytest.mark.parametrize("reuse,radius,k,cap", [(3, 12.0, 8, 20), (3, 5.0, 1, 4), (1, 5.0, 5, 2)])
def test_frames_other_knobs(port_oracle, ref_oracle, reuse, radius, k, cap):
    sd = scenes.procedural(5, 3000, 150, (56, 40))
    a = helpers.run_oracle(port_oracle, sd, 4, reuse, radius=radius, k=k, cap=cap, accumulate=True)
    b = helpers.run_oracle(ref_oracle, sd, 4, reuse, radius=radius, k=k, cap=cap, accumulate=True)
    helpers.assert_frames_equal(a, b, "port vs reference")


@pytest.mark.parametrize("passes", [2, 3])
def test_multi_pass_spatial(port_oracle, ref_oracle, passes):
    sd = scenes.procedural(7, 3000, 150, (56, 40))
    a = helpers.run_oracle(port_oracle, sd, 3, 3, radius=8.0, passes=passes)
    b = helpers.run_oracle(ref_oracle, sd, 3, 3, radius=8.0, passes=passes)
    helpers.assert_frames_equal(a, b, "port vs reference, %d passes" % passes)


def test_probe_rays(port_oracle, ref_oracle):
    sd = scenes.cornell_box((8, 8), metal_tall_box=True)
    a, b = port_oracle.scene(sd), ref_oracle.scene(sd)
    rng = np.random.default_rng(0)
    for _ in range(300):
        o = rng.uniform(-0.9, 0.9, 3) + np.array([0, 1, 0])
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        pa, oa, ma = a.intersect(o, d)
        pb, ob, mb = b.intersect(o, d)
        assert pa == pb and (pa < 0 or (np.array_equal(oa.view(np.uint32), ob.view(np.uint32)) and ma == mb))
        y = rng.uniform(-0.9, 0.9, 3) + np.array([0, 1, 0])
        assert a.occluded(o, y) == b.occluded(o, y)
    # axis-aligned and vanishing-component rays exercise the special cases of bvh.h:91-147
    for d in ([1, 0, 0], [0, -1, 0], [0, 0, -1], [0, 1e-7, -1], [1e-7, 0.6, -0.8], [0.6, 1e-8, 0.8]):
        d = np.asarray(d, np.float32)
        d = d / np.linalg.norm(d)
        for o in ([0, 1, 0.5], [0.3, 0.6, 0.9], [-0.35, 1.9, -0.3]):
            pa, oa, ma = a.intersect(o, d)
            pb, ob, mb = b.intersect(o, d)
            assert pa == pb


@pytest.mark.parametrize("name", ["cornell_tex", "gen2000_tex"])
def test_textured_port_equals_reference(port_oracle, ref_oracle, name):
    """Texture / environment-map branches: restatement == the reference's own functions at a larger size than the
    committed fixture."""
    import dataclasses
    sd = dataclasses.replace(helpers.textured_scenes()[name], resolution=(128, 96))
    a = helpers.run_oracle(port_oracle, sd, 3, 3, radius=30.0)
    b = helpers.run_oracle(ref_oracle, sd, 3, 3, radius=30.0)
    helpers.assert_frames_equal(a, b, name)


@pytest.mark.parametrize("depth,reuse", [(1, 1), (4, 1), (6, 0)])
def test_restir_indirect_port_equals_reference(port_oracle, ref_oracle, depth, reuse):
    """ReSTIRIndirect at other trace depths and a larger image than the committed fixture."""
    import dataclasses
    for name in ("cornell_glass", "cornell_tex"):
        sd = dataclasses.replace(helpers.gi_scenes()[name], resolution=(96, 72))
        a = helpers.run_oracle_gi(port_oracle, sd, 3, depth, reuse, accumulate=True)
        b = helpers.run_oracle_gi(ref_oracle, sd, 3, depth, reuse, accumulate=True)
        helpers.assert_frames_equal(a, b, "GI %s depth %d" % (name, depth))
