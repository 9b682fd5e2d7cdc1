"""Quick GPU-vs-oracle comparison (development aid; the judged checks live in tests/)."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import restir_b200 as rb
from oracle.oracle import Oracle, default_params as oparams, make_camera, orbit_camera
from restir_b200 import scenes


def diff_report(name, a, b):
    av = a.view(np.uint8).reshape(a.shape[0], -1)
    bv = b.view(np.uint8).reshape(b.shape[0], -1)
    bad = (av != bv).any(1)
    nd = int(bad.sum())
    extra = ""
    if nd and a.dtype.fields is None and a.dtype.kind == "f":
        af, bf = a.reshape(a.shape[0], -1).astype(np.float64), b.reshape(b.shape[0], -1).astype(np.float64)
        rel = np.abs(af - bf) / np.maximum(np.abs(bf), 1e-20)
        extra = " max_rel=%.3g first=%d" % (rel.max(), int(np.nonzero(bad)[0][0]))
    elif nd:
        extra = " first=%d" % int(np.nonzero(bad)[0][0])
    print("    %-15s mismatching pixels: %d / %d%s" % (name, nd, a.shape[0], extra))
    return nd


def run(sd, frames, reuse, radius=5.0, kind="port"):
    rb.init(0)
    orc = Oracle(kind)
    W, H = sd.resolution
    t0 = time.time()
    sc = rb.Scene.from_arrays(sd)
    so = orc.scene(sd)
    print("scene %s: T=%d L=%d depth=%d host build %.3fs (oracle+product %.2fs)" % (sd.name, sd.num_tris, sc.info.numLights, sc.info.bvhDepth, sc.info.buildSeconds, time.time() - t0))
    fr = sc.frame(W, H)
    fo = so.frame(W, H)
    base = rb.Camera.from_scene(sd)
    obase = make_camera(sd)
    orc.lib.orc_camera_update(C.byref(obase))
    prm, oprm = rb.default_params(reuse=reuse, radius=radius), oparams(reuse=reuse, radius=radius)
    total = 0
    for k in range(frames):
        cam, ocam = base.orbit(k), orbit_camera(orc, obase, k)
        assert bytes(cam) [:120] == bytes(ocam)[:120]
        fr.gbuffer_render(cam)
        fr.restir_direct(cam, prm, k, 0)
        ms = fr.stage_ms()
        t1 = time.time()
        fo.gbuffer_render(ocam)
        fo.restir_direct(ocam, oprm, k, 0)
        t2 = time.time()
        print("  frame %d gpu ms %s | oracle %.2fs (%d threads)" % (k, {a: round(b, 3) for a, b in ms.items() if b}, t2 - t1, orc.threads()))
        for n in ("matid", "motion", "depth", "normal", "albedo", "light_index", "reservoir", "radiance") + (("reservoir_temp",) if reuse & 2 else ()):
            if kind == "reference" and n == "light_index":
                continue
            total += diff_report(n, fr.read(n), fo.buffer(n))
        fr.gbuffer_update(cam)
        fo.gbuffer_update(ocam)
    print("  halo_miss", fr.halo_miss(), "launches", rb.launch_count())
    # PTDirect
    fr.pathtrace_direct(base, 100, 0)
    fo.pathtrace_direct(obase, 100, 0)
    print("  ptdirect ms", fr.stage_ms()["ptdirect"])
    total += diff_report("ptdirect", fr.read("radiance"), fo.buffer("radiance"))
    return total


if __name__ == "__main__":
    small = "--big" not in sys.argv
    bad = 0
    bad += run(scenes.cornell_box((320, 240) if small else (800, 800), metal_tall_box=True), 4, 3)
    bad += run(scenes.cornell_box((320, 240)), 2, 0)
    bad += run(scenes.procedural(1, 20000, 1000, (320, 240) if small else (1280, 720)), 4, 3, radius=30.0)
    print("TOTAL mismatching pixel-buffers:", bad)
