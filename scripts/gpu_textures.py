"""GPU vs oracle on textured / environment-mapped scenes: per buffer, pixels whose bytes differ and the largest
relative difference (procedural pattern and environment lookups use sinf / atan2f: libdevice vs glibc)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import restir_b200 as rb
from helpers import ALL_BUFS, mismatches, run_gpu, run_oracle
from oracle import oracle as O
from restir_b200 import scenes

rb.init(0)
orc = O.Oracle("port")
cases = {
    "cornell_tex": scenes.with_textures(scenes.cornell_box((160, 120)), env=True),
    "cornell_tex_noenv": scenes.with_textures(scenes.cornell_box((160, 120)), env=False),
    "gen3000_tex": scenes.with_textures(scenes.procedural(3, 3000, 200, (160, 90)), env=True),
    "gen3000_envonly": scenes.with_textures(scenes.procedural(3, 3000, 200, (160, 90)), env=True),
}
cases["gen3000_envonly"].materials = scenes.procedural(3, 3000, 200, (160, 90)).materials   # env map only, no maps
res = {}
for name, sd in cases.items():
    for reuse in (0, 3):
        for exact in (False, True):
            want = run_oracle(orc, sd, 3, reuse, radius=30.0, light_index=True)
            got, _ = run_gpu(rb, sd, 3, reuse, radius=30.0, light_index=True, exact=exact)
            r = {}
            for f in range(3):
                for n in want[f]:
                    bad = mismatches(got[f][n], want[f][n])
                    if bad:
                        a = np.ascontiguousarray(got[f][n]); b = np.ascontiguousarray(want[f][n])
                        if a.dtype.names is None and a.dtype == np.float32:
                            rel = float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-6)))
                        else:
                            rel = None
                        r["f%d_%s" % (f, n)] = [bad, rel]
            res["%s_reuse%d_%s" % (name, reuse, "exact" if exact else "fast")] = r
            print(name, reuse, exact, r, flush=True)
# PTDirect
for name, sd in cases.items():
    W, H = sd.resolution
    so = orc.scene(sd); fo = so.frame(W, H)
    import ctypes as C
    base = O.make_camera(sd); orc.lib.orc_camera_update(C.byref(base))
    sc = rb.Scene.from_arrays(sd); fr = sc.frame(W, H); cam = rb.Camera.from_scene(sd)
    for k in range(2):
        fo.pathtrace_direct(base, k, k); fr.pathtrace_direct(cam, k, k)
    a, b = fr.read("radiance"), fo.buffer("radiance")
    bad = mismatches(a, b)
    rel = float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-6)))
    res["%s_ptdirect" % name] = [bad, rel]
    print(name, "ptdirect", bad, rel, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "textures.json"), "w"), indent=1)
