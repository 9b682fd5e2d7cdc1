import json, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import restir_b200 as rb
from restir_b200 import scenes
from scripts.ref_cuda_compare import run_ref
W, H = 640, 360
sd = scenes.cornell_box((W, H))
tmp = tempfile.mkdtemp()
txt = scenes.write_scene_files(sd, tmp, "scene")
rb.init(0)
sc = rb.Scene.from_file(txt)
base = sc.camera
for reuse, nframes in ((0, 1), (1, 2)):
    pre = os.path.join(tmp, "ref%d_" % reuse)
    run_ref("ref_headless_r5", txt, nframes, 0, reuse, pre, nframes - 1)
    fr = sc.frame(W, H)
    prm = rb.default_params(reuse=reuse)
    for k in range(nframes):
        cam = base.orbit(k); fr.gbuffer_render(cam); fr.restir_direct(cam, prm, k, 0)
        if k < nframes - 1: fr.gbuffer_update(cam)
    ref = np.fromfile(pre + "radiance.bin", np.float32).reshape(-1, 3)
    mine = fr.read("radiance")
    mat = fr.read("matid")
    res = fr.read("reservoir")
    shaded = mat >= 0
    same = (ref == mine).all(1)
    print("reuse", reuse, "frames", nframes, "shaded frac", shaded.mean(), "bitexact all", same.mean(), "bitexact among shaded", same[shaded].mean(), "among unshaded", same[~shaded].mean())
    dif = np.nonzero(~same & shaded)[0]
    ratio = ref[dif].sum(1) / np.maximum(mine[dif].sum(1), 1e-9)
    print("  differing shaded:", len(dif), "ratio ref/mine percentiles", np.percentile(ratio, [1, 10, 50, 90, 99]))
    print("  mean ref", ref.mean(), "mean mine", mine.mean(), " mine zero frac among differing", (mine[dif].sum(1) == 0).mean(), "ref zero frac", (ref[dif].sum(1) == 0).mean())
    rr = np.fromfile(pre + "reservoir.bin", rb.api.RESERVOIR_DTYPE)
    print("  reservoir: M equal frac", (rr["M"] == res["M"]).mean(), "w bitexact frac", (rr["w"] == res["w"]).mean(), "wi bitexact", (rr["wi"] == res["wi"]).all(1).mean(), "Li equal", (rr["Li"] == res["Li"]).all(1).mean())
    for i in dif[:3]:
        print("   ref resv", rr[i], "mine", res[i])
    for i in dif[:8]:
        print("   px", i % W, i // W, "ref", ref[i], "mine", mine[i], "M", res["M"][i], "w", res["w"][i], "dist", res["dist"][i])
