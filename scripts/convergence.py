"""BASELINE config 5: accumulated ReSTIR DI (static camera, accumulate=true) vs the accumulated one-sample NEE image
(PTDirectKernel, pathtrace.cu:279) -- relMSE after N frames, for k = 1..8 neighbours and 1..3 spatial passes.

    python scripts/convergence.py [workload] [frames]

Note: the reference's RIS target/pdf and its 1/M weights make its estimator biased by design, so
the numbers below describe the reference's algorithm, not an unbiased convergence to the NEE image."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import restir_b200 as rb
from bench import WORKLOADS, make_scene

work = sys.argv[1] if len(sys.argv) > 1 else "config3"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
desc, spec, res, _, radius = WORKLOADS[work]
rb.init(0)
sd = make_scene(spec, res)
sc = rb.Scene.from_arrays(sd)
cam = rb.Camera.from_scene(sd)
W, H = res


def accumulate(fn):
    fr = sc.frame(W, H)
    fr.gbuffer_render(cam)
    for it in range(frames):
        fn(fr, it)
    img = fr.read("radiance").astype(np.float64)
    fr.close()
    return img


ref = accumulate(lambda fr, it: fr.pathtrace_direct(cam, 100000 + it, it))


def relmse(a, b):
    return float(np.mean(((a - b) ** 2).sum(1) / (b.sum(1) ** 2 + 1e-2)))


out = {"workload": work, "frames": frames, "resolution": [W, H], "mean_ptdirect": float(ref.mean()), "runs": []}
for reuse, k, passes in [(0, 5, 1), (1, 5, 1), (3, 1, 1), (3, 5, 1), (3, 8, 1), (3, 5, 2), (3, 5, 3)]:
    prm = rb.default_params(reuse=reuse, radius=radius, k=k, passes=passes)

    def step(fr, it, prm=prm):
        fr.restir_direct(cam, prm, it, it)      # static camera: the G-buffer of frame 0 stays valid; motion = identity after the first update
        if it == 0:
            fr.gbuffer_update(cam); fr.gbuffer_render(cam)
    img = accumulate(step)
    out["runs"].append({"reuse": reuse, "k": k, "passes": passes, "mean": float(img.mean()), "relMSE_vs_ptdirect": relmse(img, ref)})
print(json.dumps(out))
