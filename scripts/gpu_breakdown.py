"""Where does the RIS kernel's time go?  (development aid)  Uses the knobs only: candidates 0 removes the candidate
loop and the shadow ray (weight stays 0 -> the ray is skipped)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import restir_b200 as rb
from bench import WORKLOADS, make_scene

rb.init(0)
for w in sys.argv[1:] or ["config3"]:
    desc, spec, res, reuse, radius = WORKLOADS[w]
    sd = make_scene(spec, res)
    sc = rb.Scene.from_arrays(sd)
    fr = sc.frame(*res)
    base = rb.Camera.from_scene(sd)
    out = {}
    for name, cands, ru in (("c32_st", 32, 3), ("c0_st", 0, 3), ("c32_ris", 32, 0), ("c1_ris", 1, 0), ("c8_ris", 8, 0), ("c16_ris", 16, 0)):
        prm = rb.default_params(reuse=ru, radius=radius, candidates=cands)
        acc = []
        for k in range(8):
            cam = base.orbit(k)
            fr.gbuffer_render(cam); fr.restir_direct(cam, prm, k, 0); fr.gbuffer_update(cam)
            if k >= 3:
                acc.append(fr.stage_ms())
        out[name] = {s: round(float(np.mean([a[s] for a in acc])), 3) for s in ("gbuffer", "ris", "spatial")}
    print(w, out, "fallback rays (all runs)", sc.fallback_rays())
    shaded = (fr.read("matid") >= 0).mean()
    print("  shaded fraction", shaded)
