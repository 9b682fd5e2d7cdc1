"""Fuzz: traced-tree walk (default) vs reference-order walk (validation mode) on random scenes / cameras; every buffer must
match bit for bit.  Also a few against the CPU oracle.  (development aid; the judged checks live in tests/)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import restir_b200 as rb
from restir_b200 import scenes
import helpers
rb.init(0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
bad_total = 0
for it in range(n):
    T = int(rng.integers(2000, 60000)); L = int(rng.integers(1, max(2, T // 10)))
    W, H = int(rng.integers(200, 700)), int(rng.integers(150, 400))
    sd = scenes.procedural(int(rng.integers(1, 10**6)), T, L, (W, H)) if it % 4 else scenes.cornell_box((W, H), metal_tall_box=bool(it & 4))
    if it % 4:
        sd.eye = (float(rng.uniform(-8, 8)), float(rng.uniform(0.5, 8)), float(rng.uniform(-8, 16)))
        sd.rotation = (float(rng.uniform(-180, 180)), float(rng.uniform(-60, 10)), 0.0)
        sd.fovy = float(rng.uniform(10, 40))
    else:
        sd.eye = (float(rng.uniform(-0.6, 0.6)), float(rng.uniform(0.3, 1.7)), float(rng.uniform(1.5, 4.5)))
        sd.rotation = (float(rng.uniform(-110, -70)), float(rng.uniform(-20, 20)), 0.0)
    reuse = int(rng.integers(0, 4)); radius = float(rng.choice([5.0, 12.0, 30.0])); k = int(rng.integers(1, 9))
    sc = rb.Scene.from_arrays(sd)
    a, _ = helpers.run_gpu(rb, sd, 3, reuse, radius=radius, k=k, light_index=True, scene=sc, exact=False)
    b, _ = helpers.run_gpu(rb, sd, 3, reuse, radius=radius, k=k, light_index=True, scene=sc, exact=True)
    bad = 0
    for f in range(3):
        for nme in b[f]:
            bad += helpers.mismatches(a[f][nme], b[f][nme])
    fb = sc.fallback_rays()
    print("case %2d %-28s %4dx%-4d reuse %d r %4.1f k %d : mismatching pixel-buffers %d  fix-up pixels %s" % (it, sd.name, W, H, reuse, radius, k, bad, fb), flush=True)
    bad_total += bad
    sc.close()
print("TOTAL", bad_total)
