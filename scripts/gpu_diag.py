import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import restir_b200 as rb
from restir_b200 import scenes
import helpers
rb.init(0)
def diag(sd, orbit, frames=1, reuse=0):
    W,H = sd.resolution
    a,_ = helpers.run_gpu(rb, sd, frames, reuse, orbit=orbit, exact=False)
    b,_ = helpers.run_gpu(rb, sd, frames, reuse, orbit=orbit, exact=True)
    for f in range(frames):
        for n in ("matid","depth","albedo","normal","radiance"):
            av = np.ascontiguousarray(a[f][n]).view(np.uint8).reshape(W*H,-1); bv = np.ascontiguousarray(b[f][n]).view(np.uint8).reshape(W*H,-1)
            bad = np.nonzero((av!=bv).any(1))[0]
            if len(bad):
                print(sd.name, W,H, "frame", f, n, "differs at", len(bad), "pixels")
                for i in bad[:6]:
                    print("   px", i%W, i//W, "fast", a[f][n][i], a[f]["matid"][i], a[f]["depth"][i], "exact", b[f][n][i], b[f]["matid"][i], b[f]["depth"][i])
diag(scenes.cornell_box((33,33)), False)
diag(scenes.cornell_box((1920,1080), metal_tall_box=True), True)
diag(scenes.procedural(1,200000,10000,(1920,1080)), True)
