"""Per-kernel cost of a strip vs the full frame on ONE GPU (development aid; run under
`ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none`): the 4K config4 frame as one frame and as
the 8 strips of a measured run, each strip rendered alone (no exchange: edge rows lose their temporal / spatial neighbours,
which does not change the cost picture).  Usage: gpu_strip_probe.py [workload] [b0,b1,...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import restir_b200 as rb
from bench import WORKLOADS, make_scene, orbit_index

rb.init(0)
w = sys.argv[1] if len(sys.argv) > 1 else "config4"
bounds = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 664, 849, 1057, 1252, 1431, 1634, 1872, 2160]
desc, spec, res, reuse, radius = WORKLOADS[w]
sd = make_scene(spec, res)
sc = rb.Scene.from_arrays(sd)
base = rb.Camera.from_scene(sd)
prm = rb.default_params(reuse=reuse, radius=radius)
FRAMES = 4


def run(fr, tag):
    for k in range(FRAMES):
        cam = base.orbit(orbit_index(k))
        fr.gbuffer_render(cam); fr.restir_direct(cam, prm, k, 0); fr.gbuffer_update(cam)
    fr.sync()
    print(tag, fr.stage_ms(), flush=True)


fr = sc.frame(*res)
run(fr, "full")
fr.close()
for i in range(len(bounds) - 1):
    fr = sc.frame(res[0], res[1], rows=(bounds[i], bounds[i + 1]), halo=31)
    fr.set_halo_render(False)
    run(fr, "strip %d rows %d-%d" % (i, bounds[i], bounds[i + 1]))
    fr.close()
