"""Shadow-ray statistics of one frame (development aid; needs the RS_SHADOW_STATS build:
RSTR_LIBNAME=librestir_b200_stats.so RSTR_DEFINES=-DRS_SHADOW_STATS python scripts/gpu_shadow_stats.py [workload] [rowLo rowHi])."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import restir_b200 as rb
from bench import WORKLOADS, make_scene, orbit_index

rb.init(0)
w = sys.argv[1] if len(sys.argv) > 1 else "config4_1080p"
rows = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else None
desc, spec, res, reuse, radius = WORKLOADS[w]
sd = make_scene(spec, res)
sc = rb.Scene.from_arrays(sd)
fr = sc.frame(res[0], res[1], rows=rows, halo=31 if rows else 0)
if rows:
    fr.set_halo_render(False)
base = rb.Camera.from_scene(sd)
prm = rb.default_params(reuse=reuse, radius=radius)
L = rb.lib()
out = (C.c_ulonglong * 64)()
if not hasattr(L, "rstr_debug_shadow_stats"):     # a plain build: per-stage device times only, averaged over 12 frames
    acc = []
    for k in range(16):
        cam = base.orbit(orbit_index(k))
        fr.gbuffer_render(cam); fr.restir_direct(cam, prm, k, 0); fr.gbuffer_update(cam)
        if k >= 4:
            acc.append(fr.stage_ms())
    print(w, rows, rb.api.build_id(), "stage ms (12 frames)", {n: round(sum(a[n] for a in acc) / len(acc), 4) for n in ("ris", "spatial")})
    sys.exit(0)
for k in range(6):
    if k == 5:
        fr.sync(); L.rstr_debug_shadow_stats(None, 1)
    cam = base.orbit(orbit_index(k))
    fr.gbuffer_render(cam); fr.restir_direct(cam, prm, k, 0); fr.gbuffer_update(cam)
fr.sync()
L.rstr_debug_shadow_stats(out, 0)
v = list(out)
print(w, rows, "rays", v[32], "steps/ray %.1f" % (v[33] / max(v[32], 1)), "longest", v[34])
print("  log2 histogram of steps:", {("%d-%d" % (1 << b if b else 0, (2 << b) - 1)): v[b] for b in range(32) if v[b]})
print("  kernel %.1f us, first warp out of work after %.1f us -> tail %.1f us" % ((v[41] - v[42]) / 1e3, (v[40] - v[42]) / 1e3, (v[41] - v[40]) / 1e3))
print("  stage ms", fr.stage_ms())
