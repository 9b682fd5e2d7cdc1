"""Dump the traced tree + camera of a bench workload for scripts/travsim.cpp (development aid, CPU only)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import restir_b200 as rb
import bench

work, out = sys.argv[1], sys.argv[2]
os.makedirs(out, exist_ok=True)
desc, spec, res, reuse, radius = bench.WORKLOADS[work]
sd = bench.make_scene(spec, res)
sc = rb.Scene.from_arrays(sd)
sc.read("traced_nodes").tofile(os.path.join(out, "nodes.bin"))
sc.read("traced_tris").tofile(os.path.join(out, "tris.bin"))
cam = rb.Camera.from_scene(sd).orbit(int(sys.argv[3]) if len(sys.argv) > 3 else 10)
tan = float(np.tan(np.radians(np.float32(cam.fov[1]))))
v = list(cam.position) + list(cam.right) + list(cam.up) + list(cam.view) + [res[0] / res[1], tan, float(sc.info.tracedRoot)]
np.asarray(v, np.float32).tofile(os.path.join(out, "cam.bin"))
print(work, res, "root", sc.info.tracedRoot, "nodes", sc.info.tracedNodes)
