import sys, tempfile, ctypes as C, subprocess, os
sys.path.insert(0,'/root/repo')
import numpy as np
from oracle.oracle import Oracle
from restir_b200 import scenes
sd = scenes.cornell_box((64,64)); tmp = tempfile.mkdtemp(); txt = scenes.write_scene_files(sd, tmp, "scene")
ref = Oracle("reference")
rs = ref.lib.ref_scene_load_file(txt.encode())
out = np.zeros(16, np.float32); ref.lib.ref_probe.argtypes=[C.c_void_p, C.c_void_p]; ref.lib.ref_probe(rs, out.ctypes.data)
print("cpu probe p=%g Li.x=%g wi.y=%g dist=%g sumInv=%g L=%g w=%g g.x=%g bsdf=%g cos=%g sizeof=%g type=%g" % tuple(out[:12]))
if True:
    r = subprocess.run(["/root/repo/oracle/_ref/ref_headless_r5", txt, "1", "0", "0"], capture_output=True, text=True, env=dict(os.environ, REF_PROBE="1"))
    print(r.stdout, r.stderr[-500:])
