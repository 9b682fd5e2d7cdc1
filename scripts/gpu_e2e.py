import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import restir_b200 as rb
from restir_b200 import scenes
rb.init(0)
sd = scenes.cornell_box((1920, 1080)); sc = rb.Scene.from_arrays(sd); fr = sc.frame(1920, 1080)
base = rb.Camera.from_scene(sd); prm = rb.default_params(reuse=3, radius=30.0)
outs = [rb.pinned_empty(1920*1080*4), rb.pinned_empty(1920*1080*4)]
cams = [base.orbit(k) for k in range(1400)]
N = 100
def t(f):
    fr.sync(); t0 = time.perf_counter(); f(); fr.sync(); return (time.perf_counter() - t0) / N * 1e3
k = [0]
def device_only():
    for i in range(N):
        c = cams[k[0]]; fr.gbuffer_render(c); fr.restir_direct(c, prm, k[0], 0); fr.gbuffer_update(c); k[0] += 1
def sync_call():
    for i in range(N):
        fr.render_frame_host(cams[k[0]], prm, k[0], 0, 2, outs[0]); k[0] += 1
def async_wait_prev():
    for i in range(N):
        fr.render_frame_host_async(cams[k[0]], prm, k[0], 0, 2, outs[i & 1], i & 1); k[0] += 1
        if i: fr.wait_host((i - 1) & 1)
    fr.wait_host((N - 1) & 1)
def async_nowait():
    for i in range(N):
        fr.render_frame_host_async(cams[k[0]], prm, k[0], 0, 2, outs[i & 1], i & 1); k[0] += 1
    fr.wait_host(0); fr.wait_host(1)
def launch_only_cost():
    t0 = time.perf_counter()
    for i in range(N):
        c = cams[k[0]]; fr.gbuffer_render(c); fr.restir_direct(c, prm, k[0], 0); fr.gbuffer_update(c); k[0] += 1
    return (time.perf_counter() - t0) / N * 1e3
for name, f in (("device_only", device_only), ("sync_call", sync_call), ("async_wait_prev", async_wait_prev), ("async_nowait", async_nowait), ("device_only", device_only)):
    for rep in range(2):
        print(name, "%.3f ms/frame" % t(f))
fr.sync(); print("cpu launch cost per frame %.3f ms" % launch_only_cost())
import numpy as np, ctypes
t0 = time.perf_counter()
for i in range(50): fr.read("ldr")
print("frame_read ldr (pageable) %.3f ms" % ((time.perf_counter() - t0) / 50 * 1e3))
