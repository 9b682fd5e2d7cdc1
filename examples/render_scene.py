"""Render a scene with the ReSTIR DI pipeline and save a tone-mapped PNG (needs a B200; no CPU fallback).

    python examples/render_scene.py                       # built-in Cornell box, 64 accumulated frames -> cornell.png
    python examples/render_scene.py scene.txt out.png 128 # a scene file in the reference's grammar ("-" = built-in scene)

The loop is runCuda() of the reference (main.cpp:146-185): GBuffer::render, ReSTIRDirect, copy-out, GBuffer::update,
with `iter` counting up so that the radiance image is the running mean (restir.cu:230)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import restir_b200 as rb
from restir_b200 import scenes


def main():
    rb.init(0)
    if len(sys.argv) > 1 and sys.argv[1] not in ("", "-"):
        scene = rb.Scene.from_file(sys.argv[1])
        cam = scene.camera
    else:
        sd = scenes.cornell_box((1280, 720), metal_tall_box=True)
        scene = rb.Scene.from_arrays(sd)
        cam = rb.Camera.from_scene(sd)
    out = sys.argv[2] if len(sys.argv) > 2 else "cornell.png"
    frames = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    w, h = cam.resolution[0], cam.resolution[1]
    frame = scene.frame(w, h)
    params = rb.default_params(reuse=rb.REUSE_SPATIOTEMPORAL, radius=30.0)
    for k in range(frames):
        frame.gbuffer_render(cam)
        frame.restir_direct(cam, params, looper=k, it=k)
        frame.gbuffer_update(cam)
    frame.save_png(out, rb.TONEMAP_ACES)                 # saveImage: mirrored like the reference's (main.cpp:126)
    print("wrote", out, "(%dx%d, %d frames, %.2f ms per frame on the device)" % (w, h, frames, sum(frame.stage_ms().values())))
    frame.close()
    scene.close()


if __name__ == "__main__":
    main()
