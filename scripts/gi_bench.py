"""Device-timed ReSTIR GI (rstr_restir_indirect) on the bench workloads' scenes: ms / frame of the GI call alone
(CUDA events on the frame's stream around `steps` frames of an orbit; the G-buffer render of each frame is outside the events),
per traversal mode, with the fix-up pixel count.  One JSON line per case to stdout and to --out.

    python scripts/gi_bench.py [--workloads config2 config3] [--steps 20] [--depth 3] [--out gpurun_out/gi_bench.jsonl]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


_SCENES = {}


def run(rb, name, steps, warmup, depth, exact, bounce_exact, staged=False):
    desc, spec, res, _, _ = bench.WORKLOADS[name]
    t0 = time.time()
    if name not in _SCENES:                      # one scene per workload for all modes (GPU seconds are scarce)
        sd = bench.make_scene(spec, res)
        _SCENES.clear()
        _SCENES[name] = (sd, rb.Scene.from_arrays(sd))
    sd, sc = _SCENES[name]
    sc.set_traversal(exact)
    fr = sc.frame(*res)
    gi = rb.ReSTIRIndirect(fr)
    gi.set_bounce_walk(bounce_exact)
    gi.set_pipeline(int(staged))
    base = rb.Camera.from_scene(sd)
    setup_s = time.time() - t0
    k = 0
    total = 0.0
    for i in range(warmup + steps):
        cam = base.orbit(bench.orbit_index(k))
        fr.gbuffer_render(cam)
        fr.sync()
        fr.mark(0)
        gi.restir_indirect(cam, k, 0, depth, 1)
        fr.mark(1)
        fr.sync()
        if i >= warmup:
            total += fr.elapsed_ms(0, 1)
        fr.gbuffer_update(cam)
        k += 1
    img = gi.read()
    out = {"workload": name, "resolution": list(res), "triangles": sc.info.numTris, "emissive_triangles": sc.info.numLights, "trace_depth": depth,
           "traversal": "reference-order walk" if exact else ("packet primary + reference-order bounces" if bounce_exact else "traced tree"),
           "pipeline": "one kernel" if exact or not staged else ("staged (primary / bounce per depth / resolve)" if staged == 1 else "ray queues (primary / head, shadow walker, closest walker, tail per depth / resolve)"),
           "steps": steps, "warmup": warmup, "gi_ms_per_frame": total / steps, "mpixel_per_s": res[0] * res[1] / (total / steps * 1e-3) / 1e6,
           "fixup_pixels_per_frame": gi.fallback_pixels() / (warmup + steps), "mean_indirect": float(img.mean()), "lit_fraction": float((img.sum(1) > 0).mean()),
           "setup_s": setup_s, "build_id": rb.api.build_id()}
    gi.close()
    fr.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", nargs="+", default=["config2", "config3"])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--depth", type=int, default=3)
    ap.add_argument("--modes", nargs="+", default=["traced", "exact"], choices=["traced", "staged", "queued", "mixed", "exact"])
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "gi_bench.jsonl"))
    args = ap.parse_args()
    import restir_b200 as rb

    rb.init(0)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "a") as f:
        for name in args.workloads:
            for mode in args.modes:
                r = run(rb, name, args.steps if mode != "exact" else max(3, args.steps // 4), args.warmup, args.depth, mode == "exact", mode == "mixed", {"staged": 1, "queued": 2}.get(mode, 0))
                line = json.dumps(r)
                print(line, flush=True)
                f.write(line + "\n")
                f.flush()


if __name__ == "__main__":
    main()
