"""BASELINE config 5 on N GPUs (torchrun -n N scripts/config5_sweep.py [workload] [frames] [b0,b1,...,bN]; N = 1 works without torchrun;
the optional third argument gives the strip cuts, e.g. the ones bench.py's closed-loop refinement found on the same box):
the 4K many-light scene (default workload config4) as image strips through the library's peer data plane,

  * spatial-reuse sweep: k = 1..8 neighbours x 1..3 spatial passes, device-timed ms/frame (max over ranks) of the orbit;
  * convergence: static camera, accumulate = true, `frames` frames of the reference's spatiotemporal ReSTIR (k = 5, 1..3 passes),
    of RIS-only, and -- on one GPU per strip without halo exchange, i.e. only at N = 1 -- of the unbiased mode, against the
    `frames`-frame image of pathTraceDirect (PTDirectKernel, the "reference path-traced image"): mean relMSE over the image.

One JSON line on rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import bench
import restir_b200 as rb
from restir_b200 import strips

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
work = sys.argv[1] if len(sys.argv) > 1 else "config4"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dist = torch = None
if world > 1:
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rb.init(local)
desc, spec, (W, H), reuse, radius = bench.WORKLOADS[work]
sd = bench.make_scene(spec, (W, H))
sc = rb.Scene.from_arrays(sd)
base = rb.Camera.from_scene(sd)
motion = bench.measure_motion_rows(sc, base, W, H, rb) if world > 1 else 0
halo = strips.default_halo(radius, motion) if world > 1 else 0
bounds = strips.uniform_bounds(H, world)
if world > 1 and len(sys.argv) > 3:
    bounds = [int(v) for v in sys.argv[3].split(",")]
    assert len(bounds) == world + 1 and bounds[0] == 0 and bounds[-1] == H
elif world > 1:
    probe = sc.frame(W, H)
    probe.gbuffer_render(base.orbit(0))
    bounds = strips.balanced_bounds(strips.row_cost_from_matid(probe.read("matid"), W), world, min_rows=halo)
    probe.close()
rows = strips.strip_rows(H, world, rank, bounds)


def make():
    fr = sc.frame(W, H, rows=rows, halo=halo)
    grp = None
    if world > 1:
        grp = rb.StripGroup(fr, rank, world)
        t = torch.frombuffer(bytearray(grp.handle()), dtype=torch.uint8).cuda()
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        grp.connect([bytes(o.cpu().numpy().tobytes()) for o in outs])
    return fr, grp


def close(fr, grp):
    fr.sync()
    if world > 1:
        torch.cuda.synchronize(); dist.barrier()
        grp.close()
    fr.close()


def render(fr, grp, cam, prm, looper, it):
    if grp is not None:
        grp.render(cam, prm, looper, it)
    else:
        fr.gbuffer_render(cam); fr.restir_direct(cam, prm, looper, it); fr.gbuffer_update(cam)


def vmax(x):
    if world == 1:
        return x
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def vsum(x):
    if world == 1:
        return x
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t)
    return float(t.item())


out = {"workload": work, "resolution": [W, H], "n_gpus": world, "strip_bounds": bounds, "halo_rows": halo, "sweep_ms_per_frame": {}, "convergence": {}}
# ---- sweep: k x passes, orbit, timed on the device
fr, grp = make()
f0 = 0          # the orbit runs on across the 24 settings (back and forth over its 60 poses): every frame follows its predecessor
miss = 0
for passes in (1, 2, 3):
    for k in range(1, 9):
        prm = rb.default_params(reuse=3, radius=radius, k=k, passes=passes)
        for i in range(4):
            render(fr, grp, base.orbit(bench.orbit_index(f0 + i)), prm, f0 + i, 0)
        fr.sync()
        if world > 1:
            dist.barrier()
        if f0 == 0:
            fr.halo_miss_reset()          # frame 0 has no predecessor
        fr.mark(0)
        n = 24
        for i in range(4, 4 + n):
            render(fr, grp, base.orbit(bench.orbit_index(f0 + i)), prm, f0 + i, 0)
        fr.mark(1)
        out["sweep_ms_per_frame"]["k%d_p%d" % (k, passes)] = round(vmax(fr.elapsed_ms(0, 1)) / n, 4)
        f0 += 4 + n
miss = fr.halo_miss()
close(fr, grp)
out["halo_miss"] = int(vsum(float(miss)))

# ---- convergence against the path-traced image (static camera, running mean)
P_local = (rows[1] - rows[0]) * W


def accumulate(prm, unbiased_single=False):
    fr, grp = make()
    for it in range(frames):
        if prm is None:
            fr.pathtrace_direct(base, 100000 + it, it)
        else:
            render(fr, grp, base, prm, it, it)
    img = fr.read("radiance").astype(np.float64)
    close(fr, grp)
    return img


ref = accumulate(None)


def relmse(img):
    e = float((((img - ref) ** 2).sum(1) / (ref.sum(1) ** 2 + 1e-2)).sum())
    return vsum(e) / (W * H)


runs = [("ris_only", rb.default_params(reuse=0, radius=radius))]
runs += [("spatiotemporal_k5_p%d" % p, rb.default_params(reuse=3, radius=radius, k=5, passes=p)) for p in (1, 2, 3)]
if world == 1:
    runs += [("unbiased_spatiotemporal_k5_p%d" % p, rb.default_params(reuse=3, radius=radius, k=5, passes=p, unbiased=True)) for p in (1, 2)]
    runs += [("unbiased_spatial_k5_p1", rb.default_params(reuse=2, radius=radius, k=5, passes=1, unbiased=True))]
for name, prm in runs:
    img = accumulate(prm)
    out["convergence"][name] = {"relMSE_vs_ptdirect": relmse(img), "mean": vsum(float(img.sum())) / (3 * W * H)}
out["convergence"]["frames"] = frames
out["convergence"]["mean_ptdirect"] = vsum(float(ref.sum())) / (3 * W * H)
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
