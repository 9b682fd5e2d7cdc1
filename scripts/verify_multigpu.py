"""torchrun -n N scripts/verify_multigpu.py : strips over N GPUs (NCCL halo exchange) == the single-GPU frame, bit for bit.

Every rank renders its strip of a 1080p spatiotemporal orbit for a few frames; rank 0 additionally renders the full
frame on its own GPU; the strips' radiance / history reservoirs / light indices are gathered and compared."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import restir_b200 as rb
from bench import StripExchange
from restir_b200 import scenes, strips

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rb.init(local)
W, H, radius, frames = 1920, 1080, 30.0, 4
sd = scenes.procedural(1, 200000, 10000, (W, H))
sc = rb.Scene.from_arrays(sd)
halo = strips.default_halo(radius)
base = rb.Camera.from_scene(sd)
probe = sc.frame(W, H)
probe.gbuffer_render(base.orbit(0))
bounds = strips.balanced_bounds(strips.row_cost_from_matid(probe.read("matid"), W), world, min_rows=halo)
probe.close()
rows = strips.strip_rows(H, world, rank, bounds)
fr = sc.frame(W, H, rows=rows, halo=halo)
fr.set_stream(torch.cuda.current_stream().cuda_stream)
full = sc.frame(W, H) if rank == 0 else None
prm = rb.default_params(reuse=3, radius=radius)
plan = strips.exchange_plan(H, world, halo, bounds)


fr.set_halo_render(False)          # the benchmarked configuration: G-buffer halo rows travel with the reservoirs
exchange = StripExchange(fr, plan, rank)


bad_total = 0
for k in range(frames):
    cam = base.orbit(k)
    fr.gbuffer_render(cam)
    exchange.join()
    fr.restir_phase_a(cam, prm, k, 0)
    exchange(["geom_cur", "matid_cur", "resv_temp", "resv_out"])      # one exchange per frame
    fr.restir_phase_b(cam, prm, k, 0)
    fr.gbuffer_update(cam)
    if full is not None:
        full.gbuffer_render(cam); full.restir_direct(cam, prm, k, 0); full.gbuffer_update(cam)
    for name in ("radiance", "reservoir", "light_index", "matid", "motion"):
        mine = torch.from_numpy(np.ascontiguousarray(fr.read(name)).view(np.uint8).reshape(-1).copy())
        sizes = [(bounds[r + 1] - bounds[r]) * W * (mine.numel() // fr.npix) for r in range(world)]
        mx = max(sizes)                                   # strips are uneven: gather padded, then trim
        padded = torch.zeros(mx, dtype=torch.uint8, device="cuda")
        padded[:mine.numel()] = mine.cuda()
        parts = [torch.empty(mx, dtype=torch.uint8, device="cuda") for _ in sizes] if rank == 0 else None
        dist.gather(padded, parts, dst=0)
        if rank == 0:
            whole = np.ascontiguousarray(full.read(name)).view(np.uint8).reshape(H * W, -1)
            got = torch.cat([p_[:n_] for p_, n_ in zip(parts, sizes)]).cpu().numpy().reshape(H * W, -1)
            bad = int((whole != got).any(1).sum())
            bad_total += bad
            if bad:
                print("frame", k, name, "pixels differing:", bad)
miss = torch.tensor([fr.halo_miss()], device="cuda")
dist.all_reduce(miss)
if rank == 0:
    print(json.dumps({"verify_multigpu": "ok" if bad_total == 0 and int(miss.item()) == 0 else "MISMATCH", "n_gpus": world, "frames": frames,
                      "resolution": [W, H], "pixels_differing": bad_total, "halo_miss": int(miss.item()), "halo_rows": halo, "strip_bounds": bounds}))
dist.destroy_process_group()
