"""torchrun -n N scripts/verify_multigpu.py [workload] [frames]: strips over N GPUs == the single-GPU frame, bit for bit.

Every rank renders its strip of the workload's spatiotemporal orbit through the library's peer data plane
(rstr_strip_group_*: halo rows stored into the neighbours' memory over NVLink, gather by peer stores into rank 0's frame);
rank 0 additionally renders the full frame on its own GPU; the strips' radiance / history reservoirs / light indices /
G-buffer are gathered and compared, and so is the LDR frame rank 0 assembled.  workload: a bench.py workload name
(default config3; config4 = 1M triangles at 3840x2160)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import bench
import restir_b200 as rb
from restir_b200 import strips

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
work = sys.argv[1] if len(sys.argv) > 1 else "config3"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 4
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rb.init(local)
desc, spec, (W, H), reuse, radius = bench.WORKLOADS[work]
sd = bench.make_scene(spec, (W, H))
sc = rb.Scene.from_arrays(sd)
base = rb.Camera.from_scene(sd)
motion = bench.measure_motion_rows(sc, base, W, H, rb)
halo = strips.default_halo(radius, motion)
probe = sc.frame(W, H)
probe.gbuffer_render(base.orbit(0))
bounds = strips.balanced_bounds(strips.row_cost_from_matid(probe.read("matid"), W), world, min_rows=halo)
probe.close()
rows = strips.strip_rows(H, world, rank, bounds)
fr = sc.frame(W, H, rows=rows, halo=halo)
grp = rb.StripGroup(fr, rank, world)
t = torch.frombuffer(bytearray(grp.handle()), dtype=torch.uint8).cuda()
outs = [torch.empty_like(t) for _ in range(world)]
dist.all_gather(outs, t)
grp.connect([bytes(o.cpu().numpy().tobytes()) for o in outs])
full = sc.frame(W, H) if rank == 0 else None
prm = rb.default_params(reuse=reuse, radius=radius)
host = rb.pinned_empty(W * H * 4) if rank == 0 else None

bad_total = 0
for k in range(frames):
    cam = base.orbit(k)
    grp.render(cam, prm, k, 0)
    grp.present(rb.TONEMAP_ACES, host, k % 3)
    if rank == 0:
        grp.wait_host(k % 3)
        full.gbuffer_render(cam); full.restir_direct(cam, prm, k, 0); full.gbuffer_update(cam)
        full.tonemap(rb.TONEMAP_ACES, 1.0)
        bad = int((full.read("ldr") != host.reshape(-1, 4)).any(1).sum())
        bad_total += bad
        if bad:
            print("frame", k, "gathered LDR frame: pixels differing:", bad)
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
st = torch.tensor([fr.halo_miss(), int(grp.error())], device="cuda")
dist.all_reduce(st)
if rank == 0:
    ok = bad_total == 0 and int(st[0].item()) == 0 and int(st[1].item()) == 0
    print(json.dumps({"verify_multigpu": "ok" if ok else "MISMATCH", "exchange": "peer stores (rstr_strip_group)", "workload": work, "n_gpus": world, "frames": frames,
                      "resolution": [W, H], "pixels_differing": bad_total, "halo_miss": int(st[0].item()), "peer_timeouts": int(st[1].item()),
                      "halo_rows": halo, "motion_rows_bound": motion, "strip_bounds": bounds}))
fr.sync(); torch.cuda.synchronize(); dist.barrier()
grp.close()
dist.destroy_process_group()
