#!/usr/bin/env python
"""bench.py -- ReSTIR DI frame throughput on B200 (contract: see the task statement / DESIGN.md section 7).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config2] [--impl b200|reference]

A step is one frame of the hot path (G-buffer + ReSTIR DI, main.cpp:164-167) of the workload's camera orbit.
Default workload = BASELINE.json configs[3] (the configuration the metric and the north-star targets are quoted on):
procedural 1M-triangle scene with 100k emissive triangles, 3840x2160, spatiotemporal reuse (temporal cap 20, 5 spatial
neighbours, r = 30 px), 60-frame orbiting camera.  At N = 1 the same JSON line carries a `targets` block with the
device-timed ms/frame of config4_1080p (the <= 2 ms target), config3 and config2.  metric = Mpixel/s (higher is
better), ms_per_step = ms/frame.  `value` is timed on the device (CUDA events on the launching stream) with everything resident in HBM;
`e2e` goes through rstr_render_frame_host (camera in from the host, tone-mapped 8-bit frame back into pinned host
memory every step).  N > 1: one process per GPU under torchrun, horizontal image strips, reservoir halo rows
exchanged with NCCL send/recv; strong scaling (the image is fixed).

--impl reference times the reference's own CPU implementation of the same path (oracle/_ref/libref_harness.so when
it was built from /root/reference, else the oracle port) on the host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (description, scene factory args, resolution, reuse, radius)
    "config1": ("Cornell box 800x800 RIS-only (32 candidates, 1 spp)", ("cornell",), (800, 800), 0, 5.0),
    "config2": ("Cornell box 1920x1080 spatiotemporal (cap 20, k=5, r=30px), 60-frame orbit", ("cornell",), (1920, 1080), 3, 30.0),
    "config3": ("procedural 200k tris / 10k emissive, 1080p spatiotemporal (r=30px), orbit", ("gen", 1, 200_000, 10_000), (1920, 1080), 3, 30.0),
    "config4": ("procedural 1M tris / 100k emissive, 3840x2160 spatiotemporal (r=30px), orbit", ("gen", 1, 1_000_000, 100_000), (3840, 2160), 3, 30.0),
    "config4_1080p": ("procedural 1M tris / 100k emissive, 1080p spatiotemporal (r=30px), orbit", ("gen", 1, 1_000_000, 100_000), (1920, 1080), 3, 30.0),
}
# algorithmic HBM bytes per pixel per frame with the reference's fp32 layouts (SURVEY.md 8d / BASELINE.md section 4)
BYTES_PER_PIXEL = {0: 96, 1: 176, 2: 248, 3: 248}
# ... split per kernel for the spatiotemporal frame: G-buffer write 36 | phase A: own G-buffer 24 + previous G-buffer 20 +
# previous reservoir 36 + post-temporal reservoir 36 + history reservoir 36 = 152 | phase B: reservoir 36 + albedo 12 + radiance 12 = 60
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from `ncu --set full` captures (profiles/)
NCU_TRAFFIC = {("config2", "ris"): (452.3e6, "profiles/r01_prof_config2_r5.summary.txt (k_gbuffer_restir_a: 66.0 MB read + 386.3 MB written; round-1 capture of the fused kernel)"),
               # the staged phase A: four launches, DRAM bytes summed.  Above the 1559 MB of per-pixel algorithmic bytes because it also counts the
               # scene (tree / triangles / light records streamed from HBM: 216 + 165 + 303 MB read) and the planes that carry a pixel from one
               # stage to the next (hit records, the staged reservoir, the queue: the price of cutting the frame where the parallelism changes)
               ("config4", "ris"): (2897.1e6, "profiles/r02_c25_prof_k_{primary,candidates,shadow,temporal}_config4.summary.txt: 914.8 + 498.5 + 351.6 + 1132.1 MB (read + written)"),
               ("config4", "spatial"): (699.0e6, "profiles/r02_c25_prof_k_restir_b_config4.summary.txt (652.6 MB read + 46.4 MB written)")}
KERNEL_BYTES_PER_PIXEL = {"gbuffer": 36, "ris": {0: 60, 1: 140, 2: 116, 3: 152}, "spatial": 60}


def orbit_index(k: int) -> int:
    """The workloads are 60-frame orbits (t_k = k * 2.7 / 60); longer runs sweep that arc back and forth so that
    every step renders one of those 60 camera poses and consecutive frames stay temporally coherent."""
    m = k % 120
    return m if m < 60 else 119 - m


def make_scene(spec, resolution):
    from restir_b200 import scenes

    if spec[0] == "cornell":
        return scenes.cornell_box(resolution)
    return scenes.procedural(spec[1], spec[2], spec[3], resolution)


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path).read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for i, n in enumerate(names):
                if f[5 + i].lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


class quiet_stdout:
    """The reference's host code prints progress with std::cout (bvh.cpp:15,128); keep the JSON line alone on stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.saved)
        os.close(self.null)


# --------------------------------------------------------------------------------------------- reference arm (CPU)
def cpu_frames(kind, sd, reuse, radius, frames, warmup):
    with quiet_stdout():
        return _cpu_frames(kind, sd, reuse, radius, frames, warmup)


def _cpu_frames(kind: str, sd, reuse: float, radius: float, frames: int, warmup: int):
    """Times `frames` frames of the orbit on the host cores; returns (seconds, threads)."""
    from oracle.oracle import Oracle, default_params, make_camera, orbit_camera

    orc = Oracle(kind)
    orc.lib.orc_set_threads(0)      # all host cores: torchrun exports OMP_NUM_THREADS=1 to its workers
    W, H = sd.resolution
    so = orc.scene(sd)
    fo = so.frame(W, H)
    base = make_camera(sd)
    orc.lib.orc_camera_update(C.byref(base))
    prm = default_params(reuse=reuse, radius=radius)
    t0 = None
    for k in range(warmup + frames):
        if k == warmup:
            t0 = time.perf_counter()
        cam = orbit_camera(orc, base, orbit_index(k))
        fo.gbuffer_render(cam)
        fo.restir_direct(cam, prm, k, 0)
        fo.gbuffer_update(cam)
    dt = time.perf_counter() - t0
    return dt, orc.threads()


def workload_config(name: str, triangles: int, emissive: int) -> dict:
    """The workload description both arms print (identical dicts: the driver compares them)."""
    desc, spec, res, reuse, radius = WORKLOADS[name]
    return {"workload": name, "description": desc, "resolution": list(res), "triangles": int(triangles), "emissive_triangles": int(emissive),
            "reuse": reuse, "candidates": 32, "temporal_cap": 20, "spatial_neighbours": 5, "spatial_radius_px": radius}


def reference_kind() -> str:
    from oracle.oracle import have_reference

    return "reference" if have_reference() else "port"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    desc, spec, res, reuse, radius = WORKLOADS[args.workload]
    sd = make_scene(spec, res)
    kind = reference_kind()
    P = res[0] * res[1]
    # bound the sample so that the whole run stays within minutes: estimate from one frame
    t1, threads = cpu_frames(kind, sd, reuse, radius, 1, 0)
    steps = args.steps
    budget = 150.0
    if t1 * (steps + args.warmup) > budget:
        steps = max(1, int(budget / t1) - args.warmup)
    warm = min(args.warmup, max(0, int(budget / t1) - steps))
    dt, threads = cpu_frames(kind, sd, reuse, radius, steps, warm)
    ms = dt / steps * 1e3
    val = P / (ms * 1e-3) / 1e6
    line = {
        "impl": "reference", "metric": "ReSTIR DI Mpixel/s", "value": val, "unit": "Mpixel/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, sd.num_tris, sd.num_lights),
        "cpu_baseline": {"value": val, "unit": "Mpixel/s", "cores": threads, "kind": kind,
                         "sample": "%d frames of the orbit at %dx%d (OpenMP over rows, %d threads)" % (steps, res[0], res[1], threads)},
        "e2e": {"value": val, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- B200 arm
def window_frames(k0: int, n: int, count: int = 6):
    """Up to `count` frame indices spread over [k0, k0 + n): the stretch of the orbit that is about to be rendered.  The strip
    cuts are placed for that stretch (the balance moves with the camera: cuts that are right for the orbit's average leave the
    slowest rank 5 % above the mean on any one stretch of it)."""
    n = max(int(n), 1)
    step = max(1, n // count)
    return list(range(k0, k0 + n, step))[:count]
PIPELINES = {"auto": None, "staged": True, "fused": False}


class _DevMem:
    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class StripExchange:
    """Neighbour halo exchange of a strip frame over NCCL send/recv (NVLink): one batch_isend_irecv per call, on
    tensors that alias the frame's device planes (no staging copies).  ``deferred=True`` issues the batch on a side
    stream behind everything queued so far and returns; ``join()`` makes the main stream wait for it (used for the
    history reservoirs, which are only needed by the NEXT frame's temporal step and so travel under its G-buffer)."""

    def __init__(self, fr, plan, rank):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.fr, self.rank = torch, dist, fr, rank
        self.plan = [p for p in plan if rank in (p[0], p[1])]
        self.side = torch.cuda.Stream(priority=-1)
        self.ev_ready, self.ev_done = torch.cuda.Event(), torch.cuda.Event()
        self.pending = False
        self.keep = []

    def _ops(self, planes):
        torch, dist = self.torch, self.dist
        ops = []
        for plane in planes:
            for src, dst, r0, r1 in self.plan:
                ptr, rowbytes = self.fr.plane_row(plane, r0)
                t = torch.as_tensor(_DevMem(ptr, rowbytes * (r1 - r0)), device="cuda")
                self.keep.append(t)
                ops.append(dist.P2POp(dist.isend if self.rank == src else dist.irecv, t, dst if self.rank == src else src))
        return ops

    def __call__(self, planes, deferred=False):
        torch, dist = self.torch, self.dist
        self.keep = self.keep[-64:]
        ops = self._ops(planes)
        if not ops:
            return
        if not deferred:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            return
        main = torch.cuda.current_stream()
        self.ev_ready.record(main)
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.ev_ready)
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            self.ev_done.record(self.side)
        self.pending = True

    def join(self):
        if self.pending:
            self.torch.cuda.current_stream().wait_event(self.ev_done)
            self.pending = False


def measure_motion_rows(sc, base, W, H, rb):
    """Per-frame bound for the temporal halo: the largest |row(motion) - row| over every consecutive pose pair
    of the workload's orbit, taken from G-buffer-only renders of the full image (deterministic: every rank computes the
    same number)."""
    fr = sc.frame(W, H)
    for k in range(60):
        cam = base.orbit(k)
        fr.gbuffer_render(cam); fr.gbuffer_update(cam)
    rows = fr.motion_rows()
    fr.close()
    return rows


def time_workload(rb, name, steps, warmup, staged=None):
    """Device-timed ms/frame of another workload on this GPU (the `targets` block): same frame loop, CUDA events on the
    launching stream around `steps` frames."""
    desc, spec, res, reuse, radius = WORKLOADS[name]
    sd = make_scene(spec, res)
    sc = rb.Scene.from_arrays(sd)
    fr = sc.frame(*res)
    fr.set_pipeline(staged)
    base = rb.Camera.from_scene(sd)
    prm = rb.default_params(reuse=reuse, radius=radius)

    def frame(k):
        cam = base.orbit(orbit_index(k))
        fr.gbuffer_render(cam); fr.restir_direct(cam, prm, k, 0); fr.gbuffer_update(cam)

    k = 0
    for _ in range(max(warmup, 3)):
        frame(k); k += 1
    fr.sync()
    fr.mark(0)
    for _ in range(steps):
        frame(k); k += 1
    fr.mark(1)
    ms = fr.elapsed_ms(0, 1) / steps
    stage = {"gbuffer": [], "ris": [], "spatial": []}
    for _ in range(min(steps, 10)):
        frame(k); k += 1
        for n, v in fr.stage_ms().items():
            if n in stage and v > 0:
                stage[n].append(v)
    out = {"ms_per_frame": ms, "mpixel_per_s": res[0] * res[1] / (ms * 1e-3) / 1e6, "resolution": list(res), "triangles": sc.info.numTris,
           "emissive_triangles": sc.info.numLights, "steps": steps,
           "stage_ms": {n: (sum(v) / len(v) if v else 0.0) for n, v in stage.items()}}
    fr.close()
    sc.close()
    return out


def run_b200(args):
    import restir_b200 as rb
    from restir_b200 import strips

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = torch = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        opts = dist.ProcessGroupNCCL.Options()
        opts.is_high_priority_stream = True      # halo exchanges must get SM slots while a frame kernel fills the GPU
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
    rb.init(local)
    desc, spec, res, reuse, radius = WORKLOADS[args.workload]
    W, H = res
    P = W * H
    sd = make_scene(spec, res)
    sc = rb.Scene.from_arrays(sd)
    traced_tree = {"built": "host (binned SAH, bvh_fast.cpp)"}
    if args.traced_tree != "host":
        mode = 1 if args.traced_tree == "gpu-radix" else 0
        sc.build_traced_gpu(mode)                # first call pays the allocations / cub set-up
        traced_tree = {"built": "device (%s, bvh_gpu.cu)" % ("Morton radix tree" if mode else "PLOC"), "build_ms": sc.build_traced_gpu(mode)}
    base = rb.Camera.from_scene(sd)
    prm = rb.default_params(reuse=reuse, radius=radius)
    motion_rows = None
    halo = 0
    if world > 1:
        motion_rows = measure_motion_rows(sc, base, W, H, rb) if (reuse & 1) else 0
        halo = strips.default_halo(radius if (reuse & 2) else 0.0, motion_rows)
    bounds = strips.uniform_bounds(H, world)
    if world > 1 and not args.uniform_strips:
        # measured cost profile: a few frames of the real pipeline on the full image with the kernels' cycle accounting
        # on (SM cycles every 16x8-pixel block held its SM slot, per group of 8 rows), at poses spread over the stretch of the
        # orbit that is timed first; the cuts then give every rank the same summed cost.  Equal-height strips balance
        # badly: sky rows are almost free.  All ranks measure, the sum is all-reduced so that they agree on the cuts.
        pf = sc.frame(W, H)
        pf.row_cost(enable=True)
        for pose in window_frames(args.warmup, args.steps):
            for kk in (pose, pose + 1):
                cam = base.orbit(orbit_index(kk))
                pf.gbuffer_render(cam); pf.restir_direct(cam, prm, kk, 0); pf.gbuffer_update(cam)
        groups = torch.from_numpy(pf.row_cost(enable=False, read=True)).cuda()
        pf.close()
        dist.all_reduce(groups)
        row_cost = np.repeat(groups.cpu().numpy() / 8.0, 8)[:H]
        bounds = strips.balanced_bounds(row_cost, world, min_rows=max(8, halo))
    row_cost = row_cost if (world > 1 and not args.uniform_strips) else None
    rows = plan = fr = exchange = grp = None
    peer = world > 1 and args.exchange == "peer"

    def all_gather_bytes(blob: bytes):
        """The one piece of plumbing the peer data plane needs: every rank's 512-byte handle, in rank order."""
        t = torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda()
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return [bytes(o.cpu().numpy().tobytes()) for o in outs]

    def make_strip():
        nonlocal rows, plan, fr, exchange, grp
        if fr is not None:
            fr.sync()
            if world > 1:
                torch.cuda.synchronize()
                dist.barrier()               # nobody may still be storing into a slab that is about to be freed
            if grp is not None:
                grp.close()
                grp = None
            fr.close()
        rows = strips.strip_rows(H, world, rank, bounds)
        fr = sc.frame(W, H, rows=rows, halo=halo)
        fr.set_pipeline(PIPELINES[args.pipeline])
        fr.set_bands(args.bands)
        if args.no_fusion:
            fr.set_fusion(False)
        plan = strips.exchange_plan(H, world, halo, bounds) if world > 1 else []
        if peer:
            # the library's own data plane: halo rows and the gather are peer stores over NVLink (rstr_strip_group_*)
            grp = rb.StripGroup(fr, rank, world)
            grp.connect(all_gather_bytes(grp.handle()))
        elif world > 1:
            fr.set_stream(torch.cuda.current_stream().cuda_stream)
            fr.set_halo_render(args.render_halo)
            exchange = StripExchange(fr, plan, rank)

    make_strip()

    def frame(k):
        # One frame of a strip: G-buffer and phase A on the strip's own rows (one fused kernel); then ONE exchange carries
        # every halo row anybody needs: what phase B reads (post-temporal reservoirs + the neighbours' G-buffer rows,
        # unless those are rendered locally) and the history reservoirs phase A just wrote, which the NEXT frame's
        # temporal step reads (phase B does not modify them).  --split-exchange sends the history separately, after
        # phase B, on a side stream (the first version).
        cam = base.orbit(orbit_index(k))
        if peer:
            grp.render(cam, prm, k, 0)
            return
        fr.gbuffer_render(cam)
        if world == 1:
            fr.restir_direct(cam, prm, k, 0)
        else:
            exchange.join()
            fr.restir_phase_a(cam, prm, k, 0)
            planes = ([] if args.render_halo else ["geom_cur", "matid_cur"]) + (["resv_temp"] if reuse & 2 else [])
            if (reuse & 1) and not args.split_exchange:
                planes.append("resv_out")
            exchange(planes)
            fr.restir_phase_b(cam, prm, k, 0)
            if (reuse & 1) and args.split_exchange:
                exchange(["resv_history"], deferred=not args.no_overlap)
        fr.gbuffer_update(cam)

    def barrier():
        fr.sync()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- closed-loop refinement of the cuts: the cycle profile ranks rows correctly but is biased between cheap and
    # expensive regions (block-launch overhead, tails), so a few frames are run with the current cuts, every rank
    # reports its measured kernel time, the row costs inside each strip are rescaled by measured / predicted and the
    # cuts are placed again.  Deterministic across ranks: everything is computed from all-gathered numbers.
    # (These frames jump between camera poses: their reprojections can leave the halo, which is why the halo-miss counter
    # is reset after the warm-up below and checked only over consecutive frames of the orbit.)
    refine_log = []

    def refine_cuts(k0, n, tag):
        # measured on the very frames that follow (two lead-in frames for the history, then the stretch itself, at most 120 frames)
        nonlocal row_cost, bounds
        best = None
        for it in range(args.refine + 1):
            tot, cnt = 0.0, 0
            for kk in range(max(k0 - 2, 0), k0 + min(n, 120)):
                frame(kk)
                if kk >= k0:
                    tot += sum(v for nm, v in fr.stage_ms().items() if nm in ("gbuffer", "ris", "spatial")); cnt += 1
            t = torch.tensor([tot / cnt], device="cuda", dtype=torch.float64)
            allt = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
            measured = np.array([float(x.item()) for x in allt])
            refine_log.append({"window": tag, "bounds": list(bounds), "kernel_ms_per_rank": [round(float(x), 3) for x in measured]})
            if best is None or measured.max() < best[0]:
                best = (float(measured.max()), list(bounds))
            if it == args.refine or measured.max() / measured.mean() < 1.02:
                break
            row_cost = strips.refine_row_cost(row_cost, bounds, measured, damping=1.0 if (it == 0 and tag == "value") else 0.6)
            bounds = strips.balanced_bounds(row_cost, world, min_rows=max(8, halo))
            make_strip()
        if list(bounds) != best[1]:          # the cuts whose slowest rank was fastest, not simply the last ones tried
            bounds = best[1]
            make_strip()

    if row_cost is not None:
        refine_cuts(args.warmup, args.steps, "value")

    k = 0
    for _ in range(args.warmup):
        frame(k); k += 1
    # from here on every frame follows its predecessor on the orbit: no read may leave the resident rows
    fr.halo_miss_reset()
    # ---- pass 1: device-timed throughput (no per-frame synchronisation)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = rb.launch_count()
    fr.mark(0)
    for _ in range(args.steps):
        frame(k); k += 1
    fr.mark(1)
    barrier()
    ms_total = max_over_ranks(fr.elapsed_ms(0, 1))
    launches = rb.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    ms_step = ms_total / args.steps
    value = P / (ms_step * 1e-3) / 1e6
    # ---- pass 2: per-kernel durations (CUDA events around every launch on the launching stream)
    stage = {"gbuffer": [], "ris": [], "spatial": []}
    for _ in range(min(args.steps, 20)):
        frame(k); k += 1
        for n, v in fr.stage_ms().items():
            if n in stage and v > 0:
                stage[n].append(v)
    stage_ms = {n: (sum(v) / len(v) if v else 0.0) for n, v in stage.items()}
    shaded_local = int((fr.read("matid") >= 0).sum())     # rays actually traced: 2 primary per pixel + 1 shadow per shaded pixel
    stage_per_rank = None
    if world > 1:
        t = torch.tensor([stage_ms["gbuffer"], stage_ms["ris"], stage_ms["spatial"]], device="cuda", dtype=torch.float64)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        stage_per_rank = [[round(float(x), 4) for x in a.cpu()] for a in allt]
        t = torch.tensor([shaded_local], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        shaded = int(t.item())
    else:
        shaded = shaded_local
    # ---- pass 3: end to end through the host-facing call (N = 1) / strips gathered to rank 0 (N > 1)
    miss_before_e2e = fr.halo_miss()
    bounds_value = list(bounds)
    npix_value = (rows[1] - rows[0]) * W     # rank 0's strip while `value` and the stage times were measured
    if row_cost is not None:                 # the end-to-end frames lie further along the orbit: cuts for that stretch
        refine_cuts(k, 6 + args.steps, "e2e")
    npix_local = (rows[1] - rows[0]) * W
    e2e_ms = e2e_sync_ms = None
    if world == 1:
        # three pinned host frames in flight: the D2H of frame k overlaps the rendering of frames k+1, k+2; every frame's
        # camera goes in and every frame's image comes out inside the timed region (consumed two frames later)
        S = 3
        outs = [rb.pinned_empty(npix_local * 4) for _ in range(S)]

        def e2e_loop(n):
            nonlocal k
            for i in range(n):
                fr.render_frame_host_async(base.orbit(orbit_index(k)), prm, k, 0, rb.TONEMAP_ACES, outs[i % S], i % S); k += 1
                if i >= S - 1:
                    fr.wait_host((i - (S - 1)) % S)
            for i in range(max(0, n - (S - 1)), n):
                fr.wait_host(i % S)

        e2e_loop(6)
        fr.sync()
        t0 = time.perf_counter()
        e2e_loop(args.steps)
        e2e_ms = (time.perf_counter() - t0) / args.steps * 1e3
        assert all(o.max() > 0 for o in outs)
        # the synchronous single call, for reference
        t0 = time.perf_counter()
        for i in range(args.steps):
            fr.render_frame_host(base.orbit(orbit_index(k)), prm, k, 0, rb.TONEMAP_ACES, outs[0]); k += 1
        e2e_sync_ms = (time.perf_counter() - t0) / args.steps * 1e3
        for o in outs:
            rb.pinned_free(o)
    elif peer:
        # every rank tone-maps its strip straight into rank 0's full-frame LDR slot (peer stores); rank 0 copies the
        # assembled frame to one of three pinned host frames on a copy stream and consumes it two frames later
        S = 3
        outs = [rb.pinned_empty(P * 4) for _ in range(S)] if rank == 0 else [None] * S

        def e2e_loop(n):
            nonlocal k
            for i in range(n):
                frame(k); k += 1
                grp.present(rb.TONEMAP_ACES, outs[i % S], i % S)
                if rank == 0 and i >= S - 1:
                    grp.wait_host((i - (S - 1)) % S)
            if rank == 0:
                for i in range(max(0, n - (S - 1)), n):
                    grp.wait_host(i % S)
            fr.sync()

        e2e_loop(6)
        barrier()
        if row_cost is not None:
            fr.halo_miss_reset()             # the refinement frames above jumped between poses
        t0 = time.perf_counter()
        e2e_loop(args.steps)
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) / args.steps * 1e3)
        if rank == 0:
            assert all(o.max() > 0 for o in outs)
            for o in outs:
                rb.pinned_free(o)
    else:
        # (--exchange nccl) strips -> GPU 0 -> pinned host frame, on a side stream so that it overlaps the next frame's rendering:
        # ranks > 0 send their tone-mapped strip (NCCL over NVLink), rank 0 receives into place and copies D2H
        counts = [(bounds[r + 1] - bounds[r]) * W * 4 for r in range(world)]
        offs = [bounds[r] * W * 4 for r in range(world)]
        host = torch.empty(P * 4, dtype=torch.uint8, pin_memory=True) if rank == 0 else None
        full = torch.empty(P * 4, dtype=torch.uint8, device="cuda") if rank == 0 else None
        mine = full[offs[0]:offs[0] + counts[0]] if rank == 0 else torch.empty(counts[rank], dtype=torch.uint8, device="cuda")
        side = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        ev_ldr, ev_sent = torch.cuda.Event(), torch.cuda.Event()

        def e2e_frame(kk):
            frame(kk)
            main.wait_event(ev_sent)                 # the previous frame's strip must have left `mine`
            fr.tonemap(rb.TONEMAP_ACES, 1.0)
            fr.read_into_device("ldr", mine.data_ptr(), counts[rank])
            ev_ldr.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ev_ldr)
                if rank == 0:
                    reqs = [dist.irecv(full[offs[r]:offs[r] + counts[r]], r) for r in range(1, world)]
                    for q in reqs:
                        q.wait()
                    host.copy_(full, non_blocking=True)
                else:
                    dist.isend(mine, 0).wait()
                ev_sent.record(side)

        ev_sent.record(side)
        for _ in range(3):
            e2e_frame(k); k += 1
        barrier(); side.synchronize()
        if row_cost is not None:
            fr.halo_miss_reset()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_frame(k); k += 1
        side.synchronize()
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) / args.steps * 1e3)
        if rank == 0:
            assert int(host.max()) > 0
    halo_miss = fr.halo_miss() + (miss_before_e2e if row_cost is not None else 0)     # (re-cut strips are new frames: their counter starts at 0)
    peer_error = bool(grp.error()) if grp is not None else False
    if world > 1:
        t = torch.tensor([halo_miss, int(peer_error)], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        halo_miss, peer_error = int(t[0].item()), bool(t[1].item())

    if rank == 0:
        peak, peak_src = peaks()
        dom = max(stage_ms, key=lambda n: stage_ms[n])
        bpp = KERNEL_BYTES_PER_PIXEL[dom]
        bpp = bpp[reuse] if isinstance(bpp, dict) else bpp
        fused = stage_ms["gbuffer"] == 0.0 and stage_ms["ris"] > 0.0      # G-buffer + phase A ran as one kernel
        if fused and dom == "ris":
            bpp += KERNEL_BYTES_PER_PIXEL["gbuffer"]
        strip_px = npix_value
        achieved = bpp * strip_px / (stage_ms[dom] * 1e-3) / 1e9 if stage_ms[dom] > 0 else 0.0
        frame_bytes = BYTES_PER_PIXEL[reuse] * P
        info = sc.info
        cfg = workload_config(args.workload, info.numTris, info.numLights)
        traffic = NCU_TRAFFIC.get((args.workload, dom), (None, None))
        line = {
            "metric": "ReSTIR DI Mpixel/s", "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "pipeline": args.pipeline,
            "strips": {"parallelism": "strips%d" % world, "exchange": None if world == 1 else ("peer stores over NVLink (rstr_strip_group)" if peer else "NCCL send/recv"), "halo_rows": halo, "motion_rows_bound": motion_rows, "strip_bounds": bounds_value, "strip_bounds_e2e": bounds,
                       "gbuffer_halo": None if world == 1 else ("rendered locally" if args.render_halo else "received from the neighbours"),
                       "l2": "no explicit flush: the per-frame pixel planes (%.0f MB) exceed the 126 MB L2" % (P * 212 / 1e6)},
            "e2e": {"value": P / (e2e_ms * 1e-3) / 1e6, "unit": "Mpixel/s", "ms_per_step": e2e_ms, "ms_per_step_synchronous_call": e2e_sync_ms,
                    "api": "rstr_render_frame_host_async + rstr_frame_wait_host (three pinned host frames in flight)" if world == 1 else
                    ("rstr_strip_group_frame + rstr_strip_group_present: strips tone-mapped into rank 0's frame by peer stores, D2H on rank 0, three host frames in flight" if peer
                     else "strips gathered to rank 0 over NCCL, one D2H"),
                    "h2d_bytes_per_step": C.sizeof(rb.api.RstrCamera) + C.sizeof(rb.RstrParams), "d2h_bytes_per_step": P * 4},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": {"gbuffer": "k_gbuffer", "ris": ("k_primary + k_candidates + k_shadow + k_temporal" if (args.pipeline == "staged" or (args.pipeline == "auto" and info.tracedNodes > 1024)) else "k_gbuffer_restir_a") if fused else "k_restir_a", "spatial": "k_restir_b"}[dom],
                         "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic[0] if world == 1 else None, "traffic_source": traffic[1],
                         "algorithmic_bytes_per_pixel": bpp, "kernel_ms": stage_ms[dom],
                         "note": "traversal / light-gather bound kernel; HBM fraction reported as required, rays/s below is the telling figure"},
            "stage_ms": stage_ms,
            "stage_ms_per_rank": stage_per_rank,
            "strip_refinement": refine_log or None,
            "frame_hbm": {"algorithmic_bytes_per_frame": frame_bytes, "achieved_gbs": frame_bytes / (ms_step * 1e-3) / 1e9, "frac_of_peak": frame_bytes / (ms_step * 1e-3) / 1e9 / peak},
            "rays_per_s": (2.0 * P + shaded) / (ms_step * 1e-3),
            "rays_per_frame": {"primary": 2 * P, "shadow_upper_bound": shaded, "note": "2 primary rays per pixel (centre + jittered, one shared walk) + 1 shadow ray per shaded pixel whose reservoir weight is non-zero"},
            "halo_miss": halo_miss,
            "host_build_s": info.buildSeconds,
            "traced_tree": traced_tree,
            "build_id": rb.api.build_id(),
        }
        if world == 1 and not args.no_targets:
            # the north star's numeric targets, device-timed on the same box in the same run
            line["targets"] = {}
            for name in ("config4_1080p", "config3", "config2"):
                if name != args.workload:
                    line["targets"][name] = time_workload(rb, name, min(args.steps, 60), 5, PIPELINES[args.pipeline])
            t4 = line["targets"].get("config4_1080p")
            if t4:
                line["targets"]["north_star"] = {"config4_1080p_ms_per_frame": t4["ms_per_frame"], "target_ms": 2.0, "met": t4["ms_per_frame"] <= 2.0}
        if world == 1 and not args.no_reference_cuda:
            # the reference's own CUDA build on this B200 (oracle/_ref/ref_headless_*, built from /root/reference by
            # oracle/build_ref_cuda.sh): same scene through the reference's scene-file grammar, same orbit, same settings;
            # as shipped (device-wide sync after every launch) and with ERRORCHECK 0 (cudaUtil.h:8)
            try:
                from scripts import ref_cuda_compare as rcc

                if rcc.have_reference_cuda():
                    import shutil
                    from restir_b200 import scenes as _scenes

                    tmp = tempfile.mkdtemp()
                    txt = _scenes.write_scene_files(sd, tmp, "scene")
                    t = rcc.reference_cuda_times(txt, max(3, min(args.steps, 20)), reuse, radius, warmup=3)
                    shutil.rmtree(tmp, ignore_errors=True)
                    line["reference_cuda_ms"] = {"as_shipped": t.get("as_shipped_ms_per_frame"), "errorcheck0": t.get("errorcheck0_ms_per_frame"),
                                                 "speedup_vs_as_shipped": t["as_shipped_ms_per_frame"] / ms_step,
                                                 "note": "GBuffer::render + ReSTIRDirect of the reference's kernels compiled for sm_100a, CUDA events, same scene / orbit / settings (radius literal 30)"}
                else:
                    line["reference_cuda_ms"] = None
            except Exception as e:       # a baseline, not the product: never fails the bench
                line["reference_cuda_ms"] = {"error": str(e)[-300:]}
        if world == 1 and not args.no_targets:
            line["restir_gi"] = gi_side_measurement()
        if world == 1 and not args.no_cpu_baseline:
            kind = reference_kind()
            t1, threads = cpu_frames(kind, sd, reuse, radius, 1, 0)
            n = max(1, min(args.steps, 60, int(15.0 / max(t1, 1e-3))))
            dt, threads = cpu_frames(kind, sd, reuse, radius, n, 1 if t1 < 5 else 0)
            line["cpu_baseline"] = {"value": P / (dt / n) / 1e6, "unit": "Mpixel/s", "cores": threads, "kind": kind,
                                    "sample": "%d frames of the same orbit at %dx%d, OpenMP over rows" % (n, W, H)}
        if halo_miss != 0:
            line["invalid"] = "halo_miss = %d: a strip read rows that were not resident, the image is NOT the single-GPU frame" % halo_miss
        if peer_error:
            line["invalid"] = "a wait on a peer GPU timed out"
        print(json.dumps(line))
    if world > 1:
        fr.sync(); torch.cuda.synchronize(); dist.barrier()
    if grp is not None:
        grp.close()
    fr.close()
    sc.close()
    if world > 1:
        dist.destroy_process_group()
    if halo_miss != 0 or peer_error:
        sys.exit(3)


def gi_side_measurement(timeout_s: int = 120):
    """ReSTIR GI (rstr_restir_indirect, depth 3, temporal reuse) on the 1 M-triangle scene at 1080p, device-timed by
    scripts/gi_bench.py in a child process after the frame measurements: ms / frame of the ray-queue and the staged launch forms.  A side
    measurement of a "next" row: it can never fail or delay the bench line beyond its time-out."""
    try:
        r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "scripts", "gi_bench.py"), "--workloads", "config4_1080p",
                            "--steps", "10", "--modes", "queued", "staged", "--out", os.devnull], capture_output=True, text=True, timeout=timeout_s)
        rows = [json.loads(x) for x in r.stdout.splitlines() if x.startswith("{")]
        if not rows:
            return {"error": (r.stderr or "no output")[-300:]}
        return {"workload": "config4_1080p", "trace_depth": rows[0]["trace_depth"], "reuse": "temporal", "steps": rows[0]["steps"],
                "ms_per_frame": {("ray_queues" if "ray queues" in x["pipeline"] else "staged"): x["gi_ms_per_frame"] for x in rows},
                "fixup_pixels_per_frame": rows[0]["fixup_pixels_per_frame"], "lit_fraction": rows[0]["lit_fraction"],
                "note": "rstr_restir_indirect alone (the G-buffer render of each frame is outside the CUDA events), scripts/gi_bench.py in a child process"}
    except Exception as e:
        return {"error": str(e)[-300:]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    # 120 frames = one sweep of the 60-frame orbit there and back: ~1.7 s on the device at 4K, long enough for stable clocks / nvidia-smi samples
    ap.add_argument("--steps", type=int, default=120)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-cuda", action="store_true", help="N = 1: skip timing the reference's own CUDA build (oracle/_ref/ref_headless_*) on the same workload")
    ap.add_argument("--no-targets", action="store_true", help="N = 1: skip the `targets` block (device-timed config4_1080p / config3 / config2)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="N > 1: halo rows / gather as the library's peer stores over NVLink (default) or as NCCL send/recv issued from here (A/B)")
    ap.add_argument("--uniform-strips", action="store_true", help="equal-height strips instead of cost-balanced cuts (N > 1)")
    ap.add_argument("--refine", type=int, default=8, help="N > 1: closed-loop refinements of the strip cuts before the timed run")
    ap.add_argument("--pipeline", default="auto", choices=["auto", "staged", "fused"], help="phase A as the staged kernel pipeline, as one fused kernel, or chosen by scene size (default)")
    ap.add_argument("--bands", type=int, default=1, help="row bands of the staged pipeline whose queue kernels overlap the next band's primary walk (1 = none)")
    ap.add_argument("--no-fusion", action="store_true", help="separate G-buffer and phase-A kernels instead of the fused one (A/B)")
    ap.add_argument("--split-exchange", action="store_true", help="N > 1: history reservoirs in a second exchange after phase B (A/B)")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: wait for the history-reservoir exchange at the end of the frame instead of under the next G-buffer")
    ap.add_argument("--render-halo", action="store_true", help="N > 1: every strip renders its G-buffer halo rows itself instead of receiving them from its neighbours")
    ap.add_argument("--traced-tree", default="host", choices=["host", "gpu", "gpu-radix"], help="the tree the kernels trace: the host's binned-SAH build (default) or rebuilt on the device (rstr_scene_build_traced_gpu: PLOC, or the plain radix tree)")
    ap.add_argument("--quick", action="store_true", help="the B200 measurement only: no CPU baseline, no targets block, no reference CUDA timing (A/B runs)")
    args = ap.parse_args()
    if args.quick:
        args.no_cpu_baseline = args.no_targets = args.no_reference_cuda = True
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
