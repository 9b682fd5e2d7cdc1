"""GPU box: the reference's OWN CUDA build (oracle/_ref/ref_headless_*, see oracle/build_ref_cuda.sh) next to this
library -- timing (reference as shipped, device sync after every launch) and GPU-vs-GPU parity on the same scene file.

    python scripts/ref_cuda_compare.py [config2|config3] [frames]
"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import restir_b200 as rb
from bench import WORKLOADS, make_scene
from restir_b200 import scenes

REF = os.path.join(ROOT, "oracle", "_ref")


def run_ref(binary, scene_txt, frames, warmup, reuse, dump=None, dump_frame=-1, denoise=False):
    cmd = [os.path.join(REF, binary), scene_txt, str(frames), str(warmup), str(reuse)]
    if dump:
        cmd += [dump, str(dump_frame)]
    env = dict(os.environ, REF_DENOISE="1") if denoise else None
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=os.path.dirname(scene_txt), env=env)
    if r.returncode != 0:
        raise RuntimeError("ref_headless failed: " + r.stderr[-2000:])
    return json.loads(r.stdout.strip().splitlines()[-1])


def have_reference_cuda() -> bool:
    return all(os.path.exists(os.path.join(REF, b)) for b in ("ref_headless_r5", "ref_headless_r30"))


def reference_cuda_times(txt, frames, reuse, radius, warmup=5):
    """ms/frame of the reference's own CUDA kernels (GBuffer::render + ReSTIRDirect, CUDA events) on a scene file: as shipped
    (device-wide sync after every launch, cudaUtil.h:13-16) and, when that build exists, with ERRORCHECK 0 (cudaUtil.h:8)."""
    out = {}
    binary = "ref_headless_r30" if radius == 30.0 else "ref_headless_r5"
    out["as_shipped_ms_per_frame"] = run_ref(binary, txt, frames, warmup, reuse)["ms_per_frame"]
    if radius == 30.0 and os.path.exists(os.path.join(REF, "ref_headless_r30_nosync")):
        out["errorcheck0_ms_per_frame"] = run_ref("ref_headless_r30_nosync", txt, frames, warmup, reuse)["ms_per_frame"]
    return out


def compare(work="config2", frames=60, lib_times=True):
    desc, spec, res, reuse, radius = WORKLOADS[work]
    sd = make_scene(spec, res)
    tmp = tempfile.mkdtemp()
    txt = scenes.write_scene_files(sd, tmp, "scene")
    out = {"workload": work, "description": desc}
    # ---- timing: reference CUDA build vs this library, same scene file, same orbit
    out["reference_cuda"] = reference_cuda_times(txt, frames, reuse, radius)
    out["reference_cuda_ms_per_frame"] = out["reference_cuda"]["as_shipped_ms_per_frame"]
    rb.init(0)
    sc = rb.Scene.from_file(txt)
    base = sc.camera
    W, H = res
    P = W * H
    if lib_times:
        fr = sc.frame(*res)
        prm = rb.default_params(reuse=reuse, radius=radius)
        for k in range(5):
            cam = base.orbit(k); fr.gbuffer_render(cam); fr.restir_direct(cam, prm, k, 0); fr.gbuffer_update(cam)
        fr.sync(); fr.mark(0)
        for k in range(5, 5 + frames):
            cam = base.orbit(k); fr.gbuffer_render(cam); fr.restir_direct(cam, prm, k, 0); fr.gbuffer_update(cam)
        fr.mark(1); fr.sync()
        out["restir_b200_ms_per_frame"] = fr.elapsed_ms(0, 1) / frames
        out["speedup_vs_reference_cuda"] = out["reference_cuda_ms_per_frame"] / out["restir_b200_ms_per_frame"]
        if "errorcheck0_ms_per_frame" in out["reference_cuda"]:
            out["speedup_vs_reference_cuda_errorcheck0"] = out["reference_cuda"]["errorcheck0_ms_per_frame"] / out["restir_b200_ms_per_frame"]
        fr.close()
    # ---- parity: temporal mode (deterministic in the reference), literal radius, frame 3 of the orbit
    pre = os.path.join(tmp, "ref_")
    run_ref("ref_headless_r5", txt, 4, 0, 1, pre, 3)
    fr = sc.frame(W, H)
    prm = rb.default_params(reuse=1)
    for k in range(4):
        cam = base.orbit(k); fr.gbuffer_render(cam); fr.restir_direct(cam, prm, k, 0)
        if k < 3:
            fr.gbuffer_update(cam)
    ref = {n: np.fromfile(pre + n + ".bin", dt).reshape(shape) for n, dt, shape in
           (("matid", np.int32, (P,)), ("motion", np.int32, (P,)), ("depth", np.float32, (P,)), ("normal", np.float32, (P, 3)), ("radiance", np.float32, (P, 3)))}
    mine = {n: fr.read(n) for n in ref}
    par = {}
    par["matid_mismatch_pixels"] = int((ref["matid"] != mine["matid"]).sum())
    par["motion_mismatch_pixels"] = int((ref["motion"] != mine["motion"]).sum())
    d = np.abs(ref["depth"] - mine["depth"]) / np.maximum(np.abs(ref["depth"]), 1e-20)
    same_surface = ref["matid"] == mine["matid"]
    par["depth_max_rel"] = float(d.max()); par["depth_bitexact_fraction"] = float((ref["depth"] == mine["depth"]).mean())
    par["depth_max_rel_where_same_material"] = float(d[same_surface].max())
    par["depth_pixels_rel_gt_1e-5"] = int((d > 1e-5).sum())
    a, b = mine["radiance"].astype(np.float64), ref["radiance"].astype(np.float64)
    rel = np.abs(a - b).sum(1) / np.maximum(np.abs(b).sum(1), 1e-6)
    par["radiance_pixels_within_1e-4_rel"] = float((rel <= 1e-4).mean())
    par["radiance_pixels_bitexact"] = float((mine["radiance"] == ref["radiance"]).all(1).mean())
    par["radiance_mean_relMSE"] = float(np.mean(((a - b) ** 2).sum(1) / (b.sum(1) ** 2 + 1e-3)))
    par["mean_radiance_ref"] = float(b.mean()); par["mean_radiance_b200"] = float(a.mean())
    par["library_fmad"] = os.environ.get("RSTR_LIBNAME", "librestir_b200.so")
    out["parity_vs_reference_cuda_temporal_frame3"] = par
    fr.close(); sc.close()
    return out


def compare_denoisers(work="config2", frames=6):
    """The reference's own filters (denoiser.cu in its CUDA build, REF_DENOISE=1) vs rstr_denoiser_* on the same frames: temporal
    ReSTIR of the orbit, SVGF filtering every frame (history + moments), EAW on the last one."""
    desc, spec, res, reuse, radius = WORKLOADS[work]
    sd = make_scene(spec, res)
    tmp = tempfile.mkdtemp()
    txt = scenes.write_scene_files(sd, tmp, "scene")
    W, H = res
    P = W * H
    pre = os.path.join(tmp, "ref_")
    last = frames - 1
    run_ref("ref_headless_r5", txt, frames, 0, 1, pre, last, denoise=True)
    rb.init(0)
    sc = rb.Scene.from_file(txt)
    base = sc.camera
    fr = sc.frame(W, H)
    eaw, svgf = rb.Denoiser(fr, "eaw"), rb.Denoiser(fr, "svgf")
    prm = rb.default_params(reuse=1)
    for k in range(frames):
        cam = base.orbit(k); fr.gbuffer_render(cam); fr.restir_direct(cam, prm, k, 0)
        svgf.filter(cam)
        if k == last:
            eaw.filter(cam)
        else:
            svgf.next_frame(); fr.gbuffer_update(cam)
    mine = {"radiance": fr.read("radiance"), "eaw": eaw.read()}
    mine["svgf"], mine["svgf_var"] = svgf.read(variance=True)
    ref = {n: np.fromfile(pre + n + ".bin", np.float32).reshape(shape) for n, shape in (("radiance", (P, 3)), ("eaw", (P, 3)), ("svgf", (P, 3)), ("svgf_var", (P,)))}
    out = {"workload": work, "frames": frames, "library": os.environ.get("RSTR_LIBNAME", "librestir_b200.so")}
    for n in ref:
        a, b = mine[n].astype(np.float64).reshape(P, -1), ref[n].astype(np.float64).reshape(P, -1)
        rel = np.abs(a - b).sum(1) / np.maximum(np.abs(b).sum(1), 1e-6 * max(float(np.abs(b).max()), 1e-30))
        out[n] = {"pixels_bitexact": float((mine[n].reshape(P, -1) == ref[n].reshape(P, -1)).all(1).mean()), "pixels_within_1e-4_rel": float((rel <= 1e-4).mean()),
                  "max_rel": float(rel.max()), "mean_ref": float(b.mean()), "mean_b200": float(a.mean())}
    for d in (eaw, svgf):
        d.close()
    fr.close(); sc.close()
    return out


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "denoisers":
        print(json.dumps(compare_denoisers(sys.argv[2] if len(sys.argv) > 2 else "config2", int(sys.argv[3]) if len(sys.argv) > 3 else 6)))
        return
    work = sys.argv[1] if len(sys.argv) > 1 else "config2"
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    print(json.dumps(compare(work, frames, lib_times="parity-only" not in sys.argv)))


if __name__ == "__main__":
    main()
