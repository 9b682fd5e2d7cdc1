"""The reference's image-space filters (denoiser.cu:25-567): LeveledEAWFilter and SpatioTemporalFilter.

CPU: properties of the oracle's restatement (no g++ build of the reference's __global__ bodies exists to pin it: the pin is the
reference's own CUDA binary, tests/test_ref_cuda.py::test_denoisers_against_the_reference_cuda_build).  GPU: the CUDA kernels
against the oracle on the same frames.  Every filter weight goes through expf / powf (libdevice on the GPU, glibc in the oracle),
so colours are compared at 2e-4 relative instead of bit for bit; the discrete skeleton (which pixels pass through, which taps
take part) is the G-buffer's, which is bit-exact (test_gpu_parity.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import helpers  # noqa: E402
from oracle import oracle as orc_mod  # noqa: E402
from restir_b200 import scenes  # noqa: E402


def oracle_frames(orc, sd, frames, kind, reuse=1, modulate=False):
    W, H = sd.resolution
    so = orc.scene(sd)
    fo = so.frame(W, H)
    dn = orc_mod.OracleDenoiser(fo, kind)
    base = orc_mod.make_camera(sd)
    orc.lib.orc_camera_update(base)
    prm = orc_mod.default_params(reuse=reuse)
    out = []
    for f in range(frames):
        cam = orc_mod.orbit_camera(orc, base, f)
        fo.gbuffer_render(cam)
        fo.restir_direct(cam, prm, f, 0)
        dn.filter(cam)
        if modulate:
            dn.modulate_albedo()
        rgb, var = dn.read(variance=True) if kind == "svgf" else (dn.read(), None)
        out.append(dict(rgb=rgb, var=var, radiance=fo.buffer("radiance"), matid=fo.buffer("matid")))
        dn.next_frame()
        fo.gbuffer_update(cam)
    dn.close(); fo.close(); so.close()
    return out


def gpu_frames(rb, sd, frames, kind, reuse=1, modulate=False):
    W, H = sd.resolution
    sc = rb.Scene.from_arrays(sd)
    fr = sc.frame(W, H)
    dn = rb.Denoiser(fr, kind)
    base = rb.Camera.from_scene(sd)
    prm = rb.default_params(reuse=reuse)
    out = []
    for f in range(frames):
        cam = base.orbit(f)
        fr.gbuffer_render(cam)
        fr.restir_direct(cam, prm, f, 0)
        dn.filter(cam)
        if modulate:
            dn.modulate_albedo()
        rgb, var = dn.read(variance=True) if kind == "svgf" else (dn.read(), None)
        out.append(dict(rgb=rgb, var=var, radiance=fr.read("radiance"), matid=fr.read("matid")))
        dn.next_frame()
        fr.gbuffer_update(cam)
    dn.close(); fr.close(); sc.close()
    return out


def test_eaw_filter_properties_on_the_oracle(port_oracle):
    """Pixels without a surface pass through unchanged (denoiser.cu:77-80); on surfaces the filter averages (it lowers the pixel
    noise of the one-sample estimate and keeps the image mean), and its output stays inside the range of its input."""
    sd = scenes.cornell_box((96, 72))
    fr = oracle_frames(port_oracle, sd, 1, "eaw", reuse=0)[0]
    rgb, rad, mat = fr["rgb"].astype(np.float64), fr["radiance"].astype(np.float64), fr["matid"]
    assert np.array_equal(fr["rgb"][mat <= -1], fr["radiance"][mat <= -1])
    surf = mat >= 0
    assert surf.mean() > 0.3
    assert rgb[surf].min() >= rad[surf].min() - 1e-6 and rgb[surf].max() <= rad[surf].max() + 1e-6
    lum = lambda a: a @ np.array([.2126, .7152, .0722])                     # noqa: E731
    img = lum(rad).reshape(72, 96)
    flt = lum(rgb).reshape(72, 96)
    rough = lambda a: np.abs(np.diff(a, axis=1)).mean()                    # noqa: E731
    assert rough(flt) < 0.8 * rough(img)                                    # smoother ...
    assert abs(flt[surf.reshape(72, 96)].mean() - img[surf.reshape(72, 96)].mean()) < 0.1 * img[surf.reshape(72, 96)].mean()   # ... not darker / brighter


def test_svgf_filter_properties_on_the_oracle(port_oracle):
    """First frame: history = input, moments (lum, lum^2, 0) -> spatial variance estimate >= 0 up to rounding; later frames
    accumulate (the filtered variance of a static camera falls), light / sky pixels keep their input."""
    sd = scenes.cornell_box((64, 48))
    fr = oracle_frames(port_oracle, sd, 6, "svgf", reuse=0)
    for f in fr:
        assert np.isfinite(f["rgb"]).all() and np.isfinite(f["var"]).all()
    assert fr[0]["var"].min() > -1e-3
    surf = fr[-1]["matid"] >= 0
    assert fr[-1]["var"][surf].mean() < fr[0]["var"][surf].mean()


pytestmark_gpu = pytest.mark.gpu


def close(a, b, rtol, atol):
    return np.abs(a.astype(np.float64) - b.astype(np.float64)) <= atol + rtol * np.abs(b.astype(np.float64))


@pytest.mark.gpu
@pytest.mark.parametrize("kind,frames", [("eaw", 2), ("svgf", 6)])
@pytest.mark.parametrize("scene", ["cornell", "gen"])
def test_filters_against_oracle(gpu, port_oracle, kind, frames, scene):
    """LeveledEAWFilter::filter / SpatioTemporalFilter::filter through the C ABI vs the oracle, frame after frame of the orbit
    (temporal ReSTIR as input; SVGF crosses its 4-frame switch from the spatial to the temporal variance estimate).  The inputs are
    bit-identical (asserted); the outputs agree within 2e-4 relative (expf / powf: libdevice vs glibc) on >= 99.9 % of the
    pixels -- a weight that underflows on one side only can move a pixel whose other taps are all rejected."""
    sd = scenes.cornell_box((160, 120), metal_tall_box=True) if scene == "cornell" else scenes.procedural(1, 20000, 1000, (160, 90))
    got = gpu_frames(gpu, sd, frames, kind)
    want = oracle_frames(port_oracle, sd, frames, kind)
    for f, (g, w) in enumerate(zip(got, want)):
        assert helpers.mismatches(g["radiance"], w["radiance"]) == 0 and helpers.mismatches(g["matid"], w["matid"]) == 0
        scale = float(np.abs(w["rgb"]).max())
        ok = close(g["rgb"], w["rgb"], 2e-4, 2e-6 * scale).all(1)
        assert ok.mean() >= 0.999, (kind, scene, f, 1.0 - ok.mean())
        if kind == "svgf":
            okv = close(g["var"], w["var"], 1e-3, 1e-6 * float(np.abs(w["var"]).max()) + 1e-9)
            assert okv.mean() >= 0.999, (kind, scene, f, 1.0 - okv.mean())
        sky = w["matid"] <= -1
        assert np.array_equal(g["rgb"][sky], w["rgb"][sky])                 # pass-through pixels are copies


@pytest.mark.gpu
def test_modulate_and_add_against_oracle(gpu, port_oracle):
    """modulateAlbedo (LDRToHDR x albedo, denoiser.cu:218-228) is IEEE arithmetic only: bit-exact on bit-identical input; addImage sums."""
    sd = scenes.cornell_box((96, 72))
    W, H = sd.resolution
    sc = gpu.Scene.from_arrays(sd)
    fr = sc.frame(W, H)
    a, b = gpu.Denoiser(fr, "eaw"), gpu.Denoiser(fr, "eaw")
    a.set_sigmas(1e30, 1e30, 1e30); b.set_sigmas(1e30, 1e30, 1e30)        # all edge-stopping weights exp(-0) = 1: no libm in the result
    cam = gpu.Camera.from_scene(sd).orbit(0)
    fr.gbuffer_render(cam); fr.restir_direct(cam, gpu.default_params(reuse=0), 0, 0)
    a.filter(cam); b.filter(cam)
    flt = a.read()
    a.modulate_albedo()
    mod = a.read()
    a.add_image(b)
    both = a.read()
    assert np.array_equal(both, mod + flt)
    so = port_oracle.scene(sd)
    fo = so.frame(W, H)
    dn = orc_mod.OracleDenoiser(fo, "eaw")
    dn.set_sigmas(1e30, 1e30, 1e30)
    ocam = orc_mod.orbit_camera(port_oracle, (lambda c: (port_oracle.lib.orc_camera_update(c), c)[1])(orc_mod.make_camera(sd)), 0)
    fo.gbuffer_render(ocam); fo.restir_direct(ocam, orc_mod.default_params(reuse=0), 0, 0)
    dn.filter(ocam)
    assert helpers.mismatches(flt, dn.read()) == 0                          # Gaussian-weighted sums in the same order: bit-exact
    dn.modulate_albedo()
    assert helpers.mismatches(mod, dn.read()) == 0
    for x in (a, b):
        x.close()
    dn.close(); fo.close(); so.close(); fr.close(); sc.close()
