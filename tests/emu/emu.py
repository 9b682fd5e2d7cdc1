"""TEST INFRASTRUCTURE ONLY: ctypes driver of tests/emu/_build/libkernels_emu.so (the kernels' device code compiled for the host)."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import build as _emu_build  # noqa: E402

from restir_b200 import api  # noqa: E402   (RstrSceneDesc / RstrCamera structures and the host-side camera functions only)


class Emu:
    def __init__(self, sanitize: str = "", small_stack: bool = False):
        L = self.lib = C.CDLL(_emu_build.build(sanitize=sanitize, small_stack=small_stack))
        vp, ip = C.c_void_p, C.c_int
        L.emu_scene_create.restype = vp
        L.emu_scene_create.argtypes = [C.POINTER(api.RstrSceneDesc)]
        L.emu_scene_destroy.argtypes = [vp]
        L.emu_frame_create.restype = vp
        L.emu_frame_create.argtypes = [vp, ip, ip]
        L.emu_frame_destroy.argtypes = [vp]
        L.emu_gbuffer_render.argtypes = [vp, C.POINTER(api.RstrCamera)]
        L.emu_gbuffer_update.argtypes = [vp, C.POINTER(api.RstrCamera)]
        L.emu_gi_reset.argtypes = [vp]
        L.emu_restir_indirect.argtypes = [vp, C.POINTER(api.RstrCamera), ip, ip, ip, ip, ip]
        L.emu_gi_undecided.restype = C.c_ulonglong
        L.emu_gi_undecided.argtypes = [vp]
        L.emu_gi_indirect.restype = vp
        L.emu_gi_indirect.argtypes = [vp]
        L.emu_gi_reservoirs.restype = vp
        L.emu_gi_reservoirs.argtypes = [vp]
        L.emu_frame_plane.restype = vp
        L.emu_frame_plane.argtypes = [vp, ip]
        L.emu_di_create.restype = vp
        L.emu_di_create.argtypes = [vp, ip, ip]
        L.emu_di_destroy.argtypes = [vp]
        L.emu_di_frame.argtypes = [vp, C.POINTER(api.RstrCamera), C.POINTER(api.RstrParams), ip, ip, ip, ip]
        L.emu_di_ptdirect.argtypes = [vp, C.POINTER(api.RstrCamera), ip, ip]
        L.emu_di_update.argtypes = [vp, C.POINTER(api.RstrCamera)]
        L.emu_di_fixup_pixels.restype = C.c_uint
        L.emu_di_fixup_pixels.argtypes = [vp]
        L.emu_di_buffer.restype = vp
        L.emu_di_buffer.argtypes = [vp, ip]
        L.emu_scene_build_traced.restype = ip
        L.emu_scene_build_traced.argtypes = [vp, ip]
        L.emu_di_create_strip.restype = vp
        L.emu_di_create_strip.argtypes = [vp, ip, ip, ip, ip, ip]
        L.emu_di_phase_a.argtypes = [vp, C.POINTER(api.RstrCamera), C.POINTER(api.RstrParams), ip, ip, ip, ip]
        L.emu_di_phase_b_pass.argtypes = [vp, C.POINTER(api.RstrCamera), C.POINTER(api.RstrParams), ip, ip]
        L.emu_di_copy_rows.restype = ip
        L.emu_di_copy_rows.argtypes = [vp, vp, ip, ip, ip]
        L.emu_di_set_bands.argtypes = [vp, ip]
        L.emu_di_halo_miss.restype = C.c_uint
        L.emu_di_halo_miss.argtypes = [vp]
        L.emu_denoiser_create.restype = vp
        L.emu_denoiser_create.argtypes = [vp, ip, C.c_float, C.c_float, C.c_float]
        L.emu_denoiser_destroy.argtypes = [vp]
        L.emu_denoiser_filter.argtypes = [vp, C.POINTER(api.RstrCamera)]
        L.emu_denoiser_next_frame.argtypes = [vp]
        L.emu_denoiser_modulate_albedo.argtypes = [vp]
        L.emu_denoiser_color.restype = vp
        L.emu_denoiser_color.argtypes = [vp]
        L.emu_denoiser_variance.restype = vp
        L.emu_denoiser_variance.argtypes = [vp]

    traced_build = None      # 0 / 1: rebuild the traced tree with the device-side builder's kernels (PLOC / radix tree) after scene creation

    def scene(self, sd):
        v = np.ascontiguousarray(sd.vertices, np.float32)
        n = np.ascontiguousarray(sd.normals, np.float32)
        t = np.ascontiguousarray(sd.texcoords, np.float32)
        m = np.ascontiguousarray(sd.material_ids, np.int32)
        mats = np.ascontiguousarray(sd.materials)
        texs = [np.ascontiguousarray(x, np.float32) for x in getattr(sd, "textures", [])]
        tarr = (api.RstrTexture * max(len(texs), 1))()
        for i, x in enumerate(texs):
            tarr[i] = api.RstrTexture(int(x.shape[1]), int(x.shape[0]), x.ctypes.data)
        desc = api.RstrSceneDesc(int(m.shape[0]), v.ctypes.data, n.ctypes.data, t.ctypes.data, m.ctypes.data, len(mats), mats.ctypes.data,
                                 len(texs), C.cast(tarr, C.c_void_p) if texs else None, int(getattr(sd, "env_map", -1)) + 1)
        h = self.lib.emu_scene_create(C.byref(desc))
        assert h, "emu_scene_create failed"
        if self.traced_build is not None:
            rc = self.lib.emu_scene_build_traced(h, self.traced_build)
            assert rc == 0, "emu_scene_build_traced failed: %d" % rc
        return h

    @staticmethod
    def _view(ptr, dtype, shape):
        n = int(np.prod(shape))
        buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype).reshape(shape).copy()

    def run_gi(self, sd, frames, max_depth=3, reuse=1, accumulate=False, orbit=True, traced_tree=False, gbuffer=False, staged=False):
        """The frame loop of helpers.run_oracle_gi through the emulated device code.  Returns (frames, undecided pixel count)."""
        W, H = sd.resolution
        L = self.lib
        sc = self.scene(sd)
        fr = L.emu_frame_create(sc, W, H)
        base = api.Camera.from_scene(sd)
        out = []
        for f in range(frames):
            cam = base.orbit(f) if orbit else base
            L.emu_gbuffer_render(fr, C.byref(cam))
            L.emu_restir_indirect(fr, C.byref(cam), f, f if accumulate else 0, max_depth, reuse, (2 if staged is True else int(staged) + 1) if staged else 1 if traced_tree else 0)
            r = self._view(L.emu_gi_reservoirs(fr), np.float32, (W * H, 17))
            r[:, 15] = r[:, 15].view(np.int32).astype(np.float32)          # numSamples: int bits in the reference layout
            d = {"indirect": self._view(L.emu_gi_indirect(fr), np.float32, (W * H, 3)), "reservoir": r}
            if gbuffer:
                g = self._view(L.emu_frame_plane(fr, 0), np.float32, (W * H, 4))
                am = self._view(L.emu_frame_plane(fr, 2), np.float32, (W * H, 4))
                d.update(normal=np.ascontiguousarray(g[:, :3]), depth=np.ascontiguousarray(g[:, 3]), matid=self._view(L.emu_frame_plane(fr, 1), np.int32, (W * H,)),
                         albedo=np.ascontiguousarray(am[:, :3]), motion=np.ascontiguousarray(am[:, 3]).view(np.int32))
            out.append(d)
            L.emu_gbuffer_update(fr, C.byref(cam))
        und = int(L.emu_gi_undecided(fr))
        L.emu_frame_destroy(fr)
        L.emu_scene_destroy(sc)
        return out, und

    DI_BUF = {"albedo": (0, np.float32, 3), "normal": (1, np.float32, 3), "matid": (2, np.int32, 0), "depth": (3, np.float32, 0), "motion": (4, np.int32, 0),
              "radiance": (5, np.float32, 3), "reservoir": (6, None, 0), "reservoir_temp": (7, None, 0), "light_index": (8, np.int32, 0)}

    def run_di(self, sd, frames, reuse, radius=5.0, k=5, cap=20, candidates=32, accumulate=False, orbit=True, want=None, passes=1, drain=True, light_index=False, pipeline=0, unbiased=False, ptdirect=False, looper0=0, bands=1):
        """The frame loop of helpers.run_gpu with the staged pipeline's KERNELS run as 32-lane warps on the CPU (emu_di_frame).  Returns
        (frames, fix-up pixels of the last frame)."""
        from oracle.oracle import RESERVOIR_DTYPE

        W, H = sd.resolution
        L = self.lib
        sc = self.scene(sd)
        fr = L.emu_di_create(sc, W, H)
        L.emu_di_set_bands(fr, bands)
        base = api.Camera.from_scene(sd)
        prm = api.default_params(reuse=reuse, radius=radius, k=k, cap=cap, candidates=candidates, passes=passes, unbiased=unbiased)
        names = list(want or ("albedo", "normal", "matid", "depth", "motion", "radiance", "reservoir", "reservoir_temp")) + (["light_index"] if light_index else [])
        out = []
        fix = 0
        for f in range(frames):
            cam = base.orbit(f) if orbit else base
            if ptdirect:
                L.emu_di_ptdirect(fr, C.byref(cam), looper0 + f, f if accumulate else 0)
            else:
                L.emu_di_frame(fr, C.byref(cam), C.byref(prm), looper0 + f, f if accumulate else 0, 1 if drain else 0, pipeline)
            fix = int(L.emu_di_fixup_pixels(fr))
            d = {}
            for n in names:
                which, dt, comps = self.DI_BUF[n]
                ptr = L.emu_di_buffer(fr, which)
                d[n] = self._view(ptr, RESERVOIR_DTYPE, (W * H,)) if dt is None else self._view(ptr, dt, (W * H, comps) if comps else (W * H,))
            out.append(d)
            L.emu_di_update(fr, C.byref(cam))
        L.emu_di_destroy(fr)
        L.emu_scene_destroy(sc)
        return out, fix

    def run_denoiser(self, sd, frames, kind, reuse=1, modulate=False):
        """The loop of tests/test_denoiser.py::gpu_frames with the frame's and the filter's kernels run as warps on the CPU."""
        W, H = sd.resolution
        L = self.lib
        sc = self.scene(sd)
        fr = L.emu_di_create(sc, W, H)
        sig = {"eaw": (64.0, 0.2, 1.0), "svgf": (4.0, 128.0, 1.0)}[kind]          # LeveledEAWFilter::create / SpatioTemporalFilter::create
        dn = L.emu_denoiser_create(fr, {"eaw": 1, "svgf": 2}[kind], *sig)
        base = api.Camera.from_scene(sd)
        prm = api.default_params(reuse=reuse)
        out = []
        for f in range(frames):
            cam = base.orbit(f)
            L.emu_di_frame(fr, C.byref(cam), C.byref(prm), f, 0, 1, 0)
            L.emu_denoiser_filter(dn, C.byref(cam))
            if modulate:
                L.emu_denoiser_modulate_albedo(dn)
            rgb = self._view(L.emu_denoiser_color(dn), np.float32, (W * H, 3))
            var = self._view(L.emu_denoiser_variance(dn), np.float32, (W * H,)) if kind == "svgf" else None
            out.append(dict(rgb=rgb, var=var, radiance=self._view(L.emu_di_buffer(fr, 5), np.float32, (W * H, 3))))
            L.emu_denoiser_next_frame(dn)
            L.emu_di_update(fr, C.byref(cam))
        L.emu_denoiser_destroy(dn)
        L.emu_di_destroy(fr)
        L.emu_scene_destroy(sc)
        return out

    def run_di_strips(self, sd, frames, bounds, halo, reuse=3, radius=5.0, passes=1, names=("matid", "motion", "depth", "radiance", "reservoir", "light_index")):
        """tests/test_gpu_parity.py::test_strips_with_exchanged_gbuffer_halo_and_multi_pass with the kernels as warps: strips that render
        only their own rows, the G-buffer / reservoir halo rows copied between them where the library pushes them over NVLink
        (emu_di_copy_rows = rstr_frame_copy_rows).  Returns per frame {name: the strips' own rows concatenated} and the halo misses."""
        from oracle.oracle import RESERVOIR_DTYPE

        W, H = sd.resolution
        L = self.lib
        sc = self.scene(sd)
        strips = [(L.emu_di_create_strip(sc, W, H, bounds[i], bounds[i + 1], halo), bounds[i], bounds[i + 1]) for i in range(len(bounds) - 1)]
        PL = dict(geom_cur=0, matid_cur=1, resv_history=2, resv_temp=3, resv_temp2=4, resv_out=5)

        def exchange(plane):
            for i, (s, s0, s1) in enumerate(strips):
                for j in (i - 1, i + 1):
                    if 0 <= j < len(strips):
                        d, d0, d1 = strips[j]
                        lo, hi = max(s0, max(0, d0 - halo)), min(s1, min(H, d1 + halo))
                        if lo < hi:
                            assert L.emu_di_copy_rows(d, s, PL[plane], lo, hi) == 0

        base = api.Camera.from_scene(sd)
        prm = api.default_params(reuse=reuse, radius=radius, passes=passes)
        out = []
        for f in range(frames):
            cam = base.orbit(f)
            for s, _, _ in strips:
                L.emu_di_phase_a(s, C.byref(cam), C.byref(prm), f, 0, 1, 0)
            for plane in ("geom_cur", "matid_cur", "resv_temp"):
                exchange(plane)
            for p in range(1, passes + 1):
                for s, _, _ in strips:
                    L.emu_di_phase_b_pass(s, C.byref(cam), C.byref(prm), 0, p)
                if p < passes:
                    exchange("resv_temp2" if p & 1 else "resv_temp")
            exchange("resv_history")
            d = {}
            for n in names:
                which, dt, comps = self.DI_BUF[n]
                parts = []
                for s, s0, s1 in strips:
                    b0 = max(0, s0 - halo)
                    rows = min(H, s1 + halo) - b0
                    ptr = L.emu_di_buffer(s, which)
                    a = self._view(ptr, RESERVOIR_DTYPE, (rows * W,)) if dt is None else self._view(ptr, dt, (rows * W, comps) if comps else (rows * W,))
                    parts.append(a[(s0 - b0) * W:(s1 - b0) * W])
                d[n] = np.concatenate(parts)
            out.append(d)
            for s, _, _ in strips:
                L.emu_di_update(s, C.byref(cam))
        miss = [int(L.emu_di_halo_miss(s)) for s, _, _ in strips]
        for s, _, _ in strips:
            L.emu_di_destroy(s)
        L.emu_scene_destroy(sc)
        return out, miss
