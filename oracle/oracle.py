"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/restir_oracle.h).

``Oracle("port")`` loads oracle/_build/librestir_oracle.so (the standalone restatement, built on demand
with oracle/Makefile); ``Oracle("reference")`` loads oracle/_ref/libref_harness.so (the reference's own
sources compiled with g++; exists only where /root/reference was available at build time).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "librestir_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref_harness.so")


class OrcCamera(C.Structure):
    _fields_ = [
        ("resolution", C.c_int * 2),
        ("position", C.c_float * 3),
        ("rotation", C.c_float * 3),
        ("view", C.c_float * 3),
        ("up", C.c_float * 3),
        ("right", C.c_float * 3),
        ("fov", C.c_float * 2),
        ("pixelLength", C.c_float * 2),
        ("rotationMatInv", C.c_float * 9),
        ("viewProjection", C.c_float * 16),
        ("lensRadius", C.c_float),
        ("focalDist", C.c_float),
        ("tanFovY", C.c_float),
    ]


assert C.sizeof(OrcCamera) == 196


class OrcParams(C.Structure):
    _fields_ = [
        ("numCandidates", C.c_int),
        ("temporalCap", C.c_int),
        ("numSpatial", C.c_int),
        ("spatialRadius", C.c_float),
        ("reuse", C.c_int),
        ("spatialPasses", C.c_int),
        ("unbiased", C.c_int),
    ]


BUF = dict(albedo=0, normal=1, matid=2, depth=3, motion=4, radiance=5, reservoir=6, reservoir_temp=7, light_index=8)
RESERVOIR_DTYPE = np.dtype([("Li", "<f4", (3,)), ("wi", "<f4", (3,)), ("dist", "<f4"), ("M", "<i4"), ("w", "<f4")])
assert RESERVOIR_DTYPE.itemsize == 36


def build(target: str = "oracle") -> None:
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)


def have_reference() -> bool:
    return os.path.exists(REF_SO)


class Oracle:
    def __init__(self, kind: str = "port"):
        self.kind = kind
        if kind == "port":
            if not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < os.path.getmtime(os.path.join(HERE, "restir_oracle.cpp")):
                build("oracle")
            path = PORT_SO
        elif kind == "reference":
            if not os.path.exists(REF_SO):
                raise FileNotFoundError(REF_SO + " (build with `make -C oracle ref` where /root/reference exists)")
            path = REF_SO
        else:
            raise ValueError(kind)
        L = self.lib = C.CDLL(path)
        vp, ip, fp = C.c_void_p, C.c_int, C.c_float
        L.orc_scene_create.restype = vp
        L.orc_scene_create.argtypes = [ip, vp, vp, vp, vp, ip, vp]
        L.orc_scene_destroy.argtypes = [vp]
        L.orc_scene_set_textures.restype = ip
        L.orc_scene_set_textures.argtypes = [vp, ip, vp, vp, vp, ip]
        L.orc_scene_env_alias.restype = vp
        L.orc_scene_env_alias.argtypes = [vp, vp, vp]
        for n in ("orc_scene_bvh_size", "orc_scene_num_lights", "orc_scene_bvh_depth"):
            getattr(L, n).restype = ip
            getattr(L, n).argtypes = [vp]
        for n in ("orc_scene_boxes", "orc_scene_light_prim_ids", "orc_scene_light_radiance", "orc_scene_alias_table"):
            getattr(L, n).restype = vp
            getattr(L, n).argtypes = [vp]
        L.orc_scene_mtbvh.restype = vp
        L.orc_scene_mtbvh.argtypes = [vp, ip]
        L.orc_scene_sum_light_power.restype = fp
        L.orc_scene_sum_light_power.argtypes = [vp]
        L.orc_camera_update.argtypes = [C.POINTER(OrcCamera)]
        L.orc_frame_create.restype = vp
        L.orc_frame_create.argtypes = [vp, ip, ip]
        L.orc_frame_destroy.argtypes = [vp]
        L.orc_frame_reset.argtypes = [vp]
        L.orc_gbuffer_render.argtypes = [vp, C.POINTER(OrcCamera)]
        L.orc_gbuffer_update.argtypes = [vp, C.POINTER(OrcCamera)]
        L.orc_restir_direct.argtypes = [vp, C.POINTER(OrcCamera), C.POINTER(OrcParams), ip, ip]
        L.orc_pathtrace_direct.argtypes = [vp, C.POINTER(OrcCamera), ip, ip]
        L.orc_frame_buffer.restype = vp
        L.orc_frame_buffer.argtypes = [vp, ip]
        L.orc_set_threads.argtypes = [ip]
        L.orc_get_threads.restype = ip
        L.orc_rng_draws.argtypes = [ip, ip, ip, vp]
        L.orc_intersect.restype = ip
        L.orc_intersect.argtypes = [vp, vp, vp, vp, vp]
        L.orc_occluded.restype = ip
        L.orc_occluded.argtypes = [vp, vp, vp]
        L.orc_last_trace_stats.argtypes = [vp, vp, vp, vp]
        L.orc_alias_build.restype = fp
        L.orc_alias_build.argtypes = [ip, vp, vp]
        L.orc_gi_create.restype = vp
        L.orc_gi_create.argtypes = [vp]
        L.orc_gi_destroy.argtypes = [vp]
        L.orc_restir_indirect.argtypes = [vp, C.POINTER(OrcCamera), ip, ip, ip, ip]
        L.orc_gi_indirect.restype = vp
        L.orc_gi_indirect.argtypes = [vp]
        L.orc_gi_reservoirs.restype = vp
        L.orc_gi_reservoirs.argtypes = [vp]
        if kind == "port":            # the image-space filters are restated in the port only (denoiser.cu's kernels cannot be built by g++)
            L.orc_denoiser_create.restype = vp
            L.orc_denoiser_create.argtypes = [vp, ip]
            L.orc_denoiser_destroy.argtypes = [vp]
            L.orc_denoiser_set_sigmas.argtypes = [vp, fp, fp, fp]
            L.orc_denoiser_filter.argtypes = [vp, C.POINTER(OrcCamera)]
            L.orc_denoiser_next_frame.argtypes = [vp]
            L.orc_denoiser_modulate_albedo.argtypes = [vp]
            L.orc_denoiser_color.restype = vp
            L.orc_denoiser_color.argtypes = [vp]
            L.orc_denoiser_variance.restype = vp
            L.orc_denoiser_variance.argtypes = [vp]
        if kind == "reference":
            L.ref_scene_load_file.restype = vp
            L.ref_scene_load_file.argtypes = [C.c_char_p]
            L.ref_scene_camera.argtypes = [vp, C.POINTER(OrcCamera)]
            L.ref_scene_num_tris.restype = ip
            L.ref_scene_num_tris.argtypes = [vp]
            L.ref_scene_num_materials.restype = ip
            L.ref_scene_num_materials.argtypes = [vp]
            L.ref_scene_array.restype = vp
            L.ref_scene_array.argtypes = [vp, ip]

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _view(ptr, dtype, shape):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        if not ptr or n == 0:
            return np.zeros(shape, dtype)
        buf = (C.c_char * n).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype).reshape(shape).copy()

    def threads(self) -> int:
        return int(self.lib.orc_get_threads())

    def set_threads(self, n: int) -> None:
        self.lib.orc_set_threads(int(n))

    def rng_draws(self, looper: int, index: int, n: int) -> np.ndarray:
        out = np.zeros(n, np.float32)
        self.lib.orc_rng_draws(looper, index, n, out.ctypes.data)
        return out

    def alias_build(self, values):
        v = np.ascontiguousarray(values, np.float32)
        out = np.zeros(v.shape[0], np.dtype([("prob", "<f4"), ("failId", "<i4")]))
        total = self.lib.orc_alias_build(v.shape[0], v.ctypes.data, out.ctypes.data)
        return out, float(total)

    def scene(self, sd) -> "OracleScene":
        return OracleScene(self, sd)


def make_camera(sd, resolution=None) -> OrcCamera:
    """Camera block of the scene file -> Camera POD (scene.cpp:288-355), before update()."""
    cam = OrcCamera()
    res = tuple(resolution or sd.resolution)
    cam.resolution[0], cam.resolution[1] = int(res[0]), int(res[1])
    for i in range(3):
        cam.position[i] = sd.eye[i]
        cam.rotation[i] = sd.rotation[i]
    cam.up[0], cam.up[1], cam.up[2] = 0.0, 1.0, 0.0
    f32 = np.float32
    pi = f32(3.1415926535897932384626422832795028841971)
    fovy = f32(sd.fovy)
    yscaled = f32(np.tan(f32(fovy * f32(pi / f32(180)))))
    xscaled = f32(f32(yscaled * f32(res[0])) / f32(res[1]))
    fovx = f32(f32(np.arctan(xscaled) * f32(180)) / pi)
    cam.fov[0], cam.fov[1] = float(fovx), float(fovy)
    cam.lensRadius = sd.lens_radius
    cam.focalDist = sd.focal_dist
    cam.tanFovY = float(np.tan(np.radians(f32(fovy * f32(0.5)))))
    return cam


class OracleScene:
    def __init__(self, orc: Oracle, sd):
        self.orc, self.sd = orc, sd
        L = orc.lib
        self._keep = [np.ascontiguousarray(sd.vertices, np.float32), np.ascontiguousarray(sd.normals, np.float32),
                      np.ascontiguousarray(sd.texcoords, np.float32), np.ascontiguousarray(sd.material_ids, np.int32),
                      np.ascontiguousarray(sd.materials)]
        v, n, t, m, mats = self._keep
        self.h = L.orc_scene_create(sd.num_tris, v.ctypes.data, n.ctypes.data, t.ctypes.data, m.ctypes.data, len(mats), mats.ctypes.data)
        texs = [np.ascontiguousarray(x, np.float32) for x in getattr(sd, "textures", [])]
        env = int(getattr(sd, "env_map", -1))
        if texs or env >= 0:
            self._keep += texs
            ws = (C.c_int * len(texs))(*[x.shape[1] for x in texs])
            hs = (C.c_int * len(texs))(*[x.shape[0] for x in texs])
            ptrs = (C.c_void_p * len(texs))(*[x.ctypes.data for x in texs])
            if L.orc_scene_set_textures(self.h, len(texs), ws, hs, ptrs, env) != 0:
                raise ValueError("texture id out of range")
        self.bvh_size = L.orc_scene_bvh_size(self.h)
        self.num_lights = L.orc_scene_num_lights(self.h)          # light sampler length (emissive triangles + environment map)
        self.num_emissive = self.num_lights - (1 if env >= 0 else 0)

    def close(self):
        if self.h:
            self.orc.lib.orc_scene_destroy(self.h)
            self.h = None

    def boxes(self):
        return Oracle._view(self.orc.lib.orc_scene_boxes(self.h), np.float32, (self.bvh_size, 6))

    def mtbvh(self, i):
        return Oracle._view(self.orc.lib.orc_scene_mtbvh(self.h, i), np.int32, (self.bvh_size, 3))

    def light_prim_ids(self):
        return Oracle._view(self.orc.lib.orc_scene_light_prim_ids(self.h), np.int32, (self.num_emissive,))

    def light_radiance(self):
        return Oracle._view(self.orc.lib.orc_scene_light_radiance(self.h), np.float32, (self.num_emissive, 3))

    def alias_table(self):
        return Oracle._view(self.orc.lib.orc_scene_alias_table(self.h), np.dtype([("prob", "<f4"), ("failId", "<i4")]), (self.num_lights,))

    def env_alias(self):
        n, total = C.c_int(0), C.c_float(0)
        p = self.orc.lib.orc_scene_env_alias(self.h, C.addressof(n), C.addressof(total))
        return Oracle._view(p, np.dtype([("prob", "<f4"), ("failId", "<i4")]), (n.value,)), float(total.value)

    def sum_light_power(self):
        return float(self.orc.lib.orc_scene_sum_light_power(self.h))

    def bvh_depth(self):
        return int(self.orc.lib.orc_scene_bvh_depth(self.h))

    def intersect(self, origin, direction):
        o = np.asarray(origin, np.float32)
        d = np.asarray(direction, np.float32)
        out = np.zeros(8, np.float32)
        mat = C.c_int(-1)
        prim = self.orc.lib.orc_intersect(self.h, o.ctypes.data, d.ctypes.data, out.ctypes.data, C.addressof(mat))
        return prim, out, mat.value

    def occluded(self, x, y):
        a = np.asarray(x, np.float32)
        b = np.asarray(y, np.float32)
        return bool(self.orc.lib.orc_occluded(self.h, a.ctypes.data, b.ctypes.data))

    def frame(self, w, h) -> "OracleFrame":
        return OracleFrame(self, w, h)


class OracleFrame:
    """Mirrors the reference's frame loop pieces (main.cpp:146-185)."""

    def __init__(self, scene: OracleScene, w: int, h: int):
        self.scene, self.w, self.h = scene, w, h
        self.lib = scene.orc.lib
        self.f = self.lib.orc_frame_create(scene.h, w, h)

    def close(self):
        if self.f:
            self.lib.orc_frame_destroy(self.f)
            self.f = None

    def reset(self):
        self.lib.orc_frame_reset(self.f)

    def gbuffer_render(self, cam):
        self.lib.orc_gbuffer_render(self.f, C.byref(cam))

    def gbuffer_update(self, cam):
        self.lib.orc_gbuffer_update(self.f, C.byref(cam))

    def restir_direct(self, cam, params: OrcParams, looper: int, it: int = 0):
        self.lib.orc_restir_direct(self.f, C.byref(cam), C.byref(params), looper, it)

    def pathtrace_direct(self, cam, looper: int, it: int = 0):
        self.lib.orc_pathtrace_direct(self.f, C.byref(cam), looper, it)

    def trace_stats(self):
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        self.lib.orc_last_trace_stats(self.f, C.addressof(a), C.addressof(b), C.addressof(c))
        return dict(nodes=a.value, tris=b.value, rays=c.value)

    def buffer(self, name: str) -> np.ndarray:
        P = self.w * self.h
        ptr = self.lib.orc_frame_buffer(self.f, BUF[name])
        if name in ("albedo", "normal", "radiance"):
            return Oracle._view(ptr, np.float32, (P, 3))
        if name in ("matid", "motion", "light_index"):
            return Oracle._view(ptr, np.int32, (P,))
        if name == "depth":
            return Oracle._view(ptr, np.float32, (P,))
        return Oracle._view(ptr, RESERVOIR_DTYPE, (P,))


class OracleGI:
    """ReSTIRIndirect (restir.cu:448-476) on an oracle frame: indirect illumination + the indirect reservoirs."""

    def __init__(self, frame: OracleFrame):
        self.frame, self.lib = frame, frame.lib
        self.g = self.lib.orc_gi_create(frame.f)

    def close(self):
        if self.g:
            self.lib.orc_gi_destroy(self.g)
            self.g = None

    def restir_indirect(self, cam, looper: int, it: int = 0, max_depth: int = 3, reuse: int = 1):
        self.lib.orc_restir_indirect(self.g, C.byref(cam), looper, it, max_depth, reuse)

    def indirect(self) -> np.ndarray:
        return Oracle._view(self.lib.orc_gi_indirect(self.g), np.float32, (self.frame.w * self.frame.h, 3))

    def reservoirs(self) -> np.ndarray:
        return Oracle._view(self.lib.orc_gi_reservoirs(self.g), np.float32, (self.frame.w * self.frame.h, 17))


class OracleDenoiser:
    """LeveledEAWFilter ("eaw") / SpatioTemporalFilter ("svgf") of denoiser.h on an oracle frame (port only)."""

    def __init__(self, frame: OracleFrame, kind: str = "eaw"):
        self.frame, self.lib = frame, frame.lib
        self.d = self.lib.orc_denoiser_create(frame.f, {"eaw": 1, "svgf": 2}[kind])

    def close(self):
        if self.d:
            self.lib.orc_denoiser_destroy(self.d)
            self.d = None

    def set_sigmas(self, lumin, normal, depth):
        self.lib.orc_denoiser_set_sigmas(self.d, lumin, normal, depth)

    def filter(self, cam):
        self.lib.orc_denoiser_filter(self.d, C.byref(cam))

    def next_frame(self):
        self.lib.orc_denoiser_next_frame(self.d)

    def modulate_albedo(self):
        self.lib.orc_denoiser_modulate_albedo(self.d)

    def read(self, variance: bool = False):
        P = self.frame.w * self.frame.h
        rgb = Oracle._view(self.lib.orc_denoiser_color(self.d), np.float32, (P, 3))
        return (rgb, Oracle._view(self.lib.orc_denoiser_variance(self.d), np.float32, (P,))) if variance else rgb


def default_params(reuse: int = 0, radius: float = 5.0, k: int = 5, cap: int = 20, candidates: int = 32, passes: int = 1, unbiased: bool = False) -> OrcParams:
    return OrcParams(candidates, cap, k, radius, reuse, passes, 1 if unbiased else 0)


def orbit_camera(orc: Oracle, base: OrcCamera, k: int, speed: float = 2.7, radius: float = 1.0, fps: float = 60.0) -> OrcCamera:
    """runCuda's camera animation with a fixed clock t_k = k*speed/fps (main.cpp:149-162)."""
    cam = OrcCamera.from_buffer_copy(bytes(base))
    t = np.float32(np.float32(k / fps) * np.float32(speed))
    cam.position[0] = float(np.float32(base.position[0]) + np.float32(np.cos(t)) * np.float32(radius))
    cam.position[1] = float(np.float32(base.position[1]) + np.float32(0.0) * np.float32(radius))
    cam.position[2] = float(np.float32(base.position[2]) + np.float32(np.sin(t)) * np.float32(radius))
    orc.lib.orc_camera_update(C.byref(cam))
    return cam
