"""Host-side mirror of the reference's frame interface over the C ABI (include/restir_b200.h).

Names follow the reference: ``Scene`` (scene.h:483), ``Camera`` (sceneStructs.h:22), ``GBuffer.render/update``
(gbuffer.h:24-27) and ``ReSTIRDirect`` / ``pathTraceDirect`` (restir.h:132, pathtrace.h:15) are methods of
:class:`Frame`; ``Settings`` toggles (common.h:47-60) are the fields of :class:`RstrParams`.

The CUDA library is the only implementation: if librestir_b200.so is missing it is built with nvcc, and if
that fails (or no CUDA device is usable) every compute call raises :class:`RestirError` -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))


class RestirError(RuntimeError):
    pass


class RstrCamera(C.Structure):
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


assert C.sizeof(RstrCamera) == 196


class RstrParams(C.Structure):
    _fields_ = [
        ("numCandidates", C.c_int),
        ("temporalCap", C.c_int),
        ("numSpatial", C.c_int),
        ("spatialRadius", C.c_float),
        ("reuse", C.c_int),
        ("spatialPasses", C.c_int),
        ("unbiased", C.c_int),
    ]


class RstrSceneDesc(C.Structure):
    _fields_ = [
        ("numTris", C.c_int),
        ("vertices", C.c_void_p),
        ("normals", C.c_void_p),
        ("texcoords", C.c_void_p),
        ("materialIds", C.c_void_p),
        ("numMaterials", C.c_int),
        ("materials", C.c_void_p),
        ("numTextures", C.c_int),
        ("textures", C.c_void_p),
        ("envMap", C.c_int),
    ]


class RstrTexture(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("rgb", C.c_void_p)]


class RstrSceneInfo(C.Structure):
    _fields_ = [
        ("numTris", C.c_int),
        ("numLights", C.c_int),
        ("bvhSize", C.c_int),
        ("bvhDepth", C.c_int),
        ("numMaterials", C.c_int),
        ("sumLightPower", C.c_float),
        ("buildSeconds", C.c_double),
        ("deviceBytes", C.c_size_t),
        ("tracedBvhDepth", C.c_int),
        ("tracedBuildSeconds", C.c_double),
        ("numEmissiveTris", C.c_int),
        ("numTextures", C.c_int),
        ("envWidth", C.c_int),
        ("envHeight", C.c_int),
        ("tracedNodes", C.c_int),
        ("tracedRoot", C.c_int),
    ]


REUSE_NONE, REUSE_TEMPORAL, REUSE_SPATIAL, REUSE_SPATIOTEMPORAL = 0, 1, 2, 3
TONEMAP_NONE, TONEMAP_FILMIC, TONEMAP_ACES = 0, 1, 2

RESERVOIR_DTYPE = np.dtype([("Li", "<f4", (3,)), ("wi", "<f4", (3,)), ("dist", "<f4"), ("M", "<i4"), ("w", "<f4")])
ALIAS_DTYPE = np.dtype([("prob", "<f4"), ("failId", "<i4")])
TRACED_NODE_DTYPE = np.dtype([("lmin", "<f4", (3,)), ("lmax", "<f4", (3,)), ("rmin", "<f4", (3,)), ("rmax", "<f4", (3,)), ("left", "<i4"),
                              ("right", "<i4"), ("pad", "<i4", (2,))])
TRACED_TRI_DTYPE = np.dtype([("v", "<f4", (3, 3)), ("matId", "<i4"), ("prim", "<i4"), ("pad", "<i4")])

SCENE_ARRAYS = dict(boxes=0, mtbvh0=1, light_prim_ids=7, light_radiance=8, alias=9, vertices=10, normals=11,
                    texcoords=12, material_ids=13, materials=14, env_alias=15, traced_nodes=16, traced_tris=17)
FRAME_BUFFERS = dict(albedo=0, normal=1, matid=2, depth=3, motion=4, radiance=5, reservoir=6, reservoir_temp=7,
                     light_index=8, ldr=9)
DENOISER_EAW, DENOISER_SVGF = 1, 2
STAGES = ("gbuffer", "ris", "spatial", "ptdirect", "tonemap")
PLANES = dict(geom_cur=0, matid_cur=1, resv_history=2, resv_temp=3, resv_temp2=4, resv_out=5)

_lib = None


def lib() -> C.CDLL:
    """Load (building if needed) librestir_b200.so.  Raises if the CUDA extension cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    try:
        path = _build.build(loading=True)
    except Exception as e:  # no silent fallback
        raise RestirError("librestir_b200.so is missing and could not be built: %s" % e) from e
    L = C.CDLL(path)
    vp, ip, fp = C.c_void_p, C.c_int, C.c_float
    L.rstr_last_error.restype = C.c_char_p
    L.rstr_build_id.restype = C.c_char_p
    L.rstr_init.argtypes = [ip]
    L.rstr_params_default.argtypes = [C.POINTER(RstrParams)]
    L.rstr_scene_create.argtypes = [C.POINTER(RstrSceneDesc), C.POINTER(vp)]
    L.rstr_scene_load_file.argtypes = [C.c_char_p, C.POINTER(vp), C.POINTER(RstrCamera)]
    L.rstr_scene_destroy.argtypes = [vp]
    L.rstr_scene_info.argtypes = [vp, C.POINTER(RstrSceneInfo)]
    L.rstr_scene_read.argtypes = [vp, ip, vp, C.c_size_t]
    L.rstr_scene_texture_info.argtypes = [vp, ip, vp, vp, vp]
    L.rstr_scene_set_traversal.argtypes = [vp, ip]
    L.rstr_scene_build_traced_gpu.argtypes = [vp, ip, C.POINTER(C.c_float)]
    L.rstr_denoiser_create.argtypes = [vp, ip, C.POINTER(vp)]
    L.rstr_denoiser_destroy.argtypes = [vp]
    L.rstr_denoiser_set_sigmas.argtypes = [vp, fp, fp, fp]
    L.rstr_denoiser_filter.argtypes = [vp, C.POINTER(RstrCamera)]
    L.rstr_denoiser_next_frame.argtypes = [vp]
    L.rstr_denoiser_modulate_albedo.argtypes = [vp]
    L.rstr_denoiser_add_image.argtypes = [vp, vp]
    L.rstr_denoiser_read.argtypes = [vp, vp, vp]
    L.rstr_gi_create.argtypes = [vp, C.POINTER(vp)]
    L.rstr_gi_destroy.argtypes = [vp]
    L.rstr_gi_reset.argtypes = [vp]
    L.rstr_restir_indirect.argtypes = [vp, C.POINTER(RstrCamera), ip, ip, ip, ip, ip]
    L.rstr_gi_read.argtypes = [vp, vp, vp]
    L.rstr_gi_indirect_device.argtypes = [vp, C.POINTER(vp)]
    L.rstr_gi_set_bounce_walk.argtypes = [vp, ip]
    L.rstr_gi_set_pipeline.argtypes = [vp, ip]
    L.rstr_gi_fallback_pixels.argtypes = [vp, C.POINTER(C.c_uint), ip]
    L.rstr_scene_fallback_rays.argtypes = [vp, C.POINTER(C.c_ulonglong), ip]
    L.rstr_camera_update.argtypes = [C.POINTER(RstrCamera)]
    L.rstr_camera_orbit.argtypes = [C.POINTER(RstrCamera), ip, fp, fp, fp, C.POINTER(RstrCamera)]
    L.rstr_frame_create.argtypes = [vp, ip, ip, C.POINTER(vp)]
    L.rstr_frame_create_strip.argtypes = [vp, ip, ip, ip, ip, ip, C.POINTER(vp)]
    L.rstr_frame_destroy.argtypes = [vp]
    L.rstr_frame_reset.argtypes = [vp]
    L.rstr_gbuffer_render.argtypes = [vp, C.POINTER(RstrCamera)]
    L.rstr_gbuffer_update.argtypes = [vp, C.POINTER(RstrCamera)]
    L.rstr_restir_direct.argtypes = [vp, C.POINTER(RstrCamera), C.POINTER(RstrParams), ip, ip]
    L.rstr_restir_phase_a.argtypes = [vp, C.POINTER(RstrCamera), C.POINTER(RstrParams), ip, ip]
    L.rstr_restir_phase_b.argtypes = [vp, C.POINTER(RstrCamera), C.POINTER(RstrParams), ip, ip]
    L.rstr_pathtrace_direct.argtypes = [vp, C.POINTER(RstrCamera), ip, ip]
    L.rstr_restir_phase_b_pass.argtypes = [vp, C.POINTER(RstrCamera), C.POINTER(RstrParams), ip, ip, ip]
    L.rstr_frame_set_halo_render.argtypes = [vp, ip]
    L.rstr_frame_set_fusion.argtypes = [vp, ip]
    L.rstr_frame_set_pipeline.argtypes = [vp, ip]
    L.rstr_frame_set_bands.argtypes = [vp, ip]
    L.rstr_frame_row_cost.argtypes = [vp, ip, vp, ip]
    L.rstr_tonemap.argtypes = [vp, ip, fp]
    L.rstr_frame_save_png.argtypes = [vp, C.c_char_p, ip]
    L.rstr_image_load.argtypes = [C.c_char_p, ip, vp, vp, vp, C.c_size_t]
    L.rstr_image_write_png.argtypes = [C.c_char_p, ip, ip, vp]
    L.rstr_image_write_jpg.argtypes = [C.c_char_p, ip, ip, vp, ip]
    L.rstr_frame_save_jpg.argtypes = [vp, C.c_char_p, ip]
    L.rstr_render_frame_host.argtypes = [vp, C.POINTER(RstrCamera), C.POINTER(RstrParams), ip, ip, ip, vp, C.c_size_t]
    L.rstr_render_frame_host_async.argtypes = [vp, C.POINTER(RstrCamera), C.POINTER(RstrParams), ip, ip, ip, vp, C.c_size_t, ip]
    L.rstr_frame_wait_host.argtypes = [vp, ip]
    L.rstr_frame_sync.argtypes = [vp]
    L.rstr_frame_read.argtypes = [vp, ip, vp, C.c_size_t]
    L.rstr_frame_stage_ms.argtypes = [vp, C.POINTER(fp), ip]
    L.rstr_launch_count.restype = C.c_uint64
    L.rstr_frame_stream.restype = vp
    L.rstr_frame_stream.argtypes = [vp]
    L.rstr_frame_halo_miss.argtypes = [vp, C.POINTER(C.c_uint)]
    L.rstr_frame_halo_miss_reset.argtypes = [vp]
    L.rstr_frame_motion_rows.argtypes = [vp, C.POINTER(C.c_uint), ip]
    L.rstr_frame_plane_row.argtypes = [vp, ip, ip, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.rstr_frame_copy_rows.argtypes = [vp, vp, ip, ip, ip]
    L.rstr_frame_read_device.argtypes = [vp, ip, vp, C.c_size_t]
    L.rstr_frame_mark.argtypes = [vp, ip]
    L.rstr_frame_elapsed_ms.argtypes = [vp, ip, ip, C.POINTER(fp)]
    L.rstr_frame_set_stream.argtypes = [vp, vp]
    L.rstr_strip_group_create.argtypes = [vp, ip, ip, C.POINTER(vp)]
    L.rstr_strip_group_handle.argtypes = [vp, vp]
    L.rstr_strip_group_connect.argtypes = [vp, vp]
    L.rstr_strip_group_frame.argtypes = [vp, C.POINTER(RstrCamera), C.POINTER(RstrParams), ip, ip]
    L.rstr_strip_group_exchange.argtypes = [vp, C.c_uint]
    L.rstr_strip_group_push.argtypes = [vp, C.c_uint]
    L.rstr_strip_group_wait.argtypes = [vp]
    L.rstr_strip_group_ack.argtypes = [vp]
    L.rstr_strip_group_frame_begin.argtypes = [vp, C.POINTER(RstrCamera), C.POINTER(RstrParams), ip, ip]
    L.rstr_strip_group_frame_end.argtypes = [vp, C.POINTER(RstrCamera), C.POINTER(RstrParams), ip, ip]
    L.rstr_strip_group_present.argtypes = [vp, ip, vp, C.c_size_t, ip]
    L.rstr_strip_group_wait_host.argtypes = [vp, ip]
    L.rstr_strip_group_error.argtypes = [vp, C.POINTER(ip)]
    L.rstr_strip_group_destroy.argtypes = [vp]
    L.rstr_host_alloc.restype = vp
    L.rstr_host_alloc.argtypes = [C.c_size_t]
    L.rstr_host_free.argtypes = [vp]
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != 0:
        raise RestirError("restir_b200 error %d: %s" % (rc, lib().rstr_last_error().decode(errors="replace")))


def init(device: int = 0) -> None:
    """Bind this thread to a CUDA device; raises RestirError when none is usable."""
    _check(lib().rstr_init(device))


def default_params(reuse: int = REUSE_TEMPORAL, radius: float = 5.0, k: int = 5, cap: int = 20, candidates: int = 32, passes: int = 1,
                   unbiased: bool = False) -> RstrParams:
    """restir.cu literals (32 candidates, cap 20, 5 neighbours, radius 5 px, one spatial pass) unless overridden."""
    return RstrParams(candidates, cap, k, radius, reuse, passes, 1 if unbiased else 0)


def build_id() -> str:
    return lib().rstr_build_id().decode()


def launch_count() -> int:
    return int(lib().rstr_launch_count())


class Camera(RstrCamera):
    """Camera POD (sceneStructs.h:22-126) + the scene-file camera block (scene.cpp:288-355)."""

    @classmethod
    def from_scene(cls, sd, resolution=None) -> "Camera":
        cam = cls()
        res = tuple(resolution or sd.resolution)
        cam.resolution[0], cam.resolution[1] = int(res[0]), int(res[1])
        for i in range(3):
            cam.position[i] = sd.eye[i]
            cam.rotation[i] = sd.rotation[i]
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
        cam.update()
        return cam

    def copy(self) -> "Camera":
        return Camera.from_buffer_copy(bytes(self))

    def update(self) -> None:
        """Camera::update() (sceneStructs.h:88-102)."""
        _check(lib().rstr_camera_update(C.byref(self)))

    def orbit(self, k: int, speed: float = 2.7, radius: float = 1.0, fps: float = 60.0) -> "Camera":
        """runCuda's camera animation with the fixed clock t_k = k*speed/fps (main.cpp:149-162)."""
        cam = Camera()
        _check(lib().rstr_camera_orbit(C.byref(self), int(k), speed, radius, fps, C.byref(cam)))
        return cam


class Scene:
    """Scene + DevScene (scene.h:64-531): host build (light list, alias table, BVH) and device upload."""

    def __init__(self, handle, camera=None):
        self.h = handle
        self.camera = camera
        info = RstrSceneInfo()
        _check(lib().rstr_scene_info(self.h, C.byref(info)))
        self.info = info

    @classmethod
    def from_arrays(cls, sd) -> "Scene":
        v = np.ascontiguousarray(sd.vertices, np.float32)
        n = np.ascontiguousarray(sd.normals, np.float32)
        t = np.ascontiguousarray(sd.texcoords, np.float32)
        m = np.ascontiguousarray(sd.material_ids, np.int32)
        mats = np.ascontiguousarray(sd.materials)
        texs = [np.ascontiguousarray(x, np.float32) for x in getattr(sd, "textures", [])]
        tarr = (RstrTexture * max(len(texs), 1))()
        for i, x in enumerate(texs):
            tarr[i] = RstrTexture(int(x.shape[1]), int(x.shape[0]), x.ctypes.data)
        desc = RstrSceneDesc(int(m.shape[0]), v.ctypes.data, n.ctypes.data, t.ctypes.data, m.ctypes.data, len(mats), mats.ctypes.data,
                             len(texs), C.cast(tarr, C.c_void_p) if texs else None, int(getattr(sd, "env_map", -1)) + 1)
        h = C.c_void_p()
        _check(lib().rstr_scene_create(C.byref(desc), C.byref(h)))
        return cls(h)

    @classmethod
    def from_file(cls, path: str) -> "Scene":
        """Scene::Scene(filename) + buildDevData (scene.cpp:96, 159)."""
        h = C.c_void_p()
        cam = Camera()
        _check(lib().rstr_scene_load_file(path.encode(), C.byref(h), C.byref(cam)))
        return cls(h, cam)

    def close(self) -> None:
        if self.h:
            lib().rstr_scene_destroy(self.h)
            self.h = None

    def set_traversal(self, exact: bool) -> None:
        """False (default): binned-SAH tree + rank tie-break; True: reference-order walk for every ray."""
        _check(lib().rstr_scene_set_traversal(self.h, 1 if exact else 0))

    def build_traced_gpu(self, mode: int = 0) -> float:
        """Rebuild the traced tree on the device; returns the device time of the build in ms."""
        ms = C.c_float(0.0)
        _check(lib().rstr_scene_build_traced_gpu(self.h, mode, C.byref(ms)))
        return float(ms.value)

    def fallback_rays(self, reset: bool = True) -> dict:
        v = (C.c_ulonglong * 4)()
        _check(lib().rstr_scene_fallback_rays(self.h, v, 1 if reset else 0))
        return dict(gbuffer=int(v[0]), restir_a=int(v[1]), ptdirect=int(v[2]))

    def read(self, name: str, ordering: int = 0) -> np.ndarray:
        T, L, N, E = self.info.numTris, self.info.numLights, self.info.bvhSize, self.info.numEmissiveTris
        spec = {
            "boxes": (np.float32, (N, 6)), "mtbvh": (np.int32, (N, 3)), "light_prim_ids": (np.int32, (E,)),
            "light_radiance": (np.float32, (E, 3)), "alias": (ALIAS_DTYPE, (L,)), "vertices": (np.float32, (3 * T, 3)),
            "env_alias": (ALIAS_DTYPE, (self.info.envWidth * self.info.envHeight,)),
            "traced_nodes": (TRACED_NODE_DTYPE, (self.info.tracedNodes,)), "traced_tris": (TRACED_TRI_DTYPE, (T,)),
            "normals": (np.float32, (3 * T, 3)), "texcoords": (np.float32, (3 * T, 2)), "material_ids": (np.int32, (T,)),
            "materials": (np.dtype("V44"), (self.info.numMaterials,)),
        }[name]
        out = np.zeros(spec[1], spec[0])
        which = SCENE_ARRAYS["mtbvh0"] + ordering if name == "mtbvh" else SCENE_ARRAYS[name]
        _check(lib().rstr_scene_read(self.h, which, out.ctypes.data, out.nbytes))
        return out

    def texture(self, index: int):
        """(array (H, W, 3) f32, is_env_map) of texture ``index`` (Scene::textures)."""
        w, h, e = C.c_int(0), C.c_int(0), C.c_int(0)
        _check(lib().rstr_scene_texture_info(self.h, index, C.byref(w), C.byref(h), C.byref(e)))
        out = np.zeros((h.value, w.value, 3), np.float32)
        _check(lib().rstr_scene_read(self.h, 32 + index, out.ctypes.data, out.nbytes))
        return out, bool(e.value)

    def frame(self, width: int, height: int, rows=None, halo: int = 0) -> "Frame":
        return Frame(self, width, height, rows, halo)


class Frame:
    """G-buffer + reservoirs + radiance of one image (or one horizontal strip of it).

    Mirrors GBuffer::create/destroy (denoiser.cu:373-403), ReSTIRInit/Free/Reset (restir.cu:478-517) and the
    per-frame calls of runCuda (main.cpp:146-185).
    """

    def __init__(self, scene: Scene, width: int, height: int, rows=None, halo: int = 0):
        self.scene, self.w, self.h = scene, width, height
        self.rows = (0, height) if rows is None else (int(rows[0]), int(rows[1]))
        self.halo = halo
        self.f = C.c_void_p()
        _check(lib().rstr_frame_create_strip(scene.h, width, height, self.rows[0], self.rows[1], halo, C.byref(self.f)))
        self.npix = (self.rows[1] - self.rows[0]) * width

    def close(self) -> None:
        if self.f:
            lib().rstr_frame_destroy(self.f)
            self.f = None

    # --- reference-named entry points -------------------------------------------------------------
    def gbuffer_render(self, cam) -> None:       # GBuffer::render
        _check(lib().rstr_gbuffer_render(self.f, C.byref(cam)))

    def gbuffer_update(self, cam) -> None:       # GBuffer::update
        _check(lib().rstr_gbuffer_update(self.f, C.byref(cam)))

    def restir_direct(self, cam, params, looper: int, it: int = 0) -> None:   # ReSTIRDirect
        _check(lib().rstr_restir_direct(self.f, C.byref(cam), C.byref(params), looper, it))

    def restir_phase_a(self, cam, params, looper: int, it: int = 0) -> None:
        _check(lib().rstr_restir_phase_a(self.f, C.byref(cam), C.byref(params), looper, it))

    def restir_phase_b(self, cam, params, looper: int, it: int = 0) -> None:
        _check(lib().rstr_restir_phase_b(self.f, C.byref(cam), C.byref(params), looper, it))

    def restir_phase_b_pass(self, cam, params, looper: int, it: int, p: int) -> None:
        _check(lib().rstr_restir_phase_b_pass(self.f, C.byref(cam), C.byref(params), looper, it, p))

    def set_fusion(self, on: bool) -> None:
        """G-buffer + phase A as one kernel sharing the primary-ray tree walk (default on)."""
        _check(lib().rstr_frame_set_fusion(self.f, 1 if on else 0))

    def set_pipeline(self, staged) -> None:
        """Phase A as the staged kernel pipeline (True), as one fused kernel (False), or chosen by scene size (None, default)."""
        _check(lib().rstr_frame_set_pipeline(self.f, -1 if staged is None else (1 if staged else 0)))

    def set_bands(self, bands: int) -> None:
        """Row bands of the staged pipeline (1 = no overlap between its kernels)."""
        _check(lib().rstr_frame_set_bands(self.f, int(bands)))

    def set_halo_render(self, on: bool) -> None:
        _check(lib().rstr_frame_set_halo_render(self.f, 1 if on else 0))

    def row_cost(self, enable: bool = True, read: bool = False):
        """Start / restart (enable) or stop the per-8-row cycle profile; with ``read`` returns the counters first."""
        out = np.zeros((self.h + 7) // 8, np.float64) if read else None
        _check(lib().rstr_frame_row_cost(self.f, 1 if enable else 0, out.ctypes.data if read else None, len(out) if read else 0))
        return out

    def pathtrace_direct(self, cam, looper: int, it: int = 0) -> None:        # pathTraceDirect
        _check(lib().rstr_pathtrace_direct(self.f, C.byref(cam), looper, it))

    def reset(self) -> None:                     # ReSTIRReset
        _check(lib().rstr_frame_reset(self.f))

    def tonemap(self, mode: int = TONEMAP_ACES, scale: float = 1.0) -> None:  # copyImageToPBO
        _check(lib().rstr_tonemap(self.f, mode, scale))

    def save_png(self, path: str, tonemap: int = TONEMAP_ACES) -> None:        # saveImage (main.cpp:105-144)
        _check(lib().rstr_frame_save_png(self.f, path.encode(), tonemap))

    def save_jpg(self, path: str, tonemap: int = TONEMAP_ACES) -> None:        # saveImage(jpg = true)
        _check(lib().rstr_frame_save_jpg(self.f, path.encode(), tonemap))

    def render_frame_host(self, cam, params, looper: int, it: int = 0, tonemap: int = TONEMAP_ACES, out=None) -> None:
        """One runCuda() frame; ``out`` is a host uint8 buffer of npix*4 bytes receiving the LDR image."""
        ptr, nbytes = (None, 0) if out is None else (out.ctypes.data, out.nbytes)
        p = C.byref(params) if params is not None else None
        _check(lib().rstr_render_frame_host(self.f, C.byref(cam), p, looper, it, tonemap, ptr, nbytes))

    def render_frame_host_async(self, cam, params, looper: int, it: int, tonemap: int, out, slot: int) -> None:
        p = C.byref(params) if params is not None else None
        _check(lib().rstr_render_frame_host_async(self.f, C.byref(cam), p, looper, it, tonemap, out.ctypes.data, out.nbytes, slot))

    def wait_host(self, slot: int) -> None:
        _check(lib().rstr_frame_wait_host(self.f, slot))

    def sync(self) -> None:
        _check(lib().rstr_frame_sync(self.f))

    def stage_ms(self) -> dict:
        ms = (C.c_float * len(STAGES))()
        _check(lib().rstr_frame_stage_ms(self.f, ms, len(STAGES)))
        return {s: float(ms[i]) for i, s in enumerate(STAGES)}

    def halo_miss(self) -> int:
        v = C.c_uint(0)
        _check(lib().rstr_frame_halo_miss(self.f, C.byref(v)))
        return int(v.value)

    def halo_miss_reset(self) -> None:
        _check(lib().rstr_frame_halo_miss_reset(self.f))

    def motion_rows(self, reset: bool = False) -> int:
        """Largest |row(motion) - row| seen by the G-buffer kernels so far: the temporal halo must exceed it."""
        v = C.c_uint(0)
        _check(lib().rstr_frame_motion_rows(self.f, C.byref(v), 1 if reset else 0))
        return int(v.value)

    def stream(self) -> int:
        return int(lib().rstr_frame_stream(self.f) or 0)

    def plane_row(self, plane: str, row: int):
        p, nb = C.c_void_p(), C.c_size_t()
        _check(lib().rstr_frame_plane_row(self.f, PLANES[plane], row, C.byref(p), C.byref(nb)))
        return int(p.value or 0), int(nb.value)

    def mark(self, slot: int) -> None:
        _check(lib().rstr_frame_mark(self.f, slot))

    def elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float(0)
        _check(lib().rstr_frame_elapsed_ms(self.f, a, b, C.byref(ms)))
        return float(ms.value)

    def set_stream(self, stream: int) -> None:
        _check(lib().rstr_frame_set_stream(self.f, C.c_void_p(stream)))

    def copy_rows_from(self, src: "Frame", plane: str, row0: int, row1: int) -> None:
        _check(lib().rstr_frame_copy_rows(self.f, src.f, PLANES[plane], row0, row1))

    def read_into_device(self, name: str, dev_ptr: int, nbytes: int) -> None:
        _check(lib().rstr_frame_read_device(self.f, FRAME_BUFFERS[name], C.c_void_p(dev_ptr), nbytes))

    def read(self, name: str) -> np.ndarray:
        P = self.npix
        if name in ("albedo", "normal", "radiance"):
            out = np.zeros((P, 3), np.float32)
        elif name in ("matid", "motion", "light_index"):
            out = np.zeros((P,), np.int32)
        elif name == "depth":
            out = np.zeros((P,), np.float32)
        elif name == "ldr":
            out = np.zeros((P, 4), np.uint8)
        else:
            out = np.zeros((P,), RESERVOIR_DTYPE)
        _check(lib().rstr_frame_read(self.f, FRAME_BUFFERS[name], out.ctypes.data, out.nbytes))
        return out


STRIP_HANDLE_BYTES = 512


class Denoiser:
    """LeveledEAWFilter (kind "eaw") / SpatioTemporalFilter (kind "svgf") of denoiser.h over a full frame."""

    def __init__(self, frame: "Frame", kind: str = "eaw"):
        self.frame = frame
        self.kind = kind
        self.h = C.c_void_p()
        _check(lib().rstr_denoiser_create(frame.f, {"eaw": DENOISER_EAW, "svgf": DENOISER_SVGF}[kind], C.byref(self.h)))

    def close(self) -> None:
        if self.h:
            lib().rstr_denoiser_destroy(self.h)
            self.h = None

    def set_sigmas(self, lumin: float, normal: float, depth: float) -> None:
        _check(lib().rstr_denoiser_set_sigmas(self.h, lumin, normal, depth))

    def filter(self, cam) -> None:
        _check(lib().rstr_denoiser_filter(self.h, C.byref(cam)))

    def next_frame(self) -> None:
        _check(lib().rstr_denoiser_next_frame(self.h))

    def modulate_albedo(self) -> None:
        _check(lib().rstr_denoiser_modulate_albedo(self.h))

    def add_image(self, other: "Denoiser") -> None:
        _check(lib().rstr_denoiser_add_image(self.h, other.h))

    def read(self, variance: bool = False):
        n = self.frame.w * self.frame.h
        rgb = np.zeros((n, 3), np.float32)
        var = np.zeros(n, np.float32) if variance else None
        _check(lib().rstr_denoiser_read(self.h, rgb.ctypes.data, var.ctypes.data if variance else None))
        return (rgb, var) if variance else rgb


# Reservoir<IndirectLiSample> (restir.h:13-27, 29-117): 68 bytes
GI_RESERVOIR_DTYPE = np.dtype([("Lo", "<f4", (3,)), ("xv", "<f4", (3,)), ("nv", "<f4", (3,)), ("xs", "<f4", (3,)), ("ns", "<f4", (3,)),
                               ("numSamples", "<i4"), ("weight", "<f4")])
assert GI_RESERVOIR_DTYPE.itemsize == 68
GI_TARGET_OWN, GI_TARGET_RADIANCE = 0, 1


class ReSTIRIndirect:
    """ReSTIR GI on a full frame: ReSTIRIndirect (restir.h:133, restir.cu:448-476) with its reservoirs (the indirect share of
    ReSTIRInit / ReSTIRFree / ReSTIRReset).  Call between ``Frame.gbuffer_render`` and ``Frame.gbuffer_update``."""

    def __init__(self, frame: "Frame"):
        self.frame = frame
        self.h = C.c_void_p()
        _check(lib().rstr_gi_create(frame.f, C.byref(self.h)))

    def close(self) -> None:
        if self.h:
            lib().rstr_gi_destroy(self.h)
            self.h = None

    def reset(self) -> None:                     # ReSTIRReset
        _check(lib().rstr_gi_reset(self.h))

    def restir_indirect(self, cam, looper: int, it: int = 0, trace_depth: int = 3, reuse: int = REUSE_TEMPORAL, into_radiance: bool = False) -> None:
        _check(lib().rstr_restir_indirect(self.h, C.byref(cam), looper, it, trace_depth, reuse, GI_TARGET_RADIANCE if into_radiance else GI_TARGET_OWN))

    def read(self) -> np.ndarray:
        """devIndirectIllum (the handle's own plane) as (W*H, 3) float32."""
        out = np.zeros((self.frame.w * self.frame.h, 3), np.float32)
        _check(lib().rstr_gi_read(self.h, out.ctypes.data, None))
        return out

    def read_reservoirs(self) -> np.ndarray:
        """The reservoirs the last call wrote, in the reference's 68-byte layout."""
        out = np.zeros(self.frame.w * self.frame.h, GI_RESERVOIR_DTYPE)
        _check(lib().rstr_gi_read(self.h, None, out.ctypes.data))
        return out

    def set_bounce_walk(self, exact: bool) -> None:
        _check(lib().rstr_gi_set_bounce_walk(self.h, 1 if exact else 0))

    def set_pipeline(self, staged) -> None:
        """rstr_gi_set_pipeline: 0 / False one kernel per frame, 1 / True primary rays / one launch per bounce over the live paths / resolve,
        2 the same with the walks as persistent kernels over ray lists (same bits)."""
        _check(lib().rstr_gi_set_pipeline(self.h, int(staged)))

    def fallback_pixels(self, reset: bool = True) -> int:
        n = C.c_uint(0)
        _check(lib().rstr_gi_fallback_pixels(self.h, C.byref(n), 1 if reset else 0))
        return n.value


class StripGroup:
    """One rank of a multi-GPU strip decomposition (include/restir_b200.h, rstr_strip_group_*): halo rows and the final
    gather travel as peer stores over NVLink between the ranks' frames; the caller only moves the 512-byte handles once."""

    def __init__(self, frame: Frame, rank: int, world: int):
        self.frame, self.rank, self.world = frame, rank, world
        self.g = C.c_void_p()
        _check(lib().rstr_strip_group_create(frame.f, rank, world, C.byref(self.g)))

    def handle(self) -> bytes:
        buf = (C.c_uint8 * STRIP_HANDLE_BYTES)()
        _check(lib().rstr_strip_group_handle(self.g, buf))
        return bytes(buf)

    def connect(self, blobs) -> None:
        """``blobs``: the handles of all ranks, in rank order."""
        data = b"".join(blobs)
        assert len(data) == self.world * STRIP_HANDLE_BYTES
        buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
        _check(lib().rstr_strip_group_connect(self.g, buf))

    def render(self, cam, params, looper: int, it: int = 0) -> None:
        _check(lib().rstr_strip_group_frame(self.g, C.byref(cam), C.byref(params), looper, it))

    def render_begin(self, cam, params, looper: int, it: int = 0) -> None:
        _check(lib().rstr_strip_group_frame_begin(self.g, C.byref(cam), C.byref(params), looper, it))

    def render_end(self, cam, params, looper: int, it: int = 0) -> None:
        _check(lib().rstr_strip_group_frame_end(self.g, C.byref(cam), C.byref(params), looper, it))

    @staticmethod
    def _mask(planes) -> int:
        mask = 0
        for p in planes:
            mask |= 1 << PLANES[p]
        return mask

    def exchange(self, planes) -> None:
        _check(lib().rstr_strip_group_exchange(self.g, self._mask(planes)))

    def push(self, planes) -> None:
        _check(lib().rstr_strip_group_push(self.g, self._mask(planes)))

    def wait(self) -> None:
        _check(lib().rstr_strip_group_wait(self.g))

    def ack(self) -> None:
        _check(lib().rstr_strip_group_ack(self.g))

    def present(self, tonemap: int, out, slot: int) -> None:
        ptr, nbytes = (None, 0) if out is None else (out.ctypes.data, out.nbytes)
        _check(lib().rstr_strip_group_present(self.g, tonemap, ptr, nbytes, slot))

    def wait_host(self, slot: int) -> None:
        _check(lib().rstr_strip_group_wait_host(self.g, slot))

    def error(self) -> bool:
        v = C.c_int(0)
        _check(lib().rstr_strip_group_error(self.g, C.byref(v)))
        return bool(v.value)

    def close(self) -> None:
        if self.g:
            lib().rstr_strip_group_destroy(self.g)
            self.g = None


def load_image(path: str, flip: bool = True) -> np.ndarray:
    """Image::Image(filename) (image.cpp:16-33): PNG / JPEG / Radiance .hdr -> (H, W, 3) f32; host only."""
    w, h = C.c_int(0), C.c_int(0)
    _check(lib().rstr_image_load(path.encode(), 1 if flip else 0, C.addressof(w), C.addressof(h), None, 0))
    out = np.zeros((h.value, w.value, 3), np.float32)
    _check(lib().rstr_image_load(path.encode(), 1 if flip else 0, C.addressof(w), C.addressof(h), out.ctypes.data, out.nbytes))
    return out


def write_png(path: str, rgb: np.ndarray) -> None:
    """Image::savePNG (image.cpp:41-57): (H, W, 3) uint8."""
    a = np.ascontiguousarray(rgb, np.uint8)
    _check(lib().rstr_image_write_png(path.encode(), int(a.shape[1]), int(a.shape[0]), a.ctypes.data))


def write_jpg(path: str, rgb: np.ndarray, quality: int = 90) -> None:
    """Image::saveJPG (image.cpp:59-75): (H, W, 3) uint8 -> the file stbi_write_jpg writes."""
    a = np.ascontiguousarray(rgb, np.uint8)
    _check(lib().rstr_image_write_jpg(path.encode(), int(a.shape[1]), int(a.shape[0]), a.ctypes.data, int(quality)))


def pinned_empty(nbytes: int) -> np.ndarray:
    """Page-locked host buffer (cudaHostAlloc) as a uint8 array; freed with :func:`pinned_free`."""
    p = lib().rstr_host_alloc(nbytes)
    if not p:
        raise RestirError("cudaHostAlloc failed")
    arr = np.frombuffer((C.c_uint8 * nbytes).from_address(p), dtype=np.uint8)
    arr.flags.writeable = True
    return arr


def pinned_free(arr: np.ndarray) -> None:
    lib().rstr_host_free(arr.ctypes.data)
