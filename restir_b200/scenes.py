"""Deterministic benchmark / test scenes for the ReSTIR DI hot path.

The reference ships no scenes (its ``scenes/`` directory is git-ignored), so the
BASELINE.json configs are realised here:

* :func:`cornell_box`  -- configs 1 and 2 (36 triangles, one quad area light = 2 emissive triangles)
* :func:`procedural`   -- configs 3-5 (``gen(seed, T, L)``: height-field ground + axis-aligned boxes +
  L small downward-facing emissive triangles; PRNG = counter-based splitmix64)

Scenes are returned as flattened WORLD-SPACE triangle soup, i.e. exactly the arrays the reference's
``Scene::buildDevData`` produces (scene.cpp:159-190): ``vertices``/``normals`` (3T x 3 f32),
``texcoords`` (3T x 2 f32), ``materialIds`` (T i32), ``materials`` (44-byte PODs, material.h:258-267),
plus the camera description of the scene file's ``Camera`` block (scene.cpp:288-355).
:func:`write_scene_files` emits the same scene in the reference's text grammar + Wavefront OBJ
 so that the parser path can be exercised too.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

# material.h:113-120
LAMBERTIAN, METALLIC_WORKFLOW, DIELECTRIC, DISNEY, LIGHT = 0, 1, 2, 3, 4
_TYPE_NAMES = {LAMBERTIAN: "Lambertian", METALLIC_WORKFLOW: "MetallicWorkflow", DIELECTRIC: "Dielectric", LIGHT: "Light"}

MATERIAL_DTYPE = np.dtype(
    [
        ("type", "<i4"),
        ("baseColor", "<f4", (3,)),
        ("metallic", "<f4"),
        ("roughness", "<f4"),
        ("ior", "<f4"),
        ("baseColorMapId", "<i4"),
        ("metallicMapId", "<i4"),
        ("roughnessMapId", "<i4"),
        ("normalMapId", "<i4"),
    ]
)
assert MATERIAL_DTYPE.itemsize == 44


def make_materials(specs) -> np.ndarray:
    """specs: list of (type, (r,g,b), metallic, roughness)."""
    m = np.zeros(len(specs), dtype=MATERIAL_DTYPE)
    for i, (t, c, met, rough) in enumerate(specs):
        m[i]["type"] = t
        m[i]["baseColor"] = c
        m[i]["metallic"] = met
        m[i]["roughness"] = rough
        m[i]["ior"] = 1.5
        m[i]["baseColorMapId"] = m[i]["metallicMapId"] = m[i]["roughnessMapId"] = m[i]["normalMapId"] = -1
    return m


@dataclass
class SceneData:
    name: str
    vertices: np.ndarray      # (3T, 3) f32
    normals: np.ndarray       # (3T, 3) f32
    texcoords: np.ndarray     # (3T, 2) f32
    material_ids: np.ndarray  # (T,) i32
    materials: np.ndarray     # MATERIAL_DTYPE
    material_names: list = field(default_factory=list)
    # camera block of the scene file
    eye: tuple = (0.0, 1.0, 3.9)
    rotation: tuple = (-90.0, 0.0, 0.0)   # yaw, pitch, roll (degrees)
    fovy: float = 19.5                    # HALF vertical angle (sceneStructs.h:72)
    resolution: tuple = (800, 800)
    focal_dist: float = 1.0
    lens_radius: float = 0.0
    # textures referenced by the materials' map ids ((H, W, 3) f32 linear RGB each, image.h:7-39) and the
    # environment map (index into textures, -1 = none; scene.cpp:122-128)
    textures: list = field(default_factory=list)
    env_map: int = -1

    @property
    def num_tris(self) -> int:
        return int(self.material_ids.shape[0])

    @property
    def num_lights(self) -> int:
        return int(np.count_nonzero(self.materials["type"][self.material_ids] == LIGHT))


def _face_normals(v: np.ndarray) -> np.ndarray:
    """Per-vertex normal = geometric normal normalize(cross(v1-v0, v2-v0)), fp32."""
    t = v.reshape(-1, 3, 3)
    n = np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0]).astype(np.float32)
    ln = np.sqrt((n * n).sum(1, dtype=np.float32)).astype(np.float32)
    n = (n / np.maximum(ln, np.float32(1e-30))[:, None]).astype(np.float32)
    return np.repeat(n, 3, axis=0)


def _quad(a, b, c, d):
    """Two triangles (a,b,c),(a,c,d); geometric normal = cross(b-a, c-a)."""
    return [a, b, c, a, c, d]


def _box(center, half, yaw_deg):
    """12 triangles of an upright box rotated about +y; outward normals."""
    cx, cy, cz = center
    hx, hy, hz = half
    c, s = np.cos(np.radians(yaw_deg)), np.sin(np.radians(yaw_deg))

    def P(x, y, z):
        return (cx + c * x + s * z, cy + y, cz - s * x + c * z)

    p = [P(sx * hx, sy * hy, sz * hz) for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]
    # index = 4*ix + 2*iy + iz
    def V(ix, iy, iz):
        return p[4 * ix + 2 * iy + iz]

    tris = []
    tris += _quad(V(1, 0, 0), V(1, 1, 0), V(1, 1, 1), V(1, 0, 1))   # +x
    tris += _quad(V(0, 0, 1), V(0, 1, 1), V(0, 1, 0), V(0, 0, 0))   # -x
    tris += _quad(V(0, 1, 0), V(0, 1, 1), V(1, 1, 1), V(1, 1, 0))   # +y
    tris += _quad(V(0, 0, 1), V(0, 0, 0), V(1, 0, 0), V(1, 0, 1))   # -y
    tris += _quad(V(0, 0, 1), V(1, 0, 1), V(1, 1, 1), V(0, 1, 1))   # +z
    tris += _quad(V(1, 0, 0), V(0, 0, 0), V(0, 1, 0), V(1, 1, 0))   # -z
    return tris


def cornell_box(resolution=(800, 800), metal_tall_box: bool = False) -> SceneData:
    """Cornell box, BASELINE.json configs 1/2: 5 wall quads + short/tall boxes + one ceiling quad light.

    Room is [-1,1] x [0,2] x [-1,1], open toward +z; camera at (0,1,3.9) looking down -z
    (yaw -90 deg => view = (0,0,-1), sceneStructs.h:89-94); FovY 19.5 is the HALF angle.
    The light quad sits 0.02 below the ceiling and its winding makes cross(v1-v0, v2-v0) point down
    (lights emit single-sided along that normal, scene.h:414-418).
    """
    mats = make_materials(
        [
            (LAMBERTIAN, (0.73, 0.73, 0.73), 0.0, 1.0),   # 0 white
            (LAMBERTIAN, (0.65, 0.05, 0.05), 0.0, 1.0),   # 1 red
            (LAMBERTIAN, (0.12, 0.45, 0.15), 0.0, 1.0),   # 2 green
            (LIGHT, (17.0, 12.0, 4.0), 0.0, 1.0),         # 3 light (baseColor = radiance)
            (METALLIC_WORKFLOW if metal_tall_box else LAMBERTIAN, (0.73, 0.73, 0.73), 0.8 if metal_tall_box else 0.0, 0.4 if metal_tall_box else 1.0),  # 4 tall box
        ]
    )
    names = ["white", "red", "green", "light", "tall"]
    tris, ids = [], []

    def add(q, m):
        tris.extend(q)
        ids.extend([m] * (len(q) // 3))

    add(_quad((-1, 0, 1), (1, 0, 1), (1, 0, -1), (-1, 0, -1)), 0)     # floor, normal +y
    add(_quad((-1, 2, -1), (1, 2, -1), (1, 2, 1), (-1, 2, 1)), 0)     # ceiling, normal -y
    add(_quad((-1, 0, -1), (1, 0, -1), (1, 2, -1), (-1, 2, -1)), 0)   # back, normal +z
    add(_quad((-1, 0, 1), (-1, 0, -1), (-1, 2, -1), (-1, 2, 1)), 1)   # left (x=-1), normal +x
    add(_quad((1, 0, -1), (1, 0, 1), (1, 2, 1), (1, 2, -1)), 2)       # right (x=+1), normal -x
    add(_quad((-0.25, 1.98, -0.25), (0.25, 1.98, -0.25), (0.25, 1.98, 0.25), (-0.25, 1.98, 0.25)), 3)  # light, normal -y
    add(_box((0.33, 0.3, 0.35), (0.3, 0.3, 0.3), -18.0), 0)            # short box
    add(_box((-0.35, 0.6, -0.3), (0.3, 0.6, 0.3), 20.0), 4)            # tall box
    v = np.asarray(tris, dtype=np.float32)
    return SceneData(
        name="cornell" + ("_metal" if metal_tall_box else ""),
        vertices=v,
        normals=_face_normals(v),
        texcoords=np.zeros((v.shape[0], 2), np.float32),
        material_ids=np.asarray(ids, np.int32),
        materials=mats,
        material_names=names,
        eye=(0.0, 1.0, 3.9),
        rotation=(-90.0, 0.0, 0.0),
        fovy=19.5,
        resolution=tuple(resolution),
    )


# ----------------------------------------------------------------------------- procedural many-light scene
_GOLD = np.uint64(0x9E3779B97F4A7C15)


def synth_textures(seed: int = 7):
    """Five small deterministic textures: base colour 16x8, metallic 4x4, roughness 5x3, normal map 8x8 and an
    environment map 16x8 with a bright 'sun' texel (values from the counter-based splitmix64 stream)."""
    def u(n, stream):
        return np.minimum(_splitmix64(seed, n, stream), 0.999999).astype(np.float32)

    base = (0.1 + 0.8 * u(8 * 16 * 3, 101)).reshape(8, 16, 3)
    yy, xx = np.mgrid[0:8, 0:16]
    base = (base * np.where(((xx // 2 + yy // 2) & 1)[..., None] == 0, 1.0, 0.35)).astype(np.float32)
    metallic = np.repeat(u(16, 102).reshape(4, 4, 1), 3, axis=2).astype(np.float32)
    roughness = np.repeat((0.2 + 0.8 * u(15, 103)).reshape(3, 5, 1), 3, axis=2).astype(np.float32)
    nm = u(8 * 8 * 3, 104).reshape(8, 8, 3)
    normal = np.stack([0.5 + 0.3 * (nm[..., 0] - 0.5), 0.5 + 0.3 * (nm[..., 1] - 0.5), 0.8 + 0.2 * nm[..., 2]], axis=2).astype(np.float32)
    env = (0.05 + 0.6 * u(8 * 16 * 3, 105)).reshape(8, 16, 3).astype(np.float32)
    env[4:, :, :] *= np.float32(0.2)          # dim below the horizon
    env[1, 5] = (40.0, 36.0, 30.0)            # sun
    env[2, 11] = (3.0, 4.0, 6.0)
    return [np.ascontiguousarray(t, np.float32) for t in (base, metallic, roughness, normal, env)]


def with_textures(sd: SceneData, seed: int = 7, env: bool = True, uv_scale: float = 1.7) -> SceneData:
    """Copy of ``sd`` exercising every branch of getTexturedMaterialAndSurface (scene.h:78-99): planar texture
    coordinates (negative and > 1 values included), and the non-emissive materials cycled through base-colour map /
    procedural pattern / metallic+roughness maps (MetallicWorkflow) / normal map / base-colour + normal map; with
    ``env`` also an environment map as the last light (scene.cpp:136-152)."""
    import copy

    out = copy.copy(sd)
    out.name = sd.name + "_textured"
    v = np.asarray(sd.vertices, np.float32)
    a = np.array([0.7, 0.1, 0.3], np.float32) * np.float32(uv_scale)
    b = np.array([-0.2, 0.8, 0.45], np.float32) * np.float32(uv_scale)
    out.texcoords = np.stack([(v * a).sum(1, dtype=np.float32) - np.float32(0.6), (v * b).sum(1, dtype=np.float32) + np.float32(0.25)], axis=1).astype(np.float32)
    out.textures = synth_textures(seed)
    out.env_map = 4 if env else -1
    mats = sd.materials.copy()
    k = 0
    for i in range(len(mats)):
        if mats[i]["type"] in (LIGHT, DIELECTRIC):
            continue
        mode = k % 5
        k += 1
        if mode == 0:
            mats[i]["baseColorMapId"] = 0
        elif mode == 1:
            mats[i]["baseColorMapId"] = -2
        elif mode == 2:
            mats[i]["type"] = METALLIC_WORKFLOW
            mats[i]["metallicMapId"] = 1
            mats[i]["roughnessMapId"] = 2
        elif mode == 3:
            mats[i]["normalMapId"] = 3
        else:
            mats[i]["baseColorMapId"] = 0
            mats[i]["normalMapId"] = 3
    out.materials = mats
    return out


def _splitmix64(seed: int, n: int, stream: int) -> np.ndarray:
    """Counter-based splitmix64: u_i = mix(seed + stream*2^40 + (i+1)*GOLD), as float64 in [0,1)."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) + (np.uint64(stream) << np.uint64(40))) + (np.arange(1, n + 1, dtype=np.uint64) * _GOLD)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def procedural(seed: int = 1, num_tris: int = 200_000, num_lights: int = 10_000, resolution=(1920, 1080)) -> SceneData:
    """``gen(seed, T, L)`` of SURVEY.md 8d config 3/4.

    * height-field ground over [-10,10]^2, 2 triangles per cell, ~60 % of the non-emissive budget;
    * random axis-aligned boxes (12 triangles each) for the rest (leftover < 12 triangles go to the ground);
    * L emissive triangles, edge 0.05-0.15, at y in [2,4], normals biased downward;
    * 5 Lambertian + 1 MetallicWorkflow (roughness 0.4, metallic 0.8) + 8 Light materials.
    Triangle areas stay far above the FLT_EPSILON determinant cull (intersections.h:28) and centroids are
    distinct (coincident centroids degenerate the reference's BVH builder).
    """
    T, L = int(num_tris), int(num_lights)
    assert T > L >= 1
    mats = make_materials(
        [
            (LAMBERTIAN, (0.70, 0.70, 0.70), 0.0, 1.0),
            (LAMBERTIAN, (0.65, 0.25, 0.20), 0.0, 1.0),
            (LAMBERTIAN, (0.25, 0.55, 0.30), 0.0, 1.0),
            (LAMBERTIAN, (0.25, 0.35, 0.70), 0.0, 1.0),
            (LAMBERTIAN, (0.75, 0.65, 0.30), 0.0, 1.0),
            (METALLIC_WORKFLOW, (0.90, 0.85, 0.75), 0.8, 0.4),
        ]
        + [
            (LIGHT, c, 0.0, 1.0)
            for c in [(30, 30, 30), (40, 20, 10), (10, 25, 40), (35, 30, 10), (10, 40, 15), (40, 10, 30), (20, 20, 45), (45, 35, 25)]
        ]
    )
    names = ["lam%d" % i for i in range(5)] + ["metal"] + ["light%d" % i for i in range(8)]
    budget = T - L
    n_box = int(budget * 0.4) // 12
    n_ground = budget - n_box * 12
    # ground grid: nx*nz*2 >= n_ground, then trim is not possible (keeps a regular grid) -> choose nx, nz with nx*nz*2 == n_ground
    if n_ground % 2:
        n_ground -= 1
        extra_light = 1   # keep T exact by one more emissive triangle
    else:
        extra_light = 0
    cells = n_ground // 2
    nx = int(np.floor(np.sqrt(cells)))
    while cells % nx:
        nx -= 1
    nz = cells // nx
    L_eff = L + extra_light

    # --- ground height field
    gx = np.linspace(-10.0, 10.0, nx + 1)
    gz = np.linspace(-10.0, 10.0, nz + 1)
    X, Z = np.meshgrid(gx, gz, indexing="ij")
    ph = _splitmix64(seed, 8, 1) * 6.283185307179586
    Y = (
        0.35 * np.sin(0.9 * X + ph[0]) * np.cos(0.7 * Z + ph[1])
        + 0.15 * np.sin(2.3 * X + 1.7 * Z + ph[2])
        + 0.05 * np.sin(7.1 * X + ph[3]) * np.sin(6.3 * Z + ph[4])
    )
    jitter = (_splitmix64(seed, (nx + 1) * (nz + 1), 2).reshape(nx + 1, nz + 1) - 0.5) * 0.02
    Y = Y + jitter
    P00 = np.stack([X[:-1, :-1], Y[:-1, :-1], Z[:-1, :-1]], -1)
    P10 = np.stack([X[1:, :-1], Y[1:, :-1], Z[1:, :-1]], -1)
    P01 = np.stack([X[:-1, 1:], Y[:-1, 1:], Z[:-1, 1:]], -1)
    P11 = np.stack([X[1:, 1:], Y[1:, 1:], Z[1:, 1:]], -1)
    # winding so that the geometric normal points up (+y): cross(P01-P00, P10-P00) has +y
    g1 = np.stack([P00, P01, P11], -2)
    g2 = np.stack([P00, P11, P10], -2)
    ground = np.stack([g1, g2], 2).reshape(-1, 3, 3)
    ground_mat = (np.floor(_splitmix64(seed, cells, 3) * 5.0).astype(np.int32)).repeat(2)

    # --- boxes
    u = _splitmix64(seed, n_box * 8, 4).reshape(n_box, 8)
    bc = np.stack([u[:, 0] * 18.0 - 9.0, np.zeros(n_box), u[:, 1] * 18.0 - 9.0], -1)
    bh = np.stack([0.05 + u[:, 2] * 0.25, 0.1 + u[:, 3] * 0.6, 0.05 + u[:, 4] * 0.25], -1)
    bc[:, 1] = bh[:, 1] + 0.3 * (u[:, 5] - 0.3)
    box_mat = np.floor(u[:, 6] * 6.0).astype(np.int32)          # 0..5 (5 = metal)
    sgn = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], dtype=np.float64)
    corners = bc[:, None, :] + sgn[None, :, :] * bh[:, None, :]  # (n_box, 8, 3), index 4ix+2iy+iz
    quads = [
        (4, 6, 7, 5), (1, 3, 2, 0), (2, 3, 7, 6), (1, 0, 4, 5), (1, 5, 7, 3), (4, 0, 2, 6),
    ]
    tri_idx = []
    for a, b, c, d in quads:
        tri_idx += [(a, b, c), (a, c, d)]
    tri_idx = np.asarray(tri_idx)                                 # (12, 3)
    boxes = corners[:, tri_idx, :].reshape(-1, 3, 3)
    boxes_mat = box_mat.repeat(12)

    # --- emissive triangles
    u = _splitmix64(seed, L_eff * 12, 5).reshape(L_eff, 12)
    lc = np.stack([u[:, 0] * 19.0 - 9.5, 2.0 + u[:, 1] * 2.0, u[:, 2] * 19.0 - 9.5], -1)
    edge = 0.05 + u[:, 3] * 0.10
    # orthonormal frame around a downward-biased normal
    nrm = np.stack([(u[:, 4] - 0.5) * 0.8, -1.0 + 0.0 * u[:, 5], (u[:, 6] - 0.5) * 0.8], -1)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    ang = u[:, 7] * 6.283185307179586
    t0 = np.cross(nrm, np.array([1.0, 0.0, 0.0]))
    t0 /= np.linalg.norm(t0, axis=1, keepdims=True)
    b0 = np.cross(nrm, t0)
    ca, sa = np.cos(ang)[:, None], np.sin(ang)[:, None]
    tx = t0 * ca + b0 * sa
    bx = -t0 * sa + b0 * ca
    e = edge[:, None]
    v0 = lc - tx * e * 0.5 - bx * e * 0.2887
    v1 = lc + tx * e * 0.5 - bx * e * 0.2887
    v2 = lc + bx * e * 0.5774
    lights = np.stack([v0, v1, v2], 1)
    # make cross(v1-v0, v2-v0) point along nrm (downward); flip winding where it does not
    gn = np.cross(lights[:, 1] - lights[:, 0], lights[:, 2] - lights[:, 0])
    flip = (gn * nrm).sum(1) < 0
    lights[flip] = lights[flip][:, [0, 2, 1], :]
    light_mat = 6 + np.floor(u[:, 8] * 8.0).astype(np.int32)

    tri = np.concatenate([ground, boxes, lights], 0)
    mat_ids = np.concatenate([ground_mat, boxes_mat, light_mat]).astype(np.int32)
    # deterministic shuffle so that emissive triangles are spread through the primitive ids
    perm = np.argsort(_splitmix64(seed, tri.shape[0], 6), kind="stable")
    tri = tri[perm]
    mat_ids = mat_ids[perm]
    assert tri.shape[0] == T, (tri.shape[0], T)
    v = tri.reshape(-1, 3).astype(np.float32)
    return SceneData(
        name="gen_s%d_T%d_L%d" % (seed, T, L_eff),
        vertices=v,
        normals=_face_normals(v),
        texcoords=np.zeros((v.shape[0], 2), np.float32),
        material_ids=mat_ids,
        materials=mats,
        material_names=names,
        eye=(0.0, 5.0, 15.0),
        rotation=(-90.0, -18.0, 0.0),
        fovy=24.0,
        resolution=tuple(resolution),
    )


# ----------------------------------------------------------------------------- scene-file writer (reference grammar)
def write_scene_files(scene: SceneData, out_dir: str, stem: str | None = None, texture_files=None) -> str:
    """Write ``<stem>.txt`` + one OBJ per material in the reference's grammar.

    Triangles are grouped per material (one ``Object`` per material, identity TRS with ``Scale 1 1 1``);
    the flattening order is therefore material-major, which ``scene_file_order`` reproduces.
    """
    os.makedirs(out_dir, exist_ok=True)
    stem = stem or scene.name
    lines = []
    # ``texture_files[i]`` = image file holding scene.textures[i] (material maps / EnvMap are written as file names,
    # scene.cpp:389-429, 122-128); the OBJs then carry ``vt`` records
    tf = texture_files or []
    textured = bool(tf)
    for i, m in enumerate(scene.materials):
        bc, mm, rm, nm = (int(m[k]) for k in ("baseColorMapId", "metallicMapId", "roughnessMapId", "normalMapId"))
        lines += [
            "Material %s" % scene.material_names[i],
            "Type %s" % _TYPE_NAMES[int(m["type"])],
            "BaseColor Procedural" if bc == -2 else ("BaseColor %s" % tf[bc] if bc >= 0 else "BaseColor %r %r %r" % tuple(float(x) for x in m["baseColor"])),
            "Metallic %s" % tf[mm] if mm >= 0 else "Metallic %r" % float(m["metallic"]),
            "Roughness %s" % tf[rm] if rm >= 0 else "Roughness %r" % float(m["roughness"]),
            "Ior %r" % float(m["ior"]),
            "NormalMap %s" % tf[nm] if nm >= 0 else "NormalMap Null",
            "",
        ]
    obj_id = 0
    for i in range(len(scene.materials)):
        sel = np.nonzero(scene.material_ids == i)[0]
        if sel.size == 0:
            continue
        obj_path = os.path.join(out_dir, "%s_m%d.obj" % (stem, i))
        vi = (sel[:, None] * 3 + np.arange(3)[None, :]).reshape(-1)
        with open(obj_path, "w") as f:
            for p in scene.vertices[vi]:
                f.write("v %s %s %s\n" % (repr(float(p[0])), repr(float(p[1])), repr(float(p[2]))))
            for n in scene.normals[vi]:
                f.write("vn %s %s %s\n" % (repr(float(n[0])), repr(float(n[1])), repr(float(n[2]))))
            if textured:
                for t in scene.texcoords[vi]:
                    f.write("vt %s %s\n" % (repr(float(t[0])), repr(float(t[1]))))
            for k in range(sel.size):
                a = 3 * k + 1
                if textured:
                    f.write("f %d/%d/%d %d/%d/%d %d/%d/%d\n" % (a, a, a, a + 1, a + 1, a + 1, a + 2, a + 2, a + 2))
                else:
                    f.write("f %d//%d %d//%d %d//%d\n" % (a, a, a + 1, a + 1, a + 2, a + 2))
        lines += [
            "Object %d" % obj_id,
            obj_path,
            "Material %s" % scene.material_names[i],
            "Translate 0 0 0",
            "Rotate 0 0 0",
            "Scale 1 1 1",
            "",
        ]
        obj_id += 1
    lines += [
        "Camera",
        "Resolution %d %d" % scene.resolution,
        "FovY %r" % float(scene.fovy),
        "LensRadius %r" % float(scene.lens_radius),
        "FocalDist %r" % float(scene.focal_dist),
        "ApertureMask Null",
        "Sample 1",
        "Depth 1",
        "File %s" % stem,
        "Eye %r %r %r" % tuple(float(x) for x in scene.eye),
        "Rotation %r %r %r" % tuple(float(x) for x in scene.rotation),
        "Up 0 1 0",
        "",
        "EnvMap %s" % tf[scene.env_map] if (textured and scene.env_map >= 0) else "EnvMap Null",
        "",
    ]
    path = os.path.join(out_dir, stem + ".txt")
    with open(path, "w") as f:
        f.write("\n".join(lines))
    return path


def scene_file_order(scene: SceneData) -> np.ndarray:
    """Triangle permutation applied by :func:`write_scene_files` (material-major, stable)."""
    return np.argsort(scene.material_ids, kind="stable")
