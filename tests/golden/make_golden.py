"""Generates the committed golden fixtures from THE REFERENCE'S OWN CODE (oracle/_ref/libref_harness.so, i.e.
/root/reference/src compiled with g++ by oracle/Makefile).  Run in the build container only:

    make -C oracle ref && python tests/golden/make_golden.py

Outputs (small, committed): tests/golden/*.npz (image fixtures: tests/golden/make_images.py).  The reference ships no golden vectors of its own
(SURVEY.md section 4); these pin the oracle restatement and, through it, the CUDA path.
"""
import ctypes as C
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np

import helpers
from oracle.oracle import Oracle
from restir_b200 import scenes


def textured(ref):
    """6. textures, procedural pattern, normal / metallic / roughness maps, environment map"""
    from oracle.oracle import make_camera
    d = {}
    for name, sd in helpers.textured_scenes().items():
        so = ref.scene(sd)
        d[name + "_alias"] = np.frombuffer(so.alias_table().tobytes(), np.uint8)
        d[name + "_env_alias"] = np.frombuffer(so.env_alias()[0].tobytes(), np.uint8)
        d[name + "_sum_power"] = so.sum_light_power()
        W, H = sd.resolution
        fo = so.frame(W, H)
        cam = make_camera(sd)
        ref.lib.orc_camera_update(C.byref(cam))
        for it in range(2):
            fo.pathtrace_direct(cam, 100 + it, it)
        d[name + "_ptdirect"] = fo.buffer("radiance")
        so.close()
        for mode, reuse, radius in (("ris", 0, 5.0), ("st_r30", 3, 30.0)):
            frames = helpers.run_oracle(ref, sd, 3, reuse, radius=radius)
            for f, bufs in enumerate(frames):
                for n, a in bufs.items():
                    if n == "reservoir_temp" and not (reuse & 2):
                        continue
                    d["%s_%s_f%d_%s" % (name, mode, f, n)] = a.view(np.uint8).reshape(a.shape[0], -1) if a.dtype.fields else a
    np.savez_compressed(os.path.join(HERE, "frames_textured.npz"), **d)


TEXTURE_FILES = ["rgb8.png", "gray8.png", "gray16.png", "rgb16.png", "env_rle.hdr"]   # base colour, metallic, roughness, normal, EnvMap


def textured_scene_file(ref):
    """7. scene-file grammar with texture file names, `BaseColor Procedural`, NormalMap and EnvMap (scene.cpp:389-429, 122-128)"""
    import shutil
    from oracle.oracle import make_camera
    tmp = tempfile.mkdtemp()
    sd = scenes.with_textures(scenes.cornell_box((48, 36), metal_tall_box=True), env=True)
    for f in TEXTURE_FILES:
        shutil.copy(os.path.join(HERE, "images", f), tmp)
    path = scenes.write_scene_files(sd, tmp, "cornell_tex_file", texture_files=TEXTURE_FILES)
    txt = open(path).read().replace(tmp + "/", "")
    open(path, "w").write(txt)
    cwd = os.getcwd()
    os.chdir(tmp)
    rs = ref.lib.ref_scene_load_file(os.path.basename(path).encode())
    os.chdir(cwd)
    T = ref.lib.ref_scene_num_tris(rs)
    L = ref.lib.orc_scene_num_lights(rs)
    ref.lib.ref_scene_texture_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    d = {}
    w, h, e = C.c_int(0), C.c_int(0), C.c_int(0)
    n = ref.lib.ref_scene_texture_info(rs, 0, C.addressof(w), C.addressof(h), C.addressof(e))
    d["num_textures"] = n
    for i in range(n):
        ref.lib.ref_scene_texture_info(rs, i, C.addressof(w), C.addressof(h), C.addressof(e))
        d["texture%d" % i] = Oracle._view(ref.lib.ref_scene_array(rs, 32 + i), np.float32, (h.value, w.value, 3))
        if e.value:
            d["env_map"] = i
    nenv = C.c_int(0); tot = C.c_float(0)
    ref.lib.orc_scene_env_alias.restype = C.c_void_p
    pe = ref.lib.orc_scene_env_alias(C.c_void_p(rs), C.addressof(nenv), C.addressof(tot))
    files = {f: open(os.path.join(tmp, f)).read() for f in sorted(os.listdir(tmp)) if f.endswith((".txt", ".obj"))}
    np.savez_compressed(
        os.path.join(HERE, "scene_file_textured.npz"),
        vertices=Oracle._view(ref.lib.ref_scene_array(rs, 0), np.float32, (3 * T, 3)),
        normals=Oracle._view(ref.lib.ref_scene_array(rs, 1), np.float32, (3 * T, 3)),
        texcoords=Oracle._view(ref.lib.ref_scene_array(rs, 2), np.float32, (3 * T, 2)),
        material_ids=Oracle._view(ref.lib.ref_scene_array(rs, 3), np.int32, (T,)),
        materials=np.frombuffer(Oracle._view(ref.lib.ref_scene_array(rs, 4), np.dtype("V44"), (ref.lib.ref_scene_num_materials(rs),)).tobytes(), np.uint8),
        alias=np.frombuffer(Oracle._view(ref.lib.orc_scene_alias_table(rs), np.dtype("V8"), (L,)).tobytes(), np.uint8),
        env_alias=np.frombuffer(Oracle._view(pe, np.dtype("V8"), (nenv.value,)).tobytes(), np.uint8),
        sum_power=ref.lib.orc_scene_sum_light_power(rs), env_sum=tot.value,
        file_names=np.array(list(files.keys())), file_texts=np.array(list(files.values())), **d)


def multipass(ref):
    """8. two and three spatial passes (restir.cu:201-209, the block the reference ships commented out, driven through
    the reference's own Reservoir::preClampedMerge<4> by the harness)"""
    d = {}
    for name in ("cornell_metal", "gen2000"):
        sd = helpers.test_scenes()[name]
        for passes in (2, 3):
            frames = helpers.run_oracle(ref, sd, 3, 3, radius=12.0, passes=passes, want=("radiance", "reservoir", "reservoir_temp"))
            for f, bufs in enumerate(frames):
                for n, a in bufs.items():
                    d["%s_p%d_f%d_%s" % (name, passes, f, n)] = a.view(np.uint8).reshape(a.shape[0], -1) if a.dtype.fields else a
    np.savez_compressed(os.path.join(HERE, "frames_multipass.npz"), **d)


def ngons(ref):
    """9. OBJ polygons with 5..11 corners (tinyobj's built-in ear clipping, tiny_obj_loader.h:1536-1821): convex, star-shaped,
    clockwise, non-planar, with a duplicated corner, in all three projection-axis classes"""
    rs = np.random.RandomState(5)
    tmp = tempfile.mkdtemp()
    lines, faces, nv = [], [], 0
    for trial in range(24):
        n = rs.randint(5, 12)
        ang = np.sort(rs.rand(n)) * 2 * np.pi
        rad = 0.3 + rs.rand(n) * (1.5 if trial % 3 else 0.2)
        u, v = np.cos(ang) * rad, np.sin(ang) * rad
        if trial % 5 == 0:
            u, v = u[::-1], v[::-1]
        w = rs.randn(n) * (0.0 if trial % 2 else 0.05)
        pts = {0: np.stack([u, v, w], 1), 1: np.stack([w, u, v], 1), 2: np.stack([u, w, v], 1), 3: np.stack([u + v * 0.3, v, u * 0.5 + w], 1)}[trial % 4].astype(np.float32)
        if trial % 7 == 0:
            pts[1] = pts[0]
        pts = pts + np.float32(3.0) * np.array([trial % 6, trial // 6, 0], np.float32)
        lines += ["v %r %r %r" % tuple(float(x) for x in p) for p in pts]
        faces.append("f " + " ".join("%d//%d" % (nv + i + 1, 1 + (i & 1)) for i in range(n)))
        nv += n
    open(os.path.join(tmp, "ngons.obj"), "w").write("\n".join(lines + ["vn 0 0 1", "vn 0 1 0"] + faces) + "\n")
    open(os.path.join(tmp, "ngons.txt"), "w").write(
        "Material m\nType Lambertian\nBaseColor 0.5 0.5 0.5\nMetallic 0\nRoughness 1\nIor 1.5\nNormalMap Null\n\nObject 0\nngons.obj\nMaterial m\n"
        "Translate 0 0 0\nRotate 0 0 0\nScale 1 1 1\n\nCamera\nResolution 32 32\nFovY 20\nLensRadius 0\nFocalDist 1\nApertureMask Null\nSample 1\n"
        "Depth 1\nFile x\nEye 0 0 5\nRotation -90 0 0\nUp 0 1 0\n\nEnvMap Null\n")
    cwd = os.getcwd()
    os.chdir(tmp)
    rs_ = ref.lib.ref_scene_load_file(b"ngons.txt")
    os.chdir(cwd)
    T = ref.lib.ref_scene_num_tris(rs_)
    files = {f: open(os.path.join(tmp, f)).read() for f in ("ngons.txt", "ngons.obj")}
    np.savez_compressed(os.path.join(HERE, "scene_file_ngons.npz"),
                        vertices=Oracle._view(ref.lib.ref_scene_array(rs_, 0), np.float32, (3 * T, 3)),
                        normals=Oracle._view(ref.lib.ref_scene_array(rs_, 1), np.float32, (3 * T, 3)),
                        file_names=np.array(list(files.keys())), file_texts=np.array(list(files.values())))


def image_writers(ref):
    """10. Image::saveJPG / savePNG (image.cpp:41-75) on float images with out-of-range values: the JPEG bytes the reference
    writes (stbi_write_jpg, quality 90) and its PNG files (decoded by the tests)"""
    rs = np.random.RandomState(9)
    ref.lib.ref_image_save.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    tmp = tempfile.mkdtemp()
    d = {}
    for i, (w, h) in enumerate(((1, 1), (9, 7), (16, 16), (37, 21))):
        yy, xx = np.mgrid[0:h, 0:w]
        a = np.clip(rs.randn(h, w, 3) * 0.35 + np.stack([xx / max(w - 1, 1), yy / max(h - 1, 1), 0.5 + 0 * xx], 2), -0.3, 1.4).astype(np.float32)
        base = os.path.join(tmp, "w%d" % i)
        a = np.ascontiguousarray(a)
        ref.lib.ref_image_save(base.encode(), w, h, a.ctypes.data, 1)
        ref.lib.ref_image_save(base.encode(), w, h, a.ctypes.data, 0)
        d["img%d" % i] = a
        d["jpg%d" % i] = np.frombuffer(open(base + ".jpg", "rb").read(), np.uint8)
        d["png%d" % i] = np.frombuffer(open(base + ".png", "rb").read(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "image_writers.npz"), **d)


def gi(ref):
    """8. ReSTIRIndirect (restir.cu:242-416, 448-476): 3 orbit frames, traceDepth 3, temporal reuse on; every frame's
    devIndirectIllum and the last frame's reservoirs (17 floats each: Lo xv nv xs ns, numSamples, weight)"""
    d = {}
    for name, sd in helpers.gi_scenes().items():
        frames = helpers.run_oracle_gi(ref, sd, 3, max_depth=3, reuse=1)
        for f, bufs in enumerate(frames):
            d["%s_f%d_indirect" % (name, f)] = bufs["indirect"]
        d["%s_f2_reservoir" % name] = frames[-1]["reservoir"]
        acc = helpers.run_oracle_gi(ref, sd, 2, max_depth=2, reuse=0, accumulate=True, orbit=False)
        d["%s_acc_indirect" % name] = acc[-1]["indirect"]
    np.savez_compressed(os.path.join(HERE, "gi.npz"), **d)


def main():
    ref = Oracle("reference")
    if "--only-gi" in sys.argv:
        gi(ref)
        return
    if "--only-writers" in sys.argv:
        image_writers(ref)
        return
    if "--only-ngons" in sys.argv:
        ngons(ref)
        return
    if "--only-multipass" in sys.argv:
        multipass(ref)
        return
    if "--only-textured" in sys.argv:
        textured(ref)
        textured_scene_file(ref)
        return
    textured_scene_file(ref)
    textured(ref)
    gi(ref)
    multipass(ref)
    ngons(ref)
    image_writers(ref)
    # 1. RNG, alias known answers
    rng = {"l%d_i%d" % (l, i): ref.rng_draws(l, i, 8) for l, i in ((7, 12345), (0, 0), (59, 2073599), (1023, 8294399))}
    alias, total = ref.alias_build([1, 2, 3, 10])
    alias2, total2 = ref.alias_build(np.linspace(0.1, 7.3, 37) ** 2)
    np.savez_compressed(os.path.join(HERE, "known_answers.npz"), alias_prob=alias["prob"], alias_fail=alias["failId"], alias_sum=total,
                        alias2_prob=alias2["prob"], alias2_fail=alias2["failId"], alias2_sum=total2, **{"rng_" + k: v for k, v in rng.items()})
    # 2. host build outputs
    for name, sd in helpers.test_scenes().items():
        so = ref.scene(sd)
        d = dict(boxes=so.boxes(), light_prim_ids=so.light_prim_ids(), light_radiance=so.light_radiance(),
                 alias_prob=so.alias_table()["prob"], alias_fail=so.alias_table()["failId"], sum_power=so.sum_light_power())
        for i in range(6):
            d["mtbvh%d" % i] = so.mtbvh(i)
        np.savez_compressed(os.path.join(HERE, "host_%s.npz" % name), **d)
        so.close()
    # 3. frame loops: 3 frames, orbiting camera, every reuse mode
    for name, sd in helpers.test_scenes().items():
        d = {}
        for mode, reuse, radius in (("ris", 0, 5.0), ("temporal", 1, 5.0), ("spatial", 2, 5.0), ("st", 3, 5.0), ("st_r30", 3, 30.0)):
            frames = helpers.run_oracle(ref, sd, 3, reuse, radius=radius)
            for f, bufs in enumerate(frames):
                for n, a in bufs.items():
                    if n == "reservoir_temp" and not (reuse & 2):
                        continue
                    d["%s_f%d_%s" % (mode, f, n)] = a.view(np.uint8).reshape(a.shape[0], -1) if a.dtype.fields else a
        np.savez_compressed(os.path.join(HERE, "frames_%s.npz" % name), **d)
    # 4. PTDirect accumulation (2 iterations)
    d = {}
    for name, sd in helpers.test_scenes().items():
        so = ref.scene(sd)
        W, H = sd.resolution
        fo = so.frame(W, H)
        from oracle.oracle import make_camera
        cam = make_camera(sd)
        ref.lib.orc_camera_update(C.byref(cam))
        for it in range(2):
            fo.pathtrace_direct(cam, 100 + it, it)
        d[name] = fo.buffer("radiance")
    np.savez_compressed(os.path.join(HERE, "ptdirect.npz"), **d)
    # 5. scene-file parser + flattening with a rotated / scaled / translated instance
    tmp = tempfile.mkdtemp()
    sd = scenes.cornell_box((48, 36), metal_tall_box=True)
    path = scenes.write_scene_files(sd, tmp, "cornell_file")
    txt = open(path).read()
    # transform the first object; keep the rest at identity
    txt = txt.replace("Translate 0 0 0\nRotate 0 0 0\nScale 1 1 1", "Translate 0.25 -0.125 0.5\nRotate 17.5 -33 8.25\nScale 1.5 0.75 1.125", 1)
    txt = txt.replace(tmp + "/", "")      # relative OBJ paths: resolved against the scene file's directory
    open(path, "w").write(txt)
    cwd = os.getcwd()
    os.chdir(tmp)
    rs = ref.lib.ref_scene_load_file(os.path.basename(path).encode())
    os.chdir(cwd)
    T = ref.lib.ref_scene_num_tris(rs)
    cam = type(make_camera(sd))()
    ref.lib.ref_scene_camera(rs, C.byref(cam))
    files = {f: open(os.path.join(tmp, f)).read() for f in sorted(os.listdir(tmp))}
    np.savez_compressed(
        os.path.join(HERE, "scene_file.npz"),
        vertices=Oracle._view(ref.lib.ref_scene_array(rs, 0), np.float32, (3 * T, 3)),
        normals=Oracle._view(ref.lib.ref_scene_array(rs, 1), np.float32, (3 * T, 3)),
        texcoords=Oracle._view(ref.lib.ref_scene_array(rs, 2), np.float32, (3 * T, 2)),
        material_ids=Oracle._view(ref.lib.ref_scene_array(rs, 3), np.int32, (T,)),
        materials=np.frombuffer(Oracle._view(ref.lib.ref_scene_array(rs, 4), np.dtype("V44"), (ref.lib.ref_scene_num_materials(rs),)).tobytes(), np.uint8),
        camera=np.frombuffer(bytes(cam), np.uint8),
        light_prim_ids=Oracle._view(ref.lib.orc_scene_light_prim_ids(rs), np.int32, (ref.lib.orc_scene_num_lights(rs),)),
        alias=np.frombuffer(Oracle._view(ref.lib.orc_scene_alias_table(rs), np.dtype("V8"), (ref.lib.orc_scene_num_lights(rs),)).tobytes(), np.uint8),
        sum_power=ref.lib.orc_scene_sum_light_power(rs),
        file_names=np.array(list(files.keys())), file_texts=np.array(list(files.values())),
    )
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
