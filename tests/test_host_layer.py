"""CPU: the product's host layer (C++ behind the C ABI) -- scene build, scene-file parser, camera -- against the
oracle and the golden fixtures; and the C-ABI surface itself.  No GPU: no compute entry point is called."""
import ctypes as C
import os
import re
import tempfile

import numpy as np
import pytest

import helpers
import restir_b200 as rb
from restir_b200 import scenes

G = helpers.GOLDEN
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = rb.lib()
    hdr = open(os.path.join(ROOT, "include", "restir_b200.h")).read()
    names = set(re.findall(r"\b(rstr_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 25
    for n in sorted(names):
        assert hasattr(L, n), "librestir_b200.so does not export %s" % n


def test_struct_layouts():
    assert C.sizeof(rb.api.RstrCamera) == 196 and rb.api.RstrCamera.rotationMatInv.offset == 84 and rb.api.RstrCamera.lensRadius.offset == 184
    assert scenes.MATERIAL_DTYPE.itemsize == 44 and rb.api.RESERVOIR_DTYPE.itemsize == 36
    assert C.sizeof(rb.RstrParams) == 28 and rb.RstrParams.unbiased.offset == 24


@pytest.mark.parametrize("name", ["cornell", "cornell_metal", "gen2000", "five"])
def test_host_build_matches_golden_and_oracle(port_oracle, name):
    sd = helpers.test_scenes()[name]
    g = np.load(os.path.join(G, "host_%s.npz" % name))
    sc = rb.Scene.from_arrays(sd)
    assert sc.info.numTris == sd.num_tris and sc.info.bvhSize == 2 * sd.num_tris - 1
    assert np.array_equal(sc.read("boxes").view(np.uint32), g["boxes"].view(np.uint32))
    for i in range(6):
        assert np.array_equal(sc.read("mtbvh", i), g["mtbvh%d" % i]), "ordering %d" % i
    assert np.array_equal(sc.read("light_prim_ids"), g["light_prim_ids"])
    assert np.array_equal(sc.read("light_radiance"), g["light_radiance"])
    al = sc.read("alias")
    assert np.array_equal(al["prob"].view(np.uint32), g["alias_prob"].view(np.uint32)) and np.array_equal(al["failId"], g["alias_fail"])
    if sc.info.numLights:
        assert sc.info.sumLightPower == float(g["sum_power"])
    so = port_oracle.scene(sd)
    assert sc.info.bvhDepth >= 1 and np.array_equal(sc.read("boxes").view(np.uint32), so.boxes().view(np.uint32))
    sc.close()


def test_host_build_large_matches_oracle(port_oracle):
    sd = scenes.procedural(2, 60000, 3000, (64, 48))
    sc = rb.Scene.from_arrays(sd)
    so = port_oracle.scene(sd)
    assert np.array_equal(sc.read("boxes").view(np.uint32), so.boxes().view(np.uint32))
    for i in (0, 3, 5):
        assert np.array_equal(sc.read("mtbvh", i), so.mtbvh(i))
    assert np.array_equal(sc.read("alias").view(np.uint8), so.alias_table().view(np.uint8))
    sc.close()


def test_camera_update_matches_oracle(port_oracle):
    from oracle.oracle import make_camera, orbit_camera

    for sd in helpers.test_scenes().values():
        cam = rb.Camera.from_scene(sd)
        oc = make_camera(sd)
        port_oracle.lib.orc_camera_update(C.byref(oc))
        assert bytes(cam)[:120] == bytes(oc)[:120] and bytes(cam)[184:] == bytes(oc)[184:]
        for k in (1, 17, 59):
            assert bytes(cam.orbit(k))[:120] == bytes(orbit_camera(port_oracle, oc, k))[:120]


def test_scene_file_parser_matches_reference_fixture():
    g = np.load(os.path.join(G, "scene_file.npz"))
    tmp = tempfile.mkdtemp()
    for n, t in zip(g["file_names"], g["file_texts"]):
        open(os.path.join(tmp, str(n)), "w").write(str(t))
    sc = rb.Scene.from_file(os.path.join(tmp, "cornell_file.txt"))
    for n in ("vertices", "normals", "texcoords", "material_ids"):
        assert np.array_equal(sc.read(n).view(np.uint32), g[n].view(np.uint32)), n
    assert sc.read("materials").tobytes() == g["materials"].tobytes()
    cam = g["camera"].tobytes()
    assert bytes(sc.camera)[:120] == cam[:120] and bytes(sc.camera)[184:] == cam[184:]
    # light tables: scene.cpp:176-180 computes the power of a later instance's emitters from the wrong triangle;
    # the alias table and the power sum must reproduce that
    assert np.array_equal(sc.read("light_prim_ids"), g["light_prim_ids"])
    assert sc.read("alias").tobytes() == g["alias"].tobytes()
    assert sc.info.sumLightPower == float(g["sum_power"])
    sc.close()


def test_scene_file_errors_are_reported_not_fatal():
    tmp = tempfile.mkdtemp()
    with pytest.raises(rb.RestirError):
        rb.Scene.from_file(os.path.join(tmp, "missing.txt"))
    p = os.path.join(tmp, "bad.txt")
    open(p, "w").write("Object 0\nnope.obj\nMaterial Null\nScale 1 1 1\n\n")
    with pytest.raises(rb.RestirError):
        rb.Scene.from_file(p)
    open(p, "w").write("Camera\nResolution 8 8\nFovY 20\nLensRadius 0\nFocalDist 1\nApertureMask Null\nSample 1\nDepth 1\nFile x\nEye 0 0 0\n\n")
    with pytest.raises(rb.RestirError):      # no mesh data (scene.cpp:192-195 exits; here an error code)
        rb.Scene.from_file(p)


def test_bad_arguments():
    sd = helpers.test_scenes()["five"]
    bad = scenes.SceneData(sd.name, sd.vertices, sd.normals, sd.texcoords, np.full(5, 7, np.int32), sd.materials, sd.material_names)
    with pytest.raises(rb.RestirError):
        rb.Scene.from_arrays(bad)


def test_scene_generators_are_deterministic():
    a, b = scenes.procedural(1, 5000, 300), scenes.procedural(1, 5000, 300)
    assert np.array_equal(a.vertices, b.vertices) and np.array_equal(a.material_ids, b.material_ids)
    assert a.num_tris == 5000 and a.num_lights == 300
    c = scenes.cornell_box()
    assert c.num_tris == 36 and c.num_lights == 2


def _build_cpp_example(tmp):
    import subprocess

    exe = os.path.join(tmp, "headless")
    cmd = ["/usr/bin/g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "examples"),
           os.path.join(ROOT, "examples", "headless_main.cpp"), "-L" + os.path.join(ROOT, "restir_b200"), "-lrestir_b200",
           "-Wl,-rpath," + os.path.join(ROOT, "restir_b200"), "-o", exe]
    subprocess.run(cmd, check=True)
    return exe


def test_cpp_host_example_links_against_the_c_abi():
    """The reference-named C++ shim (examples/restir_shim.hpp) + a runCuda()-style driver compile with plain g++ and
    link against the C ABI: host code stays C++ and needs neither CUDA headers nor nvcc."""
    rb.lib()
    exe = _build_cpp_example(tempfile.mkdtemp())
    assert os.path.exists(exe)


@pytest.mark.parametrize("name", ["cornell_tex", "gen2000_tex"])
def test_textured_scene_host_tables(name):
    """Environment map: envMapSampler over the texels and the light sampler with the map appended as its last entry
    (scene.cpp:136-157), against the reference-generated fixture."""
    sd = helpers.textured_scenes()[name]
    g = np.load(os.path.join(helpers.GOLDEN, "frames_textured.npz"))
    sc = rb.Scene.from_arrays(sd)
    assert sc.info.numLights == sc.info.numEmissiveTris + 1 and (sc.info.envWidth, sc.info.envHeight) == (16, 8)
    assert sc.info.numTextures == 5
    assert np.array_equal(np.frombuffer(sc.read("alias").tobytes(), np.uint8), g[name + "_alias"])
    assert np.array_equal(np.frombuffer(sc.read("env_alias").tobytes(), np.uint8), g[name + "_env_alias"])
    assert sc.info.sumLightPower == float(g[name + "_sum_power"])
    sc.close()


def test_texture_ids_are_validated():
    sd = helpers.textured_scenes()["cornell_tex"]
    sd.materials = sd.materials.copy()
    sd.materials[0]["normalMapId"] = 9
    with pytest.raises(rb.RestirError):
        rb.Scene.from_arrays(sd)
    sd.materials[0]["normalMapId"] = -2           # the procedural id is only valid for the base colour (scene.h:80-97)
    with pytest.raises(rb.RestirError):
        rb.Scene.from_arrays(sd)


def test_image_loader_matches_reference_loader():
    """PNG (every colour type / bit depth / row filter, split IDAT, stored / fixed / dynamic deflate blocks), JPEG (baseline /
    progressive, five chroma layouts, restart markers, grey, RGB-tagged) and Radiance
    .hdr (RLE and flat) against what the reference's loader (stb_image via Image::Image) returned, bit for bit."""
    g = np.load(os.path.join(helpers.GOLDEN, "images.npz"))
    assert len(g.files) == 84
    for key in g.files:
        name, flip = key.rsplit("_flip", 1)
        a = rb.load_image(os.path.join(helpers.GOLDEN, "images", name), bool(int(flip)))
        assert a.shape == g[key].shape and np.array_equal(a.view(np.uint32), g[key].view(np.uint32)), key
    tmp = tempfile.mkdtemp()
    bad = os.path.join(tmp, "bad.png")
    open(bad, "wb").write(open(os.path.join(helpers.GOLDEN, "images", "rgb8.png"), "rb").read()[:200])
    with pytest.raises(rb.RestirError):
        rb.load_image(bad)
    with pytest.raises(rb.RestirError):
        rb.load_image(os.path.join(tmp, "missing.hdr"))


def test_png_writer_round_trip():
    """Image::savePNG replacement: a valid PNG (chunk CRCs, zlib stream with Adler-32, per-row filters, fixed-Huffman LZ77
    deflate) that decodes to the same bytes with the library's own loader, with zlib + a textbook unfilter, and shrinks
    smooth images."""
    import struct
    import zlib

    def decode(path):
        data = open(path, "rb").read()
        assert data[:8] == b"\x89PNG\r\n\x1a\n"
        pos, idat, dims = 8, b"", None
        while pos < len(data):
            n, tag = struct.unpack(">I4s", data[pos:pos + 8])
            body = data[pos + 8:pos + 8 + n]
            assert struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(tag + body) & 0xFFFFFFFF, tag
            if tag == b"IHDR":
                w, h, depth, ctype, comp, flt, lace = struct.unpack(">IIBBBBB", body)
                assert (depth, ctype, comp, flt, lace) == (8, 2, 0, 0, 0)
                dims = (w, h)
            if tag == b"IDAT":
                idat += body
            pos += 12 + n
        w, h = dims
        raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 3 * w).astype(np.int64)
        out = np.zeros((h, 3 * w), np.int64)
        for y in range(h):
            ft, row = raw[y, 0], raw[y, 1:]
            for i in range(3 * w):
                a = out[y, i - 3] if i >= 3 else 0
                b = out[y - 1, i] if y else 0
                c = out[y - 1, i - 3] if (y and i >= 3) else 0
                p = a + b - c
                pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                pred = (0, a, b, (a + b) >> 1, a if (pa <= pb and pa <= pc) else (b if pb <= pc else c))[ft]
                out[y, i] = (row[i] + pred) & 255
        return out.reshape(h, w, 3).astype(np.uint8)

    rs = np.random.RandomState(3)
    tmp = tempfile.mkdtemp()
    yy, xx = np.mgrid[0:61, 0:97]
    smooth = np.stack([xx * 255 // 96, yy * 255 // 60, (xx + yy) * 255 // 156], 2).astype(np.uint8)
    for name, img in (("noise", rs.randint(0, 256, (23, 31, 3)).astype(np.uint8)), ("smooth", smooth), ("one", np.full((1, 1, 3), 200, np.uint8)),
                      ("flat", np.full((40, 300, 3), 17, np.uint8))):
        path = os.path.join(tmp, name + ".png")
        rb.write_png(path, img)
        assert np.array_equal(rb.load_image(path, flip=False), img.astype(np.float32) / np.float32(255)), name
        assert np.array_equal(decode(path), img), name
    assert os.path.getsize(os.path.join(tmp, "smooth.png")) < smooth.nbytes // 4
    assert os.path.getsize(os.path.join(tmp, "flat.png")) < 1000


def test_textured_scene_file_matches_reference_parser():
    """Material blocks naming texture files, `BaseColor Procedural`, NormalMap and EnvMap (scene.cpp:389-429, 122-128):
    texture ids in order of first use, material textures flipped vertically and the environment map not, OBJ `vt`
    records, and the light sampler with the environment map as last entry -- against the reference's own parser."""
    import shutil
    g = np.load(os.path.join(G, "scene_file_textured.npz"))
    tmp = tempfile.mkdtemp()
    for n, t in zip(g["file_names"], g["file_texts"]):
        open(os.path.join(tmp, str(n)), "w").write(str(t))
    for f in os.listdir(os.path.join(G, "images")):
        shutil.copy(os.path.join(G, "images", f), tmp)
    sc = rb.Scene.from_file(os.path.join(tmp, "cornell_tex_file.txt"))
    for n in ("vertices", "normals", "texcoords", "material_ids"):
        assert np.array_equal(sc.read(n).view(np.uint32), g[n].view(np.uint32)), n
    assert sc.read("materials").tobytes() == g["materials"].tobytes()
    assert sc.info.numTextures == int(g["num_textures"])
    for i in range(sc.info.numTextures):
        tex, is_env = sc.texture(i)
        assert np.array_equal(tex.view(np.uint32), g["texture%d" % i].view(np.uint32)), i
        assert is_env == (i == int(g["env_map"]))
    assert sc.read("alias").tobytes() == g["alias"].tobytes()
    assert sc.read("env_alias").tobytes() == g["env_alias"].tobytes()
    assert sc.info.sumLightPower == float(g["sum_power"])
    sc.close()


@pytest.mark.parametrize("name", ["cornell", "five", "gen2000"])
def test_traced_tree_is_a_valid_bvh(name):
    """The binned-SAH tree the kernels walk (bvh_fast.cpp): every triangle sits in exactly one leaf (<= 4 per leaf) and
    every child box (padded) contains the triangles below it, so a conservative slab walk cannot miss a hit."""
    sd = helpers.test_scenes()[name]
    sc = rb.Scene.from_arrays(sd)
    nodes, tris = sc.read("traced_nodes"), sc.read("traced_tris")
    T = sc.info.numTris
    assert sorted(tris["prim"].tolist()) == list(range(T))
    assert np.array_equal(tris["v"].reshape(T, 9), np.asarray(sd.vertices, np.float32).reshape(T, 9)[tris["prim"]])
    seen = np.zeros(T, int)
    stack = [(sc.info.tracedRoot, None)]
    while stack:
        ref, box = stack.pop()
        if ref < 0:
            u = ref & 0xFFFFFFFF
            first, cnt = u & 0x07FFFFFF, ((u >> 27) & 7) + 1
            assert cnt <= 4
            seen[first:first + cnt] += 1
            if box is not None:
                v = tris["v"][first:first + cnt].reshape(-1, 3)
                assert (v >= box[0]).all() and (v <= box[1]).all()
            continue
        n = nodes[ref]
        for lo, hi, child in ((n["lmin"], n["lmax"], n["left"]), (n["rmin"], n["rmax"], n["right"])):
            if box is not None:
                assert (lo >= box[0] - 1e-4).all() and (hi <= box[1] + 1e-4).all()
            stack.append((int(child), (lo, hi)))
    assert (seen == 1).all()
    sc.close()


def test_image_loader_rejects_damaged_files_without_crashing():
    """Truncated / bit-flipped PNG and HDR files (texture files are external input): an error, or a decoded image, never a
    crash or an allocation sized by a lying header (scripts/fuzz_image_loader.py runs the long version)."""
    import random
    rnd = random.Random(7)
    src = os.path.join(helpers.GOLDEN, "images")
    files = [open(os.path.join(src, f), "rb").read() for f in sorted(os.listdir(src))]
    p = os.path.join(tempfile.mkdtemp(), "f.bin")
    outcomes = {"ok": 0, "rejected": 0}
    for it in range(400):
        b = bytearray(rnd.choice(files))
        if it % 3 == 0:
            b = b[:rnd.randrange(0, len(b))]
        else:
            for _ in range(rnd.randrange(1, 6)):
                b[rnd.randrange(len(b))] = rnd.randrange(256)
        open(p, "wb").write(b)
        try:
            a = rb.load_image(p, bool(it & 1))
            assert a.ndim == 3 and a.shape[2] == 3
            outcomes["ok"] += 1
        except rb.RestirError:
            outcomes["rejected"] += 1
    assert outcomes["ok"] > 0 and outcomes["rejected"] > 0
    open(p, "wb").write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 100000 +X 100000\n" + b"\x02\x02\x01\x86" * 10)
    with pytest.raises(rb.RestirError):
        rb.load_image(p)


def test_obj_polygons_are_split_like_the_reference_loader():
    """OBJ faces with 5-11 corners: the reference's tinyobj ear-clips them (tiny_obj_loader.h:1536-1821); the corner order it
    emits fixes the flattened primitive order.  24 polygons (convex, concave, clockwise, non-planar, duplicated corner, all
    projection-axis classes) against the reference parser's output."""
    g = np.load(os.path.join(G, "scene_file_ngons.npz"))
    tmp = tempfile.mkdtemp()
    for n, t in zip(g["file_names"], g["file_texts"]):
        open(os.path.join(tmp, str(n)), "w").write(str(t))
    sc = rb.Scene.from_file(os.path.join(tmp, "ngons.txt"))
    assert sc.info.numTris == g["vertices"].shape[0] // 3 > 24 * 3
    assert np.array_equal(sc.read("vertices").view(np.uint32), g["vertices"].view(np.uint32))
    assert np.array_equal(sc.read("normals").view(np.uint32), g["normals"].view(np.uint32))
    sc.close()


def test_image_writers_match_the_reference():
    """saveImage's file formats: the JPEG writer produces the bytes Image::saveJPG (stbi_write_jpg, quality 90) wrote for
    the same pixels, and the PNG files the reference wrote decode to the reference's 8-bit conversion
    (clamp(pix, 0, 1) * 255, truncated: image.cpp:45-49) -- so `rstr_frame_save_jpg / _png` are drop-ins for saveImage."""
    g = np.load(os.path.join(G, "image_writers.npz"))
    tmp = tempfile.mkdtemp()
    for i in range(4):
        a = g["img%d" % i]
        u8 = (np.clip(a, 0, 1) * np.float32(255)).astype(np.uint8)
        p = os.path.join(tmp, "o%d.jpg" % i)
        rb.write_jpg(p, u8, 90)
        assert open(p, "rb").read() == g["jpg%d" % i].tobytes(), i
        assert np.array_equal(rb.load_image(p, flip=False).shape, a.shape)           # and it is a readable JPEG
        q = os.path.join(tmp, "r%d.png" % i)
        open(q, "wb").write(g["png%d" % i].tobytes())
        assert np.array_equal(rb.load_image(q, flip=False), u8.astype(np.float32) / np.float32(255)), i
    # quality 0 = default 90; other qualities give valid files of monotone size
    sizes = []
    for quality in (10, 50, 90, 100):
        p = os.path.join(tmp, "q%d.jpg" % quality)
        rb.write_jpg(p, (np.clip(g["img3"], 0, 1) * 255).astype(np.uint8), quality)
        assert rb.load_image(p, flip=False).shape == g["img3"].shape
        sizes.append(os.path.getsize(p))
    assert sizes == sorted(sizes)
