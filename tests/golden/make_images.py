"""Writes the small image fixtures (PNG of every supported colour type / bit depth / row filter, Radiance .hdr flat and
RLE, JPEG baseline / progressive with the usual chroma layouts) and records what THE REFERENCE'S loader (stb_image through Image::Image, oracle/_ref/libref_harness.so) returns
for each, flipped and unflipped -> tests/golden/images/*, tests/golden/images.npz.  Build container only."""
import ctypes as C
import os
import struct
import sys
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np

IMG = os.path.join(HERE, "images")


def _chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)


def write_png(path, w, h, ctype, depth, rows, palette=None, level=6, idat_split=0):
    """rows: list of bytes objects (packed samples of one scanline each); every row gets filter (y % 5)."""
    bpp = max(1, {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype] * depth // 8)
    raw = bytearray()
    prev = bytes(len(rows[0]))
    for y, row in enumerate(rows):
        ft = y % 5
        out = bytearray(len(row))
        for i in range(len(row)):
            a = row[i - bpp] if i >= bpp else 0
            b = prev[i]
            c = prev[i - bpp] if i >= bpp else 0
            pred = (0, a, b, (a + b) >> 1, _paeth(a, b, c))[ft]
            out[i] = (row[i] - pred) & 0xFF
        raw += bytes([ft]) + out
        prev = row
    z = zlib.compress(bytes(raw), level)
    data = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0))
    if palette is not None:
        data += _chunk(b"PLTE", bytes(palette))
    data += _chunk(b"tEXt", b"Comment\0fixture")
    if idat_split:
        data += _chunk(b"IDAT", z[:idat_split]) + _chunk(b"IDAT", z[idat_split:])
    else:
        data += _chunk(b"IDAT", z)
    data += _chunk(b"IEND", b"")
    open(path, "wb").write(data)


def write_png_interlaced(path, w, h, ctype, depth, pixels, palette=None):
    """Adam7: `pixels[y][x]` = tuple of samples of one pixel; every pass is filtered on its own (filter = row % 5)."""
    nch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    bpp = max(1, nch * depth // 8)
    raw = bytearray()
    for x0, y0, dx, dy in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
        xs, ys = range(x0, w, dx), range(y0, h, dy)
        if not len(xs) or not len(ys):
            continue
        prev = None
        for r, y in enumerate(ys):
            if depth >= 8:
                row = b"".join(int(v).to_bytes(depth // 8, "big") for x in xs for v in pixels[y][x])
            else:
                bits = "".join(format(int(pixels[y][x][0]), "0%db" % depth) for x in xs)
                bits += "0" * (-len(bits) % 8)
                row = bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8))
            if prev is None:
                prev = bytes(len(row))
            ft = r % 5
            out = bytearray(len(row))
            for i in range(len(row)):
                a = row[i - bpp] if i >= bpp else 0
                b = prev[i]
                c = prev[i - bpp] if i >= bpp else 0
                out[i] = (row[i] - (0, a, b, (a + b) >> 1, _paeth(a, b, c))[ft]) & 0xFF
            raw += bytes([ft]) + out
            prev = row
    data = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 1))
    if palette is not None:
        data += _chunk(b"PLTE", bytes(palette))
    data += _chunk(b"IDAT", zlib.compress(bytes(raw), 6)) + _chunk(b"IEND", b"")
    open(path, "wb").write(data)


def write_hdr(path, img, rle):
    """img: (h, w, 4) uint8 RGBE."""
    h, w, _ = img.shape
    out = bytearray(b"#?RADIANCE\n# fixture\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=1.0\n\n-Y %d +X %d\n" % (h, w))
    for y in range(h):
        if not rle:
            out += img[y].tobytes()
            continue
        out += bytes([2, 2, w >> 8, w & 255])
        for k in range(4):
            ch = img[y, :, k].tolist()
            i = 0
            while i < w:
                run = 1
                while i + run < w and run < 127 and ch[i + run] == ch[i]:
                    run += 1
                if run >= 3:
                    out += bytes([128 + run, ch[i]])
                    i += run
                else:
                    j = i
                    while j < w and j - i < 128 and not (j + 2 < w and ch[j] == ch[j + 1] == ch[j + 2]):
                        j += 1
                    j = max(j, i + 1)
                    out += bytes([j - i]) + bytes(ch[i:j])
                    i = j
    open(path, "wb").write(out)


def make_files():
    os.makedirs(IMG, exist_ok=True)
    rs = np.random.RandomState(12345)
    smooth = lambda h, w, c: (np.clip(np.add.outer(np.arange(h) * 9, np.arange(w) * 5)[..., None] + rs.randint(0, 40, (h, w, c)), 0, 255)).astype(np.uint8)
    a = smooth(13, 17, 3); write_png(os.path.join(IMG, "rgb8.png"), 17, 13, 2, 8, [a[y].tobytes() for y in range(13)])
    a = smooth(9, 11, 4); write_png(os.path.join(IMG, "rgba8.png"), 11, 9, 6, 8, [a[y].tobytes() for y in range(9)], level=9, idat_split=40)
    a = smooth(7, 10, 1); write_png(os.path.join(IMG, "gray8.png"), 10, 7, 0, 8, [a[y].tobytes() for y in range(7)], level=1)
    a = smooth(6, 5, 2); write_png(os.path.join(IMG, "graya8.png"), 5, 6, 4, 8, [a[y].tobytes() for y in range(6)])
    a = rs.randint(0, 65536, (8, 9, 3)).astype(">u2"); write_png(os.path.join(IMG, "rgb16.png"), 9, 8, 2, 16, [a[y].tobytes() for y in range(8)])
    a = rs.randint(0, 65536, (5, 6, 1)).astype(">u2"); write_png(os.path.join(IMG, "gray16.png"), 6, 5, 0, 16, [a[y].tobytes() for y in range(5)], level=0)
    pal = rs.randint(0, 256, 3 * 29).tolist()
    a = rs.randint(0, 29, (10, 12, 1)).astype(np.uint8); write_png(os.path.join(IMG, "pal8.png"), 12, 10, 3, 8, [a[y].tobytes() for y in range(10)], palette=pal)
    for depth in (1, 2, 4):
        w, h = 13, 6
        v = rs.randint(0, 1 << depth, (h, w))
        rows = []
        for y in range(h):
            bits = "".join(format(int(x), "0%db" % depth) for x in v[y])
            bits += "0" * (-len(bits) % 8)
            rows.append(bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8)))
        write_png(os.path.join(IMG, "gray%d.png" % depth), w, h, 0, depth, rows)
        write_png(os.path.join(IMG, "pal%d.png" % depth), w, h, 3, depth, rows, palette=pal[:3 * (1 << depth)])
    big = smooth(64, 96, 3); write_png(os.path.join(IMG, "rgb8_big.png"), 96, 64, 2, 8, [big[y].tobytes() for y in range(64)], level=9)
    e = np.zeros((12, 24, 4), np.uint8)
    e[..., :3] = smooth(12, 24, 3); e[..., 3] = rs.randint(120, 136, (12, 24)); e[3:6, 4:20, :] = (200, 180, 90, 131); e[0, 0] = (1, 2, 3, 0)
    write_hdr(os.path.join(IMG, "env_rle.hdr"), e, True)
    write_hdr(os.path.join(IMG, "env_flat.hdr"), e, False)
    write_hdr(os.path.join(IMG, "narrow_flat.hdr"), e[:, :5].copy(), False)
    # interlaced PNG (Adam7), including sizes that leave some passes empty
    ia = rs.randint(0, 256, (11, 13, 3))
    write_png_interlaced(os.path.join(IMG, "adam7_rgb8.png"), 13, 11, 2, 8, ia)
    write_png_interlaced(os.path.join(IMG, "adam7_rgba16.png"), 5, 3, 6, 16, rs.randint(0, 65536, (3, 5, 4)))
    write_png_interlaced(os.path.join(IMG, "adam7_gray2.png"), 9, 10, 0, 2, rs.randint(0, 4, (10, 9, 1)))
    write_png_interlaced(os.path.join(IMG, "adam7_pal4.png"), 1, 7, 3, 4, rs.randint(0, 16, (7, 1, 1)), palette=pal[:48])
    # JPEG: baseline / progressive, 4:4:4 / 4:2:2 / 4:2:0 / 4:1:1 / 4:4:0, optimised Huffman tables, restart markers, grey,
    # RGB-tagged components (written with Pillow and OpenCV, which are in the build image)
    from PIL import Image
    import cv2
    pic = np.clip(smooth(29, 37, 3).astype(int) + (((np.mgrid[0:29, 0:37][1] // 5) % 2) * 60)[..., None], 0, 255).astype(np.uint8)
    P = lambda name, arr, **kw: Image.fromarray(arr).save(os.path.join(IMG, name), "JPEG", **kw)
    P("base444_q90.jpg", pic, quality=90, subsampling=0)
    P("base422_q60_opt.jpg", pic, quality=60, subsampling=1, optimize=True)
    P("base420_q75_rst.jpg", pic, quality=75, subsampling=2, restart_marker_blocks=2)
    P("prog420_q80.jpg", pic, quality=80, subsampling=2, progressive=True)
    P("prog444_q35.jpg", pic[:13, :9].copy(), quality=35, subsampling=0, progressive=True)
    P("gray_q70.jpg", pic[..., 1].copy(), quality=70)
    P("rgb_tagged_q85.jpg", pic, quality=85, subsampling=0, keep_rgb=True)
    P("one_pixel.jpg", pic[:1, :1].copy(), quality=95)
    cv2.imwrite(os.path.join(IMG, "cv411_rst.jpg"), pic, [cv2.IMWRITE_JPEG_QUALITY, 70, cv2.IMWRITE_JPEG_SAMPLING_FACTOR,
                                                         cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411, cv2.IMWRITE_JPEG_RST_INTERVAL, 2])
    # BMP (24-bit, 32-bit with masks, 8- and 4-bit palette, 1-bit, 16-bit 565 bit fields) and TGA (true colour / grey /
    # colour-mapped, raw and run-length, both row orders)
    small = pic[:9, :14].copy()
    small[:, :6] = small[:1, :1]
    sim = Image.fromarray(small)
    sim.save(os.path.join(IMG, "rgb24.bmp"))
    Image.fromarray(np.dstack([small, small[..., :1]]), "RGBA").save(os.path.join(IMG, "rgba32.bmp"))
    sim.convert("P", palette=Image.ADAPTIVE, colors=200).save(os.path.join(IMG, "pal8.bmp"))
    sim.convert("P", palette=Image.ADAPTIVE, colors=16).save(os.path.join(IMG, "pal4.bmp"), bits=4)
    sim.convert("1").save(os.path.join(IMG, "mono1.bmp"))
    v565 = ((small[..., 0] >> 3).astype(np.uint16) << 11) | ((small[..., 1] >> 2).astype(np.uint16) << 5) | (small[..., 2] >> 3)
    rows = [v565[y].astype("<u2").tobytes() for y in range(9)][::-1]
    body = b"".join(r + b"\0" * (-len(r) % 4) for r in rows)
    hdr = struct.pack("<IiiHHIIiiII", 40, 14, 9, 1, 16, 3, len(body), 2835, 2835, 0, 0) + struct.pack("<III", 0xF800, 0x07E0, 0x001F)
    open(os.path.join(IMG, "rgb565.bmp"), "wb").write(b"BM" + struct.pack("<IHHI", 14 + len(hdr) + len(body), 0, 0, 14 + len(hdr)) + hdr + body)
    sim.save(os.path.join(IMG, "rgb24.tga"))
    sim.save(os.path.join(IMG, "rgb24_rle_topdown.tga"), compression="tga_rle", orientation=1)
    Image.fromarray(np.dstack([small, small[..., :1]]), "RGBA").save(os.path.join(IMG, "rgba32_rle.tga"), compression="tga_rle")
    sim.convert("L").save(os.path.join(IMG, "grey8.tga"))
    sim.convert("P", palette=Image.ADAPTIVE, colors=64).save(os.path.join(IMG, "mapped8_rle.tga"), compression="tga_rle")
    cv2.imwrite(os.path.join(IMG, "cv440_prog.jpg"), pic, [cv2.IMWRITE_JPEG_QUALITY, 50, cv2.IMWRITE_JPEG_SAMPLING_FACTOR,
                                                          cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440, cv2.IMWRITE_JPEG_PROGRESSIVE, 1])


def main():
    make_files()
    from oracle.oracle import REF_SO
    ref = C.CDLL(REF_SO)
    ref.ref_image_load.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    d = {}
    for name in sorted(os.listdir(IMG)):
        for flip in (0, 1):
            w, h = C.c_int(0), C.c_int(0)
            p = os.path.join(IMG, name).encode()
            assert ref.ref_image_load(p, flip, C.addressof(w), C.addressof(h), None, 0) == 0, name
            out = np.zeros((h.value, w.value, 3), np.float32)
            assert ref.ref_image_load(p, flip, C.addressof(w), C.addressof(h), out.ctypes.data, out.nbytes) == 0
            d["%s_flip%d" % (name, flip)] = out
    np.savez_compressed(os.path.join(HERE, "images.npz"), **d)
    print("image fixtures:", len(d) // 2, "files")


if __name__ == "__main__":
    main()
