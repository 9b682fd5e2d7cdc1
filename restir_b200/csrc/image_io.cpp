// image_io.cpp -- texture / environment-map files -> linear float RGB, and the PNG writer of the output step.
//
// The reference loads every texture with stbi_loadf(file, &w, &h, &n, 3) after stbi_ldr_to_hdr_gamma(1.f)
// (image.cpp:16-33, scene.cpp:97-98): 8-bit images become v / 255.0f, Radiance .hdr files are decoded from RGBE,
// rows are flipped vertically for material textures and kept for the environment map (scene.cpp:98, 124-126).
// stb_image is a third-party single-header dependency of the reference (external/include/stb_image.h, v2.x); this
// file restates the formats scenes use -- PNG (RFC 2083 + RFC 1950/1951 inflate), Radiance RGBE, and JPEG (image_jpeg.cpp) --
// and is pinned against stb itself through oracle/ref_harness.cpp (ref_image_load) on the committed fixtures.
// Writing: Image::savePNG (image.cpp:41-57) stores 8-bit RGB; here with per-row filter selection and a fixed-Huffman LZ77 deflate.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "scene_host.h"

namespace rs {

bool decodeJPEG(const std::vector<uint8_t>& file, int& W, int& H, std::vector<uint8_t>& rgb, std::string& err);   // image_jpeg.cpp

namespace {

bool readFile(const std::string& path, std::vector<uint8_t>& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    size_t got = out.empty() ? 0 : fread(out.data(), 1, out.size(), f);
    fclose(f);
    return got == out.size();
}

// ------------------------------------------------------------------------------------------------ inflate (RFC 1951)
struct BitReader {
    const uint8_t* p; size_t n, pos = 0; uint32_t acc = 0; int cnt = 0; bool bad = false;
    BitReader(const uint8_t* d, size_t len) : p(d), n(len) {}
    uint32_t bits(int k) {
        while (cnt < k) {
            if (pos >= n) { bad = true; return 0; }
            acc |= (uint32_t)p[pos++] << cnt; cnt += 8;
        }
        uint32_t v = acc & ((k == 32) ? 0xffffffffu : ((1u << k) - 1u));
        acc >>= k; cnt -= k;
        return v;
    }
    void alignByte() { acc = 0; cnt = 0; }
};

struct Huffman {           // canonical code, decoded bit by bit through first-code / first-symbol tables
    uint16_t count[16] = {}, symbol[320] = {};
    bool build(const uint8_t* lens, int n) {
        memset(count, 0, sizeof count);
        for (int i = 0; i < n; i++) count[lens[i]]++;
        count[0] = 0;
        uint16_t offs[16]; offs[1] = 0;
        for (int l = 1; l < 15; l++) offs[l + 1] = offs[l] + count[l];
        for (int i = 0; i < n; i++) if (lens[i]) symbol[offs[lens[i]]++] = (uint16_t)i;
        return true;
    }
    int decode(BitReader& br) const {
        int code = 0, first = 0, index = 0;
        for (int l = 1; l <= 15; l++) {
            code |= (int)br.bits(1);
            if (br.bad) return -1;
            int c = count[l];
            if (code - c < first) return symbol[index + (code - first)];
            index += c; first += c; first <<= 1; code <<= 1;
        }
        return -1;
    }
};

bool inflateZlib(const std::vector<uint8_t>& in, std::vector<uint8_t>& out, std::string& err) {
    if (in.size() < 6 || (in[0] & 0x0f) != 8 || ((in[0] << 8 | in[1]) % 31) != 0 || (in[1] & 0x20)) { err = "bad zlib header"; return false; }
    BitReader br(in.data() + 2, in.size() - 2);
    static const uint16_t lenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t lenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t distBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t distExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    static const uint8_t clOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    for (;;) {
        uint32_t final = br.bits(1), type = br.bits(2);
        if (br.bad) { err = "truncated deflate stream"; return false; }
        if (type == 0) {
            br.alignByte();
            if (br.pos + 4 > br.n) { err = "truncated stored block"; return false; }
            uint32_t len = br.p[br.pos] | br.p[br.pos + 1] << 8, nlen = br.p[br.pos + 2] | br.p[br.pos + 3] << 8;
            br.pos += 4;
            if ((len ^ 0xffffu) != nlen || br.pos + len > br.n) { err = "bad stored block"; return false; }
            out.insert(out.end(), br.p + br.pos, br.p + br.pos + len);
            br.pos += len;
        } else if (type == 1 || type == 2) {
            Huffman lit, dist;
            uint8_t lens[320];
            if (type == 1) {
                for (int i = 0; i < 288; i++) lens[i] = i < 144 ? 8 : (i < 256 ? 9 : (i < 280 ? 7 : 8));
                lit.build(lens, 288);
                for (int i = 0; i < 30; i++) lens[i] = 5;
                dist.build(lens, 30);
            } else {
                int hlit = br.bits(5) + 257, hdist = br.bits(5) + 1, hclen = br.bits(4) + 4;
                uint8_t cl[19] = {};
                for (int i = 0; i < hclen; i++) cl[clOrder[i]] = (uint8_t)br.bits(3);
                Huffman clh; clh.build(cl, 19);
                int i = 0;
                while (i < hlit + hdist) {
                    int sym = clh.decode(br);
                    if (sym < 0) { err = "bad code lengths"; return false; }
                    if (sym < 16) lens[i++] = (uint8_t)sym;
                    else {
                        int rep, val = 0;
                        if (sym == 16) { if (i == 0) { err = "bad repeat"; return false; } val = lens[i - 1]; rep = 3 + br.bits(2); }
                        else if (sym == 17) rep = 3 + br.bits(3);
                        else rep = 11 + br.bits(7);
                        if (i + rep > hlit + hdist) { err = "bad repeat"; return false; }
                        while (rep--) lens[i++] = (uint8_t)val;
                    }
                }
                uint8_t dl[32];
                memcpy(dl, lens + hlit, hdist);
                lit.build(lens, hlit);
                dist.build(dl, hdist);
            }
            for (;;) {
                int sym = lit.decode(br);
                if (sym < 0 || br.bad) { err = "bad literal/length code"; return false; }
                if (sym < 256) out.push_back((uint8_t)sym);
                else if (sym == 256) break;
                else {
                    sym -= 257;
                    if (sym >= 29) { err = "bad length symbol"; return false; }
                    int len = lenBase[sym] + (int)br.bits(lenExtra[sym]);
                    int ds = dist.decode(br);
                    if (ds < 0 || ds >= 30) { err = "bad distance code"; return false; }
                    size_t d = distBase[ds] + br.bits(distExtra[ds]);
                    if (d > out.size()) { err = "distance too far back"; return false; }
                    size_t from = out.size() - d;
                    for (int k = 0; k < len; k++) out.push_back(out[from + k]);
                }
            }
        } else { err = "bad deflate block type"; return false; }
        if (final) break;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------ PNG -> 8-bit RGB
uint32_t be32(const uint8_t* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

int paeth(int a, int b, int c) {
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

bool decodePNG(const std::vector<uint8_t>& file, int& W, int& H, std::vector<uint8_t>& rgb, std::string& err) {
    size_t pos = 8;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, palette;
    bool haveHdr = false;
    while (pos + 12 <= file.size()) {
        uint32_t len = be32(&file[pos]);
        const uint8_t* tag = &file[pos + 4];
        const uint8_t* data = &file[pos + 8];
        if (pos + 12 + (size_t)len > file.size()) { err = "truncated PNG chunk"; return false; }
        if (!memcmp(tag, "IHDR", 4)) {
            if (len != 13) { err = "bad IHDR"; return false; }
            W = (int)be32(data); H = (int)be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
            haveHdr = true;
        } else if (!memcmp(tag, "PLTE", 4)) palette.assign(data, data + len);
        else if (!memcmp(tag, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!memcmp(tag, "IEND", 4)) break;
        pos += 12 + (size_t)len;
    }
    if (!haveHdr || W <= 0 || H <= 0) { err = "PNG without a valid IHDR"; return false; }
    if ((size_t)W * H > idat.size() * 1100 + 1024) { err = "PNG header claims more pixels than its data can hold"; return false; }   // deflate: <= 1032x
    int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!channels || !(depth == 8 || depth == 16 || ((ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4))) || (ctype == 3 && depth == 16)) {
        err = "unsupported PNG colour type / bit depth"; return false;
    }
    if (interlace > 1) { err = "PNG: unknown interlace method"; return false; }
    std::vector<uint8_t> raw;
    if (!inflateZlib(idat, raw, err)) { err = "PNG: " + err; return false; }
    rgb.assign((size_t)W * H * 3, 0);
    const int bpp = (channels * depth + 7) / 8;
    const int stride = depth == 16 ? 2 : 1;                          // 16-bit samples: the high byte (stb: >> 8)
    // one pass = a sub-image of its own scanlines and filter bytes (RFC 2083 section 2.6): the whole image, or the
    // seven Adam7 passes {x0, y0, dx, dy}
    static const int adam7[7][4] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
    static const int whole[1][4] = {{0, 0, 1, 1}};
    const int (*passes)[4] = interlace ? adam7 : whole;
    size_t rawPos = 0;
    std::vector<uint8_t> img;
    for (int pass = 0; pass < (interlace ? 7 : 1); pass++) {
        const int x0 = passes[pass][0], y0 = passes[pass][1], dx = passes[pass][2], dy = passes[pass][3];
        const int pw = (W - x0 + dx - 1) / dx, ph = (H - y0 + dy - 1) / dy;
        if (pw <= 0 || ph <= 0) continue;
        const size_t rowBytes = ((size_t)pw * channels * depth + 7) / 8;
        if (raw.size() < rawPos + (rowBytes + 1) * (size_t)ph) { err = "PNG: not enough pixel data"; return false; }
        img.assign(rowBytes * ph, 0);
        for (int y = 0; y < ph; y++) {                               // RFC 2083 section 6: the five row filters
            const uint8_t* src = &raw[rawPos + (rowBytes + 1) * y];
            uint8_t* cur = &img[rowBytes * y];
            const uint8_t* up = y ? cur - rowBytes : nullptr;
            int ft = src[0];
            if (ft > 4) { err = "PNG: bad filter"; return false; }
            for (size_t i = 0; i < rowBytes; i++) {
                int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= (size_t)bpp) ? up[i - bpp] : 0;
                int pred = ft == 0 ? 0 : ft == 1 ? a : ft == 2 ? b : ft == 3 ? ((a + b) >> 1) : paeth(a, b, c);
                cur[i] = (uint8_t)(src[1 + i] + pred);
            }
        }
        rawPos += (rowBytes + 1) * (size_t)ph;
        for (int y = 0; y < ph; y++) {
            const uint8_t* row = &img[rowBytes * y];
            for (int x = 0; x < pw; x++) {
                uint8_t* o = &rgb[((size_t)(y0 + y * dy) * W + (x0 + x * dx)) * 3];
                if (depth < 8) {
                    int perByte = 8 / depth, shift = (perByte - 1 - x % perByte) * depth;
                    int v = (row[x / perByte] >> shift) & ((1 << depth) - 1);
                    if (ctype == 3) {
                        if ((size_t)v * 3 + 2 >= palette.size()) { err = "PNG: palette index out of range"; return false; }
                        o[0] = palette[v * 3]; o[1] = palette[v * 3 + 1]; o[2] = palette[v * 3 + 2];
                    } else {
                        int g = v * (depth == 1 ? 0xff : depth == 2 ? 0x55 : 0x11);
                        o[0] = o[1] = o[2] = (uint8_t)g;
                    }
                    continue;
                }
                const uint8_t* p = row + (size_t)x * channels * stride;
                switch (ctype) {
                case 0: case 4: o[0] = o[1] = o[2] = p[0]; break;
                case 2: case 6: o[0] = p[0]; o[1] = p[stride]; o[2] = p[2 * stride]; break;
                case 3:
                    if ((size_t)p[0] * 3 + 2 >= palette.size()) { err = "PNG: palette index out of range"; return false; }
                    o[0] = palette[p[0] * 3]; o[1] = palette[p[0] * 3 + 1]; o[2] = palette[p[0] * 3 + 2];
                    break;
                }
            }
        }
    }
    return true;
}

// ------------------------------------------------------------------------------------------------ Radiance RGBE
bool hdrLine(const std::vector<uint8_t>& f, size_t& pos, std::string& line) {
    line.clear();
    if (pos >= f.size()) return false;
    while (pos < f.size() && f[pos] != '\n') line.push_back((char)f[pos++]);
    if (pos < f.size()) pos++;
    return true;
}

void rgbeToFloat(const uint8_t* rgbe, f3& out) {
    if (rgbe[3] != 0) {
        float scale = (float)ldexp(1.0f, (int)rgbe[3] - (128 + 8));
        out = mk3(rgbe[0] * scale, rgbe[1] * scale, rgbe[2] * scale);
    } else out = mk3(0.f, 0.f, 0.f);
}

bool decodeHDR(const std::vector<uint8_t>& f, int& W, int& H, std::vector<f3>& px, std::string& err) {
    size_t pos = 0;
    std::string line;
    hdrLine(f, pos, line);
    if (line != "#?RADIANCE" && line != "#?RGBE") { err = "corrupt HDR header"; return false; }
    bool valid = false;
    for (;;) {
        if (!hdrLine(f, pos, line)) { err = "truncated HDR header"; return false; }
        if (line.empty()) break;
        if (line == "FORMAT=32-bit_rle_rgbe") valid = true;
    }
    if (!valid) { err = "unsupported HDR format"; return false; }
    hdrLine(f, pos, line);
    char* end = nullptr;
    if (line.compare(0, 3, "-Y ") != 0) { err = "unsupported HDR data layout"; return false; }
    H = (int)strtol(line.c_str() + 3, &end, 10);
    while (*end == ' ') end++;
    if (strncmp(end, "+X ", 3) != 0) { err = "unsupported HDR data layout"; return false; }
    W = (int)strtol(end + 3, nullptr, 10);
    if (W <= 0 || H <= 0) { err = "bad HDR size"; return false; }
    if ((size_t)W * H > f.size() * 128) { err = "HDR header claims more pixels than the file can hold"; return false; }   // RLE: <= 64 px per 2 bytes per channel
    px.resize((size_t)W * H);
    auto flat = [&](size_t firstPixel) {
        for (size_t i = firstPixel; i < (size_t)W * H; i++) {
            if (pos + 4 > f.size()) return false;
            rgbeToFloat(&f[pos], px[i]); pos += 4;
        }
        return true;
    };
    if (W < 8 || W >= 32768) {
        if (!flat(0)) { err = "truncated HDR data"; return false; }
        return true;
    }
    std::vector<uint8_t> scan((size_t)W * 4);
    for (int j = 0; j < H; j++) {
        if (pos + 4 > f.size()) { err = "truncated HDR data"; return false; }
        if (f[pos] != 2 || f[pos + 1] != 2 || (f[pos + 2] & 0x80)) {
            // not run-length encoded: the file holds flat RGBE pixels (only meaningful on the first row)
            if (j != 0) { err = "HDR: mixed flat / RLE scanlines"; return false; }
            if (!flat(0)) { err = "truncated HDR data"; return false; }
            return true;
        }
        int len = f[pos + 2] << 8 | f[pos + 3];
        pos += 4;
        if (len != W) { err = "HDR: invalid decoded scanline length"; return false; }
        for (int k = 0; k < 4; k++) {
            int i = 0;
            while (i < W) {
                if (pos >= f.size()) { err = "truncated HDR data"; return false; }
                int count = f[pos++];
                if (count > 128) {
                    count -= 128;
                    if (count > W - i || pos >= f.size()) { err = "HDR: bad RLE data"; return false; }
                    uint8_t v = f[pos++];
                    for (int z = 0; z < count; z++) scan[(size_t)(i++) * 4 + k] = v;
                } else {
                    if (count > W - i || pos + count > f.size()) { err = "HDR: bad RLE data"; return false; }
                    for (int z = 0; z < count; z++) scan[(size_t)(i++) * 4 + k] = f[pos++];
                }
            }
        }
        for (int i = 0; i < W; i++) rgbeToFloat(&scan[(size_t)i * 4], px[(size_t)j * W + i]);
    }
    return true;
}

// ------------------------------------------------------------------------------------------------ BMP / TGA -> 8-bit RGB
// The uncompressed Windows bitmap variants (palette 1/4/8 bit, 16/24/32 bit with the default or explicit channel
// masks; header sizes 12/40/56/108/124) and Truevision TGA (true colour 15/16/24/32 bit, grey 8/16 bit, colour-mapped,
// raw or run-length packets, either row order) as the reference's loader reads them (stb_image.h:5071-5460, 5469-5700):
// channel masks are widened to 8 bits by bit replication, 5-bit TGA channels by v * 255 / 31, alpha is dropped.
struct Bytes {
    const std::vector<uint8_t>& f; size_t pos = 0;
    int u8() { return pos < f.size() ? f[pos++] : 0; }
    int u16() { int a = u8(); return a | u8() << 8; }
    uint32_t u32() { uint32_t a = (uint32_t)u16(); return a | (uint32_t)u16() << 16; }
    void skip(long n) { if (n > 0) pos += (size_t)n; }
};
int highBit(uint32_t z) { int n = -1; while (z) { n++; z >>= 1; } return n; }
int bitCount(uint32_t a) { int n = 0; while (a) { n += a & 1; a >>= 1; } return n; }
int widenChannel(uint32_t v, int shift, int bits) {                 // stb_image.h:5096-5114
    static const unsigned mul[9] = {0, 0xff, 0x55, 0x49, 0x11, 0x21, 0x41, 0x81, 0x01};
    static const unsigned sh[9] = {0, 0, 0, 1, 0, 2, 4, 6, 0};
    if (shift < 0) v <<= -shift; else v >>= shift;
    if (bits < 0 || bits > 8) return 0;
    v >>= (8 - bits);
    return (int)(v * mul[bits]) >> sh[bits];
}

bool decodeBMP(const std::vector<uint8_t>& file, int& W, int& H, std::vector<uint8_t>& rgb, std::string& err) {
    Bytes s{file};
    s.skip(2); s.u32(); s.u16(); s.u16();
    const long offset = (long)s.u32();
    const int hsz = (int)s.u32();
    uint32_t mr = 0, mg = 0, mb = 0, ma = 0;
    if (hsz != 12 && hsz != 40 && hsz != 56 && hsz != 108 && hsz != 124) { err = "BMP: unknown header size"; return false; }
    long w, h;
    if (hsz == 12) { w = s.u16(); h = s.u16(); } else { w = (int32_t)s.u32(); h = (int32_t)s.u32(); }
    if (s.u16() != 1) { err = "BMP: bad plane count"; return false; }
    const int bpp = s.u16();
    if (hsz != 12) {
        const int compress = (int)s.u32();
        if (compress == 1 || compress == 2) { err = "BMP: run-length compressed files are not supported"; return false; }
        for (int i = 0; i < 5; i++) s.u32();
        if (hsz == 40 || hsz == 56) {
            if (hsz == 56) for (int i = 0; i < 4; i++) s.u32();
            if (bpp == 16 || bpp == 32) {
                if (compress == 0) {
                    if (bpp == 32) { mr = 0xffu << 16; mg = 0xffu << 8; mb = 0xffu; ma = 0xffu << 24; }
                    else { mr = 31u << 10; mg = 31u << 5; mb = 31u; }
                } else if (compress == 3) {
                    mr = s.u32(); mg = s.u32(); mb = s.u32();
                    if (mr == mg && mg == mb) { err = "BMP: bad channel masks"; return false; }
                } else { err = "BMP: unsupported compression"; return false; }
            }
        } else {
            mr = s.u32(); mg = s.u32(); mb = s.u32(); ma = s.u32();
            s.u32();
            for (int i = 0; i < 12; i++) s.u32();
            if (hsz == 124) for (int i = 0; i < 4; i++) s.u32();
        }
    }
    const bool bottomUp = h > 0;
    if (h < 0) h = -h;
    if (w <= 0 || h <= 0 || (size_t)w * (size_t)h > file.size() * 8 + 64) { err = "BMP: bad size"; return false; }
    W = (int)w; H = (int)h;
    rgb.assign((size_t)W * H * 3, 0);
    size_t z = 0;
    if (bpp < 16) {
        int psize = hsz == 12 ? (int)((offset - 14 - 24) / 3) : (int)((offset - 14 - hsz) >> 2);
        if (psize <= 0 || psize > 256) { err = "BMP: bad palette"; return false; }
        uint8_t pal[256][3] = {};
        for (int i = 0; i < psize; i++) {
            pal[i][2] = (uint8_t)s.u8(); pal[i][1] = (uint8_t)s.u8(); pal[i][0] = (uint8_t)s.u8();
            if (hsz != 12) s.u8();
        }
        s.skip(offset - 14 - hsz - (long)psize * (hsz == 12 ? 3 : 4));
        int width;
        if (bpp == 1) width = (W + 7) >> 3; else if (bpp == 4) width = (W + 1) >> 1; else if (bpp == 8) width = W;
        else { err = "BMP: bad bit depth"; return false; }
        const int pad = (-width) & 3;
        for (int j = 0; j < H; j++) {
            int v = 0;
            for (int i = 0; i < W; i++) {
                int c;
                if (bpp == 1) { if ((i & 7) == 0) v = s.u8(); c = (v >> (7 - (i & 7))) & 1; }
                else if (bpp == 4) { if ((i & 1) == 0) v = s.u8(); c = (i & 1) ? (v & 15) : (v >> 4); }
                else c = s.u8();
                rgb[z++] = pal[c][0]; rgb[z++] = pal[c][1]; rgb[z++] = pal[c][2];
            }
            s.skip(pad);
        }
    } else {
        s.skip(offset - 14 - hsz);
        const int width = bpp == 24 ? 3 * W : (bpp == 16 ? 2 * W : 0), pad = (-width) & 3;
        if (bpp != 16 && bpp != 24 && bpp != 32) { err = "BMP: bad bit depth"; return false; }
        const bool easy = bpp == 24 || (bpp == 32 && mb == 0xff && mg == 0xff00 && mr == 0x00ff0000 && ma == 0xff000000);
        int rs = 0, gs = 0, bs = 0, rc = 0, gc = 0, bc = 0;
        if (!easy) {
            if (!mr || !mg || !mb) { err = "BMP: bad channel masks"; return false; }
            rs = highBit(mr) - 7; rc = bitCount(mr); gs = highBit(mg) - 7; gc = bitCount(mg); bs = highBit(mb) - 7; bc = bitCount(mb);
        }
        for (int j = 0; j < H; j++) {
            for (int i = 0; i < W; i++) {
                if (easy) {
                    rgb[z + 2] = (uint8_t)s.u8(); rgb[z + 1] = (uint8_t)s.u8(); rgb[z] = (uint8_t)s.u8();
                    if (bpp == 32) s.u8();
                } else {
                    uint32_t v = bpp == 16 ? (uint32_t)s.u16() : s.u32();
                    rgb[z] = (uint8_t)widenChannel(v & mr, rs, rc); rgb[z + 1] = (uint8_t)widenChannel(v & mg, gs, gc); rgb[z + 2] = (uint8_t)widenChannel(v & mb, bs, bc);
                }
                z += 3;
            }
            s.skip(pad);
        }
    }
    if (bottomUp)
        for (int j = 0; j < H / 2; j++)
            for (int i = 0; i < W * 3; i++) std::swap(rgb[(size_t)j * W * 3 + i], rgb[(size_t)(H - 1 - j) * W * 3 + i]);
    return true;
}

bool looksLikeTGA(const std::vector<uint8_t>& f) {                  // the format has no signature: stb_image.h:5469-5499
    if (f.size() < 18) return false;
    int mapType = f[1], type = f[2];
    if (mapType > 1) return false;
    if (mapType == 1) {
        if (type != 1 && type != 9) return false;
        int pb = f[7];
        if (pb != 8 && pb != 15 && pb != 16 && pb != 24 && pb != 32) return false;
    } else if (type != 2 && type != 3 && type != 10 && type != 11) return false;
    if ((f[12] | f[13] << 8) < 1 || (f[14] | f[15] << 8) < 1) return false;
    int bpp = f[16];
    if (mapType == 1 && bpp != 8 && bpp != 16) return false;
    return bpp == 8 || bpp == 15 || bpp == 16 || bpp == 24 || bpp == 32;
}

bool decodeTGA(const std::vector<uint8_t>& file, int& W, int& H, std::vector<uint8_t>& rgb, std::string& err) {
    Bytes s{file};
    const int idLen = s.u8(), indexed = s.u8();
    int type = s.u8();
    const int palStart = s.u16(), palLen = s.u16(), palBits = s.u8();
    s.u16(); s.u16();
    W = s.u16(); H = s.u16();
    const int bpp = s.u8(), desc = s.u8();
    const bool rle = type >= 8;
    if (rle) type -= 8;
    const bool bottomUp = ((desc >> 5) & 1) == 0;
    auto compOf = [&](int bits, bool grey, bool& rgb16) {
        rgb16 = false;
        switch (bits) {
        case 8: return 1;
        case 16: if (grey) return 2;     // fall through
        case 15: rgb16 = true; return 3;
        case 24: return 3;
        case 32: return 4;
        }
        return 0;
    };
    bool rgb16 = false;
    const int comp = indexed ? compOf(palBits, false, rgb16) : compOf(bpp, type == 3, rgb16);
    if (!comp || W < 1 || H < 1) { err = "TGA: unsupported pixel format"; return false; }
    if ((size_t)W * H > file.size() * 128 + 64) { err = "TGA: header claims more pixels than the file can hold"; return false; }
    std::vector<uint8_t> data((size_t)W * H * comp);
    auto read16 = [&](uint8_t* out) {
        int px = s.u16();
        out[0] = (uint8_t)((((px >> 10) & 31) * 255) / 31); out[1] = (uint8_t)((((px >> 5) & 31) * 255) / 31); out[2] = (uint8_t)(((px & 31) * 255) / 31);
    };
    s.skip(idLen);
    std::vector<uint8_t> palette;
    if (indexed) {
        s.skip(palStart);
        palette.assign((size_t)palLen * comp, 0);
        for (int i = 0; i < palLen; i++) {
            if (rgb16) read16(&palette[(size_t)i * comp]);
            else for (int j = 0; j < comp; j++) palette[(size_t)i * comp + j] = (uint8_t)s.u8();
        }
    }
    uint8_t px[4] = {0, 0, 0, 0};
    int count = 0;
    bool repeating = false;
    for (size_t i = 0; i < (size_t)W * H; i++) {
        bool readNext = true;
        if (rle) {
            if (count == 0) { int cmd = s.u8(); count = 1 + (cmd & 127); repeating = (cmd >> 7) != 0; }
            else if (repeating) readNext = false;
        }
        if (readNext) {
            if (indexed) {
                int idx = bpp == 8 ? s.u8() : s.u16();
                if (idx >= palLen) idx = 0;
                for (int j = 0; j < comp; j++) px[j] = palette.empty() ? 0 : palette[(size_t)idx * comp + j];
            } else if (rgb16) read16(px);
            else for (int j = 0; j < comp; j++) px[j] = (uint8_t)s.u8();
        }
        for (int j = 0; j < comp; j++) data[i * comp + j] = px[j];
        --count;
    }
    rgb.resize((size_t)W * H * 3);
    for (int y = 0; y < H; y++) {
        const int sy = bottomUp ? H - 1 - y : y;
        for (int x = 0; x < W; x++) {
            const uint8_t* p = &data[((size_t)sy * W + x) * comp];
            uint8_t* o = &rgb[((size_t)y * W + x) * 3];
            if (comp <= 2) o[0] = o[1] = o[2] = p[0];
            else if (rgb16) { o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; }
            else { o[0] = p[2]; o[1] = p[1]; o[2] = p[0]; }       // stored BGR(A)
        }
    }
    return true;
}

uint32_t crc32(const uint8_t* p, size_t n, uint32_t crc = 0) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1; table[i] = c; }
        init = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    return ~crc;
}

void putChunk(std::vector<uint8_t>& out, const char* tag, const std::vector<uint8_t>& data) {
    uint32_t n = (uint32_t)data.size();
    const uint8_t len[4] = {(uint8_t)(n >> 24), (uint8_t)(n >> 16), (uint8_t)(n >> 8), (uint8_t)n};
    out.insert(out.end(), len, len + 4);
    size_t start = out.size();
    out.insert(out.end(), tag, tag + 4);
    out.insert(out.end(), data.begin(), data.end());
    uint32_t c = crc32(&out[start], out.size() - start);
    const uint8_t cb[4] = {(uint8_t)(c >> 24), (uint8_t)(c >> 16), (uint8_t)(c >> 8), (uint8_t)c};
    out.insert(out.end(), cb, cb + 4);
}

}  // namespace

// Image::Image(filename) (image.cpp:16-33): stbi_loadf(..., 3) with ldr->hdr gamma 1; flipY = stbi_set_flip_vertically_on_load
bool loadImageRGB(const std::string& path, bool flipY, HostTexture& out, std::string& err) {
    std::vector<uint8_t> file;
    if (!readFile(path, file)) { err = "cannot read image file " + path; return false; }
    static const uint8_t pngSig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    int W = 0, H = 0;
    std::vector<f3> px;
    if (file.size() >= 8 && !memcmp(file.data(), pngSig, 8)) {
        std::vector<uint8_t> rgb;
        if (!decodePNG(file, W, H, rgb, err)) { err = path + ": " + err; return false; }
        px.resize((size_t)W * H);
        for (size_t i = 0; i < px.size(); i++) px[i] = mk3(rgb[3 * i] / 255.0f, rgb[3 * i + 1] / 255.0f, rgb[3 * i + 2] / 255.0f);   // stbi__ldr_to_hdr, gamma 1
    } else if (file.size() >= 4 && file[0] == 0xff && file[1] == 0xd8) {
        std::vector<uint8_t> rgb;
        if (!decodeJPEG(file, W, H, rgb, err)) { err = path + ": " + err; return false; }
        px.resize((size_t)W * H);
        for (size_t i = 0; i < px.size(); i++) px[i] = mk3(rgb[3 * i] / 255.0f, rgb[3 * i + 1] / 255.0f, rgb[3 * i + 2] / 255.0f);
    } else if (file.size() >= 11 && (!memcmp(file.data(), "#?RADIANCE", 10) || !memcmp(file.data(), "#?RGBE", 6))) {
        if (!decodeHDR(file, W, H, px, err)) { err = path + ": " + err; return false; }
    } else if ((file.size() >= 26 && file[0] == 'B' && file[1] == 'M') || looksLikeTGA(file)) {
        std::vector<uint8_t> rgb;
        const bool bmp = file[0] == 'B' && file[1] == 'M';
        if (!(bmp ? decodeBMP(file, W, H, rgb, err) : decodeTGA(file, W, H, rgb, err))) { err = path + ": " + err; return false; }
        px.resize((size_t)W * H);
        for (size_t i = 0; i < px.size(); i++) px[i] = mk3(rgb[3 * i] / 255.0f, rgb[3 * i + 1] / 255.0f, rgb[3 * i + 2] / 255.0f);
    } else {
        err = path + ": unsupported image format (PNG, JPEG, BMP, TGA and Radiance .hdr are supported)";
        return false;
    }
    if (flipY)
        for (int y = 0; y < H / 2; y++)
            for (int x = 0; x < W; x++) std::swap(px[(size_t)y * W + x], px[(size_t)(H - 1 - y) * W + x]);
    out.w = W; out.h = H; out.rgb.swap(px);
    return true;
}

// ---- deflate (RFC 1951) with the fixed Huffman code: greedy LZ77 over a 32 KiB window, 3-byte hash chains
namespace {
struct BitWriter {
    std::vector<uint8_t>& out; uint32_t acc = 0; int cnt = 0;
    void put(uint32_t v, int n) { acc |= v << cnt; cnt += n; while (cnt >= 8) { out.push_back((uint8_t)acc); acc >>= 8; cnt -= 8; } }
    void putHuff(uint32_t code, int n) { uint32_t r = 0; for (int i = 0; i < n; i++) r |= ((code >> i) & 1u) << (n - 1 - i); put(r, n); }   // codes go MSB first
    void flush() { if (cnt) { out.push_back((uint8_t)acc); acc = 0; cnt = 0; } }
};
void fixedLiteral(BitWriter& w, int sym) {            // RFC 1951 section 3.2.6
    if (sym < 144) w.putHuff(0x30 + sym, 8);
    else if (sym < 256) w.putHuff(0x190 + sym - 144, 9);
    else if (sym < 280) w.putHuff(sym - 256, 7);
    else w.putHuff(0xc0 + sym - 280, 8);
}
void deflateFixed(const std::vector<uint8_t>& in, std::vector<uint8_t>& out) {
    static const uint16_t lenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t lenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t distBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t distExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    BitWriter w{out};
    w.put(1, 1); w.put(1, 2);                           // one final block, fixed codes
    const size_t n = in.size();
    const int HASH = 1 << 15;
    std::vector<int> head(HASH, -1), prev(n ? n : 1, -1);
    auto hashAt = [&](size_t i) { return (int)(((in[i] << 10) ^ (in[i + 1] << 5) ^ in[i + 2]) & (HASH - 1)); };
    size_t i = 0;
    while (i < n) {
        int bestLen = 0, bestDist = 0;
        if (i + 2 < n) {
            int h = hashAt(i), cand = head[h], chain = 0;
            while (cand >= 0 && i - (size_t)cand <= 32768 && chain++ < 48) {
                int l = 0;
                const int maxLen = (int)std::min<size_t>(258, n - i);
                while (l < maxLen && in[cand + l] == in[i + l]) l++;
                if (l > bestLen) { bestLen = l; bestDist = (int)(i - cand); if (l == maxLen) break; }
                cand = prev[cand];
            }
        }
        auto insert = [&](size_t k) { if (k + 2 < n) { int h = hashAt(k); prev[k] = head[h]; head[h] = (int)k; } };
        if (bestLen >= 3) {
            int lc = 28; while (lenBase[lc] > bestLen) lc--;
            fixedLiteral(w, 257 + lc); w.put(bestLen - lenBase[lc], lenExtra[lc]);
            int dc = 29; while (distBase[dc] > bestDist) dc--;
            w.putHuff(dc, 5); w.put(bestDist - distBase[dc], distExtra[dc]);
            for (int k = 0; k < bestLen; k++) insert(i + k);
            i += bestLen;
        } else {
            fixedLiteral(w, in[i]);
            insert(i);
            i++;
        }
    }
    fixedLiteral(w, 256);
    w.flush();
}
}  // namespace

// Image::savePNG (image.cpp:41-57): 8-bit RGB rows, top row first.  Every row takes the PNG filter with the smallest sum of
// absolute residuals (the heuristic stb_image_write uses too), then zlib / deflate with the fixed Huffman code.
bool writePNG(const std::string& path, int W, int H, const uint8_t* rgb, std::string& err) {
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<uint8_t> ihdr = {(uint8_t)(W >> 24), (uint8_t)(W >> 16), (uint8_t)(W >> 8), (uint8_t)W, (uint8_t)(H >> 24), (uint8_t)(H >> 16), (uint8_t)(H >> 8), (uint8_t)H, 8, 2, 0, 0, 0};
    putChunk(out, "IHDR", ihdr);
    const size_t rowBytes = (size_t)W * 3;
    std::vector<uint8_t> raw;
    raw.reserve((rowBytes + 1) * H);
    std::vector<uint8_t> cand(rowBytes), best(rowBytes);
    for (int y = 0; y < H; y++) {
        const uint8_t* cur = rgb + rowBytes * y;
        const uint8_t* up = y ? cur - rowBytes : nullptr;
        long bestSum = -1; int bestFilter = 0;
        for (int ft = 0; ft < 5; ft++) {
            long sum = 0;
            for (size_t i = 0; i < rowBytes; i++) {
                int a = i >= 3 ? cur[i - 3] : 0, b = up ? up[i] : 0, c = (up && i >= 3) ? up[i - 3] : 0;
                int pred = ft == 0 ? 0 : ft == 1 ? a : ft == 2 ? b : ft == 3 ? ((a + b) >> 1) : paeth(a, b, c);
                cand[i] = (uint8_t)(cur[i] - pred);
                sum += abs((int)(int8_t)cand[i]);
            }
            if (bestSum < 0 || sum < bestSum) { bestSum = sum; bestFilter = ft; best.swap(cand); }
        }
        raw.push_back((uint8_t)bestFilter);
        raw.insert(raw.end(), best.begin(), best.end());
    }
    std::vector<uint8_t> z = {0x78, 0x5e};
    deflateFixed(raw, z);
    uint32_t a = 1, b = 0;
    for (uint8_t v : raw) { a = (a + v) % 65521u; b = (b + a) % 65521u; }
    uint32_t ad = b << 16 | a;
    z.push_back((uint8_t)(ad >> 24)); z.push_back((uint8_t)(ad >> 16)); z.push_back((uint8_t)(ad >> 8)); z.push_back((uint8_t)ad);
    putChunk(out, "IDAT", z);
    putChunk(out, "IEND", {});
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { err = "cannot write " + path; return false; }
    bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    if (!ok) err = "short write to " + path;
    return ok;
}

}  // namespace rs
