// image_jpeg.cpp -- JPEG (JFIF / Adobe, Huffman, baseline + extended sequential + progressive, 8-bit, 1 or 3
// components) -> 8-bit RGB, for texture files.
//
// The reference reads textures through stb_image (external/include/stb_image.h, stbi_loadf(..., 3)); JPEG decoders
// differ in their last bit (IDCT rounding, chroma upsampling filter, colour matrix), and the decoded texels feed the
// bit-exact hot path, so this decoder reproduces stb_image's arithmetic choices exactly:
//   * ITU T.81 entropy decoding; coefficients are kept as 16-bit values (products wrap like stb's `short`);
//   * the inverse DCT is the 12-bit fixed-point "islow" factorisation with stb's rounding (column pass keeps 2 extra
//     bits: +512 >> 10; row pass +65536 + (128 << 17) >> 17, clamped)            stb_image.h:2225-2328
//   * chroma upsampling: 3:1 triangle filters for 2x horizontal / vertical / both, nearest for every other factor,
//     with stb's row stepping (a sample row is blended with the NEXT one in the upper half of its span and with the
//     previous one in the lower half)                                             stb_image.h:3236-3430, 3660-3700
//   * YCbCr -> RGB in 20-bit fixed point with the constants rounded to 12 bits first                stb_image.h:3431-3455
//   * components tagged 'R','G','B' (or Adobe transform 0 without a JFIF header) are taken as RGB      stb_image.h:3648
// Pinned against stb itself (oracle/ref_harness.cpp:ref_image_load) on the fixtures under tests/golden/images/.
// CMYK / YCCK (4 components), 12-bit and arithmetic-coded files are rejected.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "scene_host.h"

namespace rs {

namespace {

const uint8_t kZigzag[64 + 15] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,
                                  6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,
                                  39, 46, 53, 60, 61, 54, 47, 55, 62, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};

struct HuffTable {
    bool present = false;
    uint8_t symbols[256];
    int minCode[18], maxCode[18], firstIndex[18];     // per code length 1..16 (T.81 F.2.2.3)
    bool build(const int counts[16], const uint8_t* vals, int n) {
        memcpy(symbols, vals, n);
        int code = 0, k = 0;
        for (int len = 1; len <= 16; len++) {
            firstIndex[len] = k; minCode[len] = code;
            code += counts[len - 1]; k += counts[len - 1];
            maxCode[len] = counts[len - 1] ? code - 1 : -1;
            if (code > (1 << len)) return false;
            code <<= 1;
        }
        present = true;
        return true;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, hd = 0, ha = 0, dcPred = 0;
    int x = 0, y = 0, w2 = 0, h2 = 0, blocksW = 0;
    std::vector<uint8_t> data;      // w2 x h2 samples
    std::vector<int16_t> coeff;     // progressive: 64 per block
};

struct Decoder {
    const uint8_t* p; size_t n, pos = 0;
    std::string err;
    HuffTable dc[4], ac[4];
    uint16_t dequant[4][64] = {};
    Component comp[4];
    int ncomp = 0, W = 0, H = 0, hMax = 1, vMax = 1, mcuX = 0, mcuY = 0;
    bool progressive = false, jfif = false;
    int adobeTransform = -1, rgbTags = 0;
    int restartInterval = 0;
    // scan state
    int scanN = 0, order[4] = {}, specStart = 0, specEnd = 63, succHigh = 0, succLow = 0, eobRun = 0, todo = 0;
    uint32_t bitBuf = 0; int bitCnt = 0; int marker = -1; bool noMore = false;

    int get8() { return pos < n ? p[pos++] : 0; }
    int get16() { int a = get8(); return a << 8 | get8(); }
    bool fail(const char* m) { if (err.empty()) err = m; return false; }

    // ---- bit reader over the entropy-coded segment: a 0xFF followed by a non-zero byte is a marker and ends the data
    void fill() {
        do {
            unsigned b = noMore ? 0 : (unsigned)get8();
            if (b == 0xff) {
                int c = get8();
                while (c == 0xff) c = get8();
                if (c != 0) { marker = c; noMore = true; return; }
            }
            bitBuf |= b << (24 - bitCnt);
            bitCnt += 8;
        } while (bitCnt <= 24);
    }
    int bits(int k) {
        if (k == 0) return 0;
        if (bitCnt < k) fill();
        int v = (int)(bitBuf >> (32 - k));
        bitBuf <<= k; bitCnt -= k;
        return v;
    }
    int bit() { return bits(1); }
    int extend(int k) {               // RECEIVE + EXTEND (T.81 F.2.2.1)
        if (k == 0) return 0;
        int v = bits(k);
        return v < (1 << (k - 1)) ? v - (1 << k) + 1 : v;
    }
    int decodeSymbol(const HuffTable& t) {
        if (bitCnt < 16) fill();
        int code = 0;
        for (int len = 1; len <= 16; len++) {
            code = code << 1 | (int)(bitBuf >> 31);
            bitBuf <<= 1; bitCnt--;
            if (t.maxCode[len] >= 0 && code <= t.maxCode[len] && code >= t.minCode[len]) return t.symbols[t.firstIndex[len] + code - t.minCode[len]];
        }
        return -1;
    }
    void resetScan() {
        bitBuf = 0; bitCnt = 0; noMore = false; marker = -1; eobRun = 0;
        for (Component& c : comp) c.dcPred = 0;
        todo = restartInterval ? restartInterval : 0x7fffffff;
    }

    // ---- blocks
    bool blockBaseline(int16_t* d, Component& c) {
        int t = decodeSymbol(dc[c.hd]);
        if (t < 0 || t > 15) return fail("JPEG: bad Huffman code");
        memset(d, 0, 64 * sizeof(int16_t));
        c.dcPred += extend(t);
        const uint16_t* q = dequant[c.tq];
        d[0] = (int16_t)(c.dcPred * q[0]);
        for (int k = 1; k < 64;) {
            int rs = decodeSymbol(ac[c.ha]);
            if (rs < 0) return fail("JPEG: bad Huffman code");
            int s = rs & 15, r = rs >> 4;
            if (s == 0) {
                if (rs != 0xf0) break;
                k += 16;
            } else {
                k += r;
                int zig = kZigzag[k++];
                d[zig] = (int16_t)(extend(s) * q[zig]);
            }
        }
        return true;
    }
    bool blockProgDC(int16_t* d, Component& c) {
        if (specEnd != 0) return fail("JPEG: DC scan with AC coefficients");
        if (succHigh == 0) {
            memset(d, 0, 64 * sizeof(int16_t));
            int t = decodeSymbol(dc[c.hd]);
            if (t < 0 || t > 15) return fail("JPEG: bad Huffman code");
            c.dcPred += extend(t);
            d[0] = (int16_t)(c.dcPred << succLow);
        } else if (bit()) d[0] = (int16_t)(d[0] + (int16_t)(1 << succLow));
        return true;
    }
    void refine(int16_t* p, int16_t bitv) {
        if (bit() && (*p & bitv) == 0) *p = (int16_t)(*p > 0 ? *p + bitv : *p - bitv);
    }
    bool blockProgAC(int16_t* d, Component& c) {
        if (specStart == 0) return fail("JPEG: AC scan starting at the DC coefficient");
        if (succHigh == 0) {                                  // first pass over this band
            if (eobRun) { --eobRun; return true; }
            int k = specStart;
            do {
                int rs = decodeSymbol(ac[c.ha]);
                if (rs < 0) return fail("JPEG: bad Huffman code");
                int s = rs & 15, r = rs >> 4;
                if (s == 0) {
                    if (r < 15) {
                        eobRun = (1 << r) + (r ? bits(r) : 0) - 1;
                        break;
                    }
                    k += 16;
                } else {
                    k += r;
                    int zig = kZigzag[k++];
                    d[zig] = (int16_t)(extend(s) << succLow);
                }
            } while (k <= specEnd);
        } else {                                              // refinement pass (T.81 G.1.2.3)
            const int16_t bitv = (int16_t)(1 << succLow);
            if (eobRun) {
                --eobRun;
                for (int k = specStart; k <= specEnd; ++k) {
                    int16_t* q = &d[kZigzag[k]];
                    if (*q != 0) refine(q, bitv);
                }
            } else {
                int k = specStart;
                do {
                    int rs = decodeSymbol(ac[c.ha]);
                    if (rs < 0) return fail("JPEG: bad Huffman code");
                    int s = rs & 15, r = rs >> 4;
                    if (s == 0) {
                        if (r < 15) {
                            eobRun = (1 << r) - 1 + (r ? bits(r) : 0);
                            r = 64;                           // finish the band: only refinements remain
                        }
                    } else {
                        if (s != 1) return fail("JPEG: bad refinement code");
                        s = bit() ? bitv : -bitv;
                    }
                    while (k <= specEnd) {
                        int16_t* q = &d[kZigzag[k++]];
                        if (*q != 0) refine(q, bitv);
                        else {
                            if (r == 0) { *q = (int16_t)s; break; }
                            --r;
                        }
                    }
                } while (k <= specEnd);
            }
        }
        return true;
    }

    // ---- inverse DCT, stb_image.h:2225-2328
    static void idct1d(int s0, int s1, int s2, int s3, int s4, int s5, int s6, int s7, int x[4], int t[4]) {
        auto f = [](double v) { return (int)(v * 4096 + 0.5); };
        int p2 = s2, p3 = s6;
        int p1 = (p2 + p3) * f(0.5411961f);
        int t2 = p1 + p3 * f(-1.847759065f), t3 = p1 + p2 * f(0.765366865f);
        int t0 = (s0 + s4) * 4096, t1 = (s0 - s4) * 4096;
        x[0] = t0 + t3; x[3] = t0 - t3; x[1] = t1 + t2; x[2] = t1 - t2;
        t0 = s7; t1 = s5; t2 = s3; t3 = s1;
        p3 = t0 + t2; int p4 = t1 + t3; p1 = t0 + t3; p2 = t1 + t2;
        int p5 = (p3 + p4) * f(1.175875602f);
        t0 = t0 * f(0.298631336f); t1 = t1 * f(2.053119869f); t2 = t2 * f(3.072711026f); t3 = t3 * f(1.501321110f);
        p1 = p5 + p1 * f(-0.899976223f); p2 = p5 + p2 * f(-2.562915447f);
        p3 = p3 * f(-1.961570560f); p4 = p4 * f(-0.390180644f);
        t[3] = t3 + p1 + p4; t[2] = t2 + p2 + p3; t[1] = t1 + p2 + p4; t[0] = t0 + p1 + p3;
    }
    static uint8_t clamp8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
    static void idct(uint8_t* out, int stride, const int16_t* d) {
        int val[64];
        for (int i = 0; i < 8; i++) {
            const int16_t* c = d + i;
            int* v = val + i;
            if (c[8] == 0 && c[16] == 0 && c[24] == 0 && c[32] == 0 && c[40] == 0 && c[48] == 0 && c[56] == 0) {
                int dcterm = c[0] * 4;
                for (int r = 0; r < 8; r++) v[8 * r] = dcterm;
            } else {
                int x[4], t[4];
                idct1d(c[0], c[8], c[16], c[24], c[32], c[40], c[48], c[56], x, t);
                for (int k = 0; k < 4; k++) x[k] += 512;
                v[0] = (x[0] + t[3]) >> 10; v[56] = (x[0] - t[3]) >> 10;
                v[8] = (x[1] + t[2]) >> 10; v[48] = (x[1] - t[2]) >> 10;
                v[16] = (x[2] + t[1]) >> 10; v[40] = (x[2] - t[1]) >> 10;
                v[24] = (x[3] + t[0]) >> 10; v[32] = (x[3] - t[0]) >> 10;
            }
        }
        for (int i = 0; i < 8; i++) {
            const int* v = val + 8 * i;
            uint8_t* o = out + (size_t)stride * i;
            int x[4], t[4];
            idct1d(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], x, t);
            for (int k = 0; k < 4; k++) x[k] += 65536 + (128 << 17);
            o[0] = clamp8((x[0] + t[3]) >> 17); o[7] = clamp8((x[0] - t[3]) >> 17);
            o[1] = clamp8((x[1] + t[2]) >> 17); o[6] = clamp8((x[1] - t[2]) >> 17);
            o[2] = clamp8((x[2] + t[1]) >> 17); o[5] = clamp8((x[2] - t[1]) >> 17);
            o[3] = clamp8((x[3] + t[0]) >> 17); o[4] = clamp8((x[3] - t[0]) >> 17);
        }
    }

    // ---- one scan
    bool restartCheck(bool& stop) {               // after every MCU: at the end of a restart interval the next marker must be RSTn
        stop = false;
        if (--todo <= 0) {
            if (bitCnt < 24) fill();
            if (!(marker >= 0xd0 && marker <= 0xd7)) { stop = true; return true; }
            resetScan();
        }
        return true;
    }
    bool decodeBlockAt(Component& c, int bx, int by) {
        if (!progressive) {
            int16_t d[64];
            if (!blockBaseline(d, c)) return false;
            idct(&c.data[(size_t)c.w2 * by * 8 + bx * 8], c.w2, d);
            return true;
        }
        int16_t* d = &c.coeff[64 * ((size_t)bx + (size_t)by * c.blocksW)];
        return specStart == 0 ? blockProgDC(d, c) : blockProgAC(d, c);
    }
    bool scan() {
        resetScan();
        bool stop;
        if (scanN == 1) {
            Component& c = comp[order[0]];
            int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
            for (int j = 0; j < h; j++)
                for (int i = 0; i < w; i++) {
                    if (!decodeBlockAt(c, i, j)) return false;
                    restartCheck(stop);
                    if (stop) return true;
                }
            return true;
        }
        if (progressive && specStart != 0) return fail("JPEG: interleaved AC scan");
        for (int j = 0; j < mcuY; j++)
            for (int i = 0; i < mcuX; i++) {
                for (int k = 0; k < scanN; k++) {
                    Component& c = comp[order[k]];
                    for (int y = 0; y < c.v; y++)
                        for (int x = 0; x < c.h; x++)
                            if (!decodeBlockAt(c, i * c.h + x, j * c.v + y)) return false;
                }
                restartCheck(stop);
                if (stop) return true;
            }
        return true;
    }

    // ---- markers
    int nextMarker() {
        if (marker >= 0) { int m = marker; marker = -1; return m; }
        int x = get8();
        if (x != 0xff) return -1;
        while (x == 0xff) x = get8();
        return x;
    }
    bool tables(int m) {
        if (m == -1) return fail("JPEG: expected a marker");
        if (m == 0xdd) {
            if (get16() != 4) return fail("JPEG: bad DRI length");
            restartInterval = get16();
            return true;
        }
        if (m == 0xdb) {
            int L = get16() - 2;
            while (L > 0) {
                int q = get8(), prec = q >> 4, t = q & 15;
                if (prec > 1 || t > 3) return fail("JPEG: bad DQT");
                for (int i = 0; i < 64; i++) dequant[t][kZigzag[i]] = (uint16_t)(prec ? get16() : get8());
                L -= prec ? 129 : 65;
            }
            return L == 0 || fail("JPEG: bad DQT length");
        }
        if (m == 0xc4) {
            int L = get16() - 2;
            while (L > 0) {
                int q = get8(), tc = q >> 4, th = q & 15, counts[16], total = 0;
                if (tc > 1 || th > 3) return fail("JPEG: bad DHT");
                for (int i = 0; i < 16; i++) { counts[i] = get8(); total += counts[i]; }
                if (total > 256) return fail("JPEG: bad DHT");
                uint8_t vals[256];
                for (int i = 0; i < total; i++) vals[i] = (uint8_t)get8();
                if (!(tc ? ac[th] : dc[th]).build(counts, vals, total)) return fail("JPEG: bad code lengths");
                L -= 17 + total;
            }
            return L == 0 || fail("JPEG: bad DHT length");
        }
        if ((m >= 0xe0 && m <= 0xef) || m == 0xfe) {
            int L = get16();
            if (L < 2) return fail("JPEG: bad segment length");
            L -= 2;
            if (m == 0xe0 && L >= 5) {
                static const char tag[5] = {'J', 'F', 'I', 'F', 0};
                bool ok = true;
                for (int i = 0; i < 5; i++) if (get8() != (uint8_t)tag[i]) ok = false;
                L -= 5;
                if (ok) jfif = true;
            } else if (m == 0xee && L >= 12) {
                static const char tag[6] = {'A', 'd', 'o', 'b', 'e', 0};
                bool ok = true;
                for (int i = 0; i < 6; i++) if (get8() != (uint8_t)tag[i]) ok = false;
                L -= 6;
                if (ok) { get8(); get16(); get16(); adobeTransform = get8(); L -= 6; }
            }
            pos += (size_t)L;
            return true;
        }
        return fail("JPEG: unsupported marker (arithmetic coding, lossless and hierarchical files are not supported)");
    }
    bool frameHeader() {
        int Lf = get16();
        if (Lf < 11) return fail("JPEG: bad SOF length");
        if (get8() != 8) return fail("JPEG: only 8-bit samples are supported");
        H = get16(); W = get16();
        if (H == 0 || W == 0) return fail("JPEG: empty image");
        ncomp = get8();
        if (ncomp == 4) return fail("JPEG: CMYK / YCCK files are not supported");
        if (ncomp != 1 && ncomp != 3) return fail("JPEG: bad component count");
        if (Lf != 8 + 3 * ncomp) return fail("JPEG: bad SOF length");
        rgbTags = 0;
        for (int i = 0; i < ncomp; i++) {
            Component& c = comp[i];
            c.id = get8();
            if (ncomp == 3 && c.id == "RGB"[i]) rgbTags++;
            int q = get8();
            c.h = q >> 4; c.v = q & 15; c.tq = get8();
            if (!c.h || c.h > 4 || !c.v || c.v > 4 || c.tq > 3) return fail("JPEG: bad sampling factors");
            hMax = c.h > hMax ? c.h : hMax; vMax = c.v > vMax ? c.v : vMax;
        }
        if ((size_t)W * H > n * 4096 + 65536) return fail("JPEG: header claims more pixels than the file can hold");
        mcuX = (W + hMax * 8 - 1) / (hMax * 8); mcuY = (H + vMax * 8 - 1) / (vMax * 8);
        for (int i = 0; i < ncomp; i++) {
            Component& c = comp[i];
            c.x = (W * c.h + hMax - 1) / hMax; c.y = (H * c.v + vMax - 1) / vMax;
            c.w2 = mcuX * c.h * 8; c.h2 = mcuY * c.v * 8; c.blocksW = c.w2 / 8;
            c.data.assign((size_t)c.w2 * c.h2, 0);
            if (progressive) c.coeff.assign((size_t)c.w2 * c.h2, 0);
        }
        return true;
    }
    bool scanHeader() {
        int Ls = get16();
        scanN = get8();
        if (scanN < 1 || scanN > ncomp || Ls != 6 + 2 * scanN) return fail("JPEG: bad SOS");
        for (int i = 0; i < scanN; i++) {
            int id = get8(), q = get8(), which = 0;
            while (which < ncomp && comp[which].id != id) which++;
            if (which == ncomp) return fail("JPEG: SOS names an unknown component");
            comp[which].hd = q >> 4; comp[which].ha = q & 15;
            if (comp[which].hd > 3 || comp[which].ha > 3) return fail("JPEG: bad Huffman table index");
            order[i] = which;
        }
        specStart = get8(); specEnd = get8();
        int aa = get8();
        succHigh = aa >> 4; succLow = aa & 15;
        if (progressive) {
            if (specStart > 63 || specEnd > 63 || specStart > specEnd || succHigh > 13 || succLow > 13) return fail("JPEG: bad SOS");
        } else {
            if (specStart != 0 || succHigh != 0 || succLow != 0) return fail("JPEG: bad SOS");
            specEnd = 63;
        }
        for (int i = 0; i < scanN; i++) {
            const Component& c = comp[order[i]];
            if ((specStart == 0 && (!progressive || succHigh == 0) && !dc[c.hd].present) || ((!progressive || specStart > 0) && !ac[c.ha].present))
                return fail("JPEG: scan uses a Huffman table that was not defined");
        }
        return true;
    }
    bool decode() {
        if (nextMarker() != 0xd8) return fail("JPEG: no SOI");
        int m = nextMarker();
        while (!(m == 0xc0 || m == 0xc1 || m == 0xc2)) {
            if (!tables(m)) return false;
            m = nextMarker();
            while (m == -1) {
                if (pos >= n) return fail("JPEG: no SOF");
                m = nextMarker();
            }
        }
        progressive = m == 0xc2;
        if (!frameHeader()) return false;
        m = nextMarker();
        while (m != 0xd9) {
            if (m == 0xda) {
                if (!scanHeader() || !scan()) return false;
                if (marker < 0) {                       // trailing zero bytes after the entropy-coded data
                    while (pos < n) {
                        if (get8() == 255) { marker = get8(); break; }
                    }
                }
            } else if (m == 0xdc) {
                int Ld = get16(), NL = get16();
                if (Ld != 4 || NL != H) return fail("JPEG: bad DNL");
            } else if (!tables(m)) return false;
            if (pos >= n && marker < 0) return fail("JPEG: truncated file");
            m = nextMarker();
        }
        if (progressive)
            for (int i = 0; i < ncomp; i++) {
                Component& c = comp[i];
                int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
                for (int j = 0; j < h; j++)
                    for (int b = 0; b < w; b++) {
                        int16_t* d = &c.coeff[64 * ((size_t)b + (size_t)j * c.blocksW)];
                        for (int k = 0; k < 64; k++) d[k] = (int16_t)(d[k] * dequant[c.tq][k]);
                        idct(&c.data[(size_t)c.w2 * j * 8 + b * 8], c.w2, d);
                    }
            }
        return true;
    }

    // ---- upsampling + colour, stb_image.h:3236-3455, 3636-3720
    static const uint8_t* resample(uint8_t* out, const uint8_t* nearRow, const uint8_t* farRow, int w, int hs, int vs) {
        if (hs == 1 && vs == 1) return nearRow;
        if (hs == 1 && vs == 2) {
            for (int i = 0; i < w; i++) out[i] = (uint8_t)((3 * nearRow[i] + farRow[i] + 2) >> 2);
            return out;
        }
        if (hs == 2 && vs == 1) {
            const uint8_t* in = nearRow;
            if (w == 1) { out[0] = out[1] = in[0]; return out; }
            out[0] = in[0];
            out[1] = (uint8_t)((in[0] * 3 + in[1] + 2) >> 2);
            int i;
            for (i = 1; i < w - 1; i++) {
                int m = 3 * in[i] + 2;
                out[i * 2] = (uint8_t)((m + in[i - 1]) >> 2);
                out[i * 2 + 1] = (uint8_t)((m + in[i + 1]) >> 2);
            }
            out[i * 2] = (uint8_t)((in[w - 2] * 3 + in[w - 1] + 2) >> 2);
            out[i * 2 + 1] = in[w - 1];
            return out;
        }
        if (hs == 2 && vs == 2) {
            if (w == 1) { out[0] = out[1] = (uint8_t)((3 * nearRow[0] + farRow[0] + 2) >> 2); return out; }
            int t1 = 3 * nearRow[0] + farRow[0];
            out[0] = (uint8_t)((t1 + 2) >> 2);
            for (int i = 1; i < w; i++) {
                int t0 = t1;
                t1 = 3 * nearRow[i] + farRow[i];
                out[i * 2 - 1] = (uint8_t)((3 * t0 + t1 + 8) >> 4);
                out[i * 2] = (uint8_t)((3 * t1 + t0 + 8) >> 4);
            }
            out[w * 2 - 1] = (uint8_t)((t1 + 2) >> 2);
            return out;
        }
        for (int i = 0; i < w; i++)
            for (int j = 0; j < hs; j++) out[i * hs + j] = nearRow[i];
        return out;
    }
    void toRGB(std::vector<uint8_t>& rgb) {
        rgb.resize((size_t)W * H * 3);
        const bool isRGB = ncomp == 3 && (rgbTags == 3 || (adobeTransform == 0 && !jfif));
        struct Up { int hs, vs, ystep, wLores, ypos; size_t line0, line1; std::vector<uint8_t> buf; } up[3];
        for (int k = 0; k < ncomp; k++) {
            up[k].hs = hMax / comp[k].h; up[k].vs = vMax / comp[k].v;
            up[k].ystep = up[k].vs >> 1; up[k].wLores = (W + up[k].hs - 1) / up[k].hs;
            up[k].ypos = 0; up[k].line0 = up[k].line1 = 0;
            up[k].buf.assign((size_t)W + 8, 0);
        }
        const uint8_t* rows[3] = {nullptr, nullptr, nullptr};
        for (int j = 0; j < H; j++) {
            for (int k = 0; k < ncomp; k++) {
                Up& r = up[k];
                const uint8_t* base = comp[k].data.data();
                bool bottom = r.ystep >= (r.vs >> 1);
                rows[k] = resample(r.buf.data(), base + (bottom ? r.line1 : r.line0), base + (bottom ? r.line0 : r.line1), r.wLores, r.hs, r.vs);
                if (++r.ystep >= r.vs) {
                    r.ystep = 0;
                    r.line0 = r.line1;
                    if (++r.ypos < comp[k].y) r.line1 += comp[k].w2;
                }
            }
            uint8_t* out = &rgb[(size_t)j * W * 3];
            if (ncomp == 1) {
                for (int i = 0; i < W; i++) out[3 * i] = out[3 * i + 1] = out[3 * i + 2] = rows[0][i];
            } else if (isRGB) {
                for (int i = 0; i < W; i++) { out[3 * i] = rows[0][i]; out[3 * i + 1] = rows[1][i]; out[3 * i + 2] = rows[2][i]; }
            } else {
                auto fx = [](float v) { return ((int)(v * 4096.0f + 0.5f)) << 8; };
                for (int i = 0; i < W; i++) {
                    int yf = (rows[0][i] << 20) + (1 << 19);
                    int cr = rows[2][i] - 128, cb = rows[1][i] - 128;
                    int r = yf + cr * fx(1.40200f);
                    int g = yf + (cr * -fx(0.71414f)) + ((cb * -fx(0.34414f)) & 0xffff0000);
                    int b = yf + cb * fx(1.77200f);
                    out[3 * i] = clamp8(r >> 20); out[3 * i + 1] = clamp8(g >> 20); out[3 * i + 2] = clamp8(b >> 20);
                }
            }
        }
    }
};

}  // namespace

bool decodeJPEG(const std::vector<uint8_t>& file, int& W, int& H, std::vector<uint8_t>& rgb, std::string& err) {
    Decoder d{file.data(), file.size()};
    if (!d.decode()) { err = d.err.empty() ? "JPEG: corrupt file" : d.err; return false; }
    d.toRGB(rgb);
    W = d.W; H = d.H;
    return true;
}

}  // namespace rs

// ------------------------------------------------------------------------------------------------ JPEG writer
// Image::saveJPG (image.cpp:59-75) = stbi_write_jpg(..., quality 90) of stb_image_write 1.10: baseline JFIF, 4:4:4, the
// Annex K Huffman and quantisation tables (scaled by the IJG quality rule), the Arai-Agui-Nakajima float DCT with the
// scale factors folded into the divisors, coefficients rounded half away from zero.  Restated so that the file is the
// one the reference writes, byte for byte (pinned through oracle/ref_harness.cpp:ref_image_save).
namespace rs {

namespace {

const uint8_t kDcLumaCounts[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const uint8_t kDcChromaCounts[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
const uint8_t kDcValues[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t kAcLumaCounts[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
const uint8_t kAcChromaCounts[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
const uint8_t kAcLumaValues[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1,
    0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37,
    0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a,
    0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
    0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3,
    0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
const uint8_t kAcChromaValues[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1,
    0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36,
    0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69,
    0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a,
    0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca,
    0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
const int kLumaQ[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
                        18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const int kChromaQ[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                          99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
// position of natural-order coefficient i in the zigzag sequence
const uint8_t kToZigzag[64] = {0,  1,  5,  6,  14, 15, 27, 28, 2,  4,  7,  13, 16, 26, 29, 42, 3,  8,  12, 17, 25, 30, 41, 43, 9,  11, 18, 24, 31, 40, 44, 53,
                               10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};

struct Code { uint16_t bits, len; };
void canonicalCodes(const uint8_t counts[16], const uint8_t* values, Code table[256]) {    // T.81 Annex C
    memset(table, 0, 256 * sizeof(Code));
    int code = 0, k = 0;
    for (int len = 1; len <= 16; len++) {
        for (int i = 0; i < counts[len - 1]; i++) table[values[k++]] = Code{(uint16_t)code++, (uint16_t)len};
        code <<= 1;
    }
}

struct JpegOut {
    std::vector<uint8_t> out;
    int bitBuf = 0, bitCnt = 0;
    void put(const void* p, size_t n) { out.insert(out.end(), (const uint8_t*)p, (const uint8_t*)p + n); }
    void bits(Code c) {
        bitCnt += c.len;
        bitBuf |= c.bits << (24 - bitCnt);
        while (bitCnt >= 8) {
            uint8_t b = (uint8_t)((bitBuf >> 16) & 255);
            out.push_back(b);
            if (b == 255) out.push_back(0);
            bitBuf <<= 8; bitCnt -= 8;
        }
    }
};

void fdct8(float& d0, float& d1, float& d2, float& d3, float& d4, float& d5, float& d6, float& d7) {   // AAN forward DCT, one dimension
    float t0 = d0 + d7, t7 = d0 - d7, t1 = d1 + d6, t6 = d1 - d6, t2 = d2 + d5, t5 = d2 - d5, t3 = d3 + d4, t4 = d3 - d4;
    float t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    float e0 = t10 + t11, e4 = t10 - t11;
    float z1 = (t12 + t13) * 0.707106781f;
    float e2 = t13 + z1, e6 = t13 - z1;
    t10 = t4 + t5; t11 = t5 + t6; t12 = t6 + t7;
    float z5 = (t10 - t12) * 0.382683433f;
    float z2 = t10 * 0.541196100f + z5, z4 = t12 * 1.306562965f + z5, z3 = t11 * 0.707106781f;
    float z11 = t7 + z3, z13 = t7 - z3;
    d5 = z13 + z2; d3 = z13 - z2; d1 = z11 + z4; d7 = z11 - z4;
    d0 = e0; d2 = e2; d4 = e4; d6 = e6;
}

Code magnitude(int val) {
    int mag = val < 0 ? -val : val;
    val = val < 0 ? val - 1 : val;
    uint16_t len = 1;
    while (mag >>= 1) ++len;
    return Code{(uint16_t)(val & ((1 << len) - 1)), len};
}

int encodeBlock(JpegOut& o, float* blk, const float* divisors, int prevDC, const Code* dcTab, const Code* acTab) {
    for (int r = 0; r < 64; r += 8) fdct8(blk[r], blk[r + 1], blk[r + 2], blk[r + 3], blk[r + 4], blk[r + 5], blk[r + 6], blk[r + 7]);
    for (int c = 0; c < 8; c++) fdct8(blk[c], blk[c + 8], blk[c + 16], blk[c + 24], blk[c + 32], blk[c + 40], blk[c + 48], blk[c + 56]);
    int q[64];
    for (int i = 0; i < 64; i++) {
        float v = blk[i] * divisors[i];
        q[kToZigzag[i]] = (int)(v < 0 ? v - 0.5f : v + 0.5f);
    }
    int diff = q[0] - prevDC;
    if (diff == 0) o.bits(dcTab[0]);
    else { Code m = magnitude(diff); o.bits(dcTab[m.len]); o.bits(m); }
    int last = 63;
    while (last > 0 && q[last] == 0) --last;
    if (last == 0) { o.bits(acTab[0x00]); return q[0]; }
    for (int i = 1; i <= last; ++i) {
        int start = i;
        while (q[i] == 0 && i <= last) ++i;
        int zeros = i - start;
        for (int k = 0; k < (zeros >> 4); k++) o.bits(acTab[0xf0]);
        zeros &= 15;
        Code m = magnitude(q[i]);
        o.bits(acTab[(zeros << 4) + m.len]);
        o.bits(m);
    }
    if (last != 63) o.bits(acTab[0x00]);
    return q[0];
}

}  // namespace

bool writeJPG(const std::string& path, int W, int H, const unsigned char* rgb, int quality, std::string& err) {
    if (W <= 0 || H <= 0 || W > 65535 || H > 65535) { err = "JPEG: bad image size"; return false; }
    quality = quality ? quality : 90;
    quality = quality < 1 ? 1 : (quality > 100 ? 100 : quality);
    quality = quality < 50 ? 5000 / quality : 200 - quality * 2;
    uint8_t lumaT[64], chromaT[64];
    for (int i = 0; i < 64; i++) {
        int y = (kLumaQ[i] * quality + 50) / 100, c = (kChromaQ[i] * quality + 50) / 100;
        lumaT[kToZigzag[i]] = (uint8_t)(y < 1 ? 1 : (y > 255 ? 255 : y));
        chromaT[kToZigzag[i]] = (uint8_t)(c < 1 ? 1 : (c > 255 ? 255 : c));
    }
    static const float aan[8] = {1.0f * 2.828427125f, 1.387039845f * 2.828427125f, 1.306562965f * 2.828427125f, 1.175875602f * 2.828427125f,
                                 1.0f * 2.828427125f, 0.785694958f * 2.828427125f, 0.541196100f * 2.828427125f, 0.275899379f * 2.828427125f};
    float divY[64], divC[64];
    for (int r = 0, k = 0; r < 8; r++)
        for (int c = 0; c < 8; c++, k++) {
            divY[k] = 1 / (lumaT[kToZigzag[k]] * aan[r] * aan[c]);
            divC[k] = 1 / (chromaT[kToZigzag[k]] * aan[r] * aan[c]);
        }
    Code dcY[256], dcC[256], acY[256], acC[256];
    canonicalCodes(kDcLumaCounts, kDcValues, dcY); canonicalCodes(kDcChromaCounts, kDcValues, dcC);
    canonicalCodes(kAcLumaCounts, kAcLumaValues, acY); canonicalCodes(kAcChromaCounts, kAcChromaValues, acC);
    JpegOut o;
    static const uint8_t head0[] = {0xFF, 0xD8, 0xFF, 0xE0, 0, 0x10, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0, 0xFF, 0xDB, 0, 0x84, 0};
    o.put(head0, sizeof head0); o.put(lumaT, 64);
    o.out.push_back(1); o.put(chromaT, 64);
    const uint8_t sof[] = {0xFF, 0xC0, 0, 0x11, 8, (uint8_t)(H >> 8), (uint8_t)H, (uint8_t)(W >> 8), (uint8_t)W, 3, 1, 0x11, 0, 2, 0x11, 1, 3, 0x11, 1, 0xFF, 0xC4, 0x01, 0xA2, 0};
    o.put(sof, sizeof sof);
    o.put(kDcLumaCounts, 16); o.put(kDcValues, 12);
    o.out.push_back(0x10); o.put(kAcLumaCounts, 16); o.put(kAcLumaValues, 162);
    o.out.push_back(1); o.put(kDcChromaCounts, 16); o.put(kDcValues, 12);
    o.out.push_back(0x11); o.put(kAcChromaCounts, 16); o.put(kAcChromaValues, 162);
    static const uint8_t sos[] = {0xFF, 0xDA, 0, 0xC, 3, 1, 0, 2, 0x11, 3, 0x11, 0, 0x3F, 0};
    o.put(sos, sizeof sos);
    int dY = 0, dU = 0, dV = 0;
    for (int y = 0; y < H; y += 8)
        for (int x = 0; x < W; x += 8) {
            float Y[64], U[64], V[64];
            for (int row = y, pos = 0; row < y + 8; ++row) {
                const int rr = row < H ? row : H - 1;                 // edge blocks repeat the last row / column
                for (int col = x; col < x + 8; ++col, ++pos) {
                    const unsigned char* p = rgb + ((size_t)rr * W + (col < W ? col : W - 1)) * 3;
                    float r = p[0], g = p[1], b = p[2];
                    Y[pos] = +0.29900f * r + 0.58700f * g + 0.11400f * b - 128;
                    U[pos] = -0.16874f * r - 0.33126f * g + 0.50000f * b;
                    V[pos] = +0.50000f * r - 0.41869f * g - 0.08131f * b;
                }
            }
            dY = encodeBlock(o, Y, divY, dY, dcY, acY);
            dU = encodeBlock(o, U, divC, dU, dcC, acC);
            dV = encodeBlock(o, V, divC, dV, dcC, acC);
        }
    o.bits(Code{0x7F, 7});
    o.out.push_back(0xFF); o.out.push_back(0xD9);
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { err = "cannot write " + path; return false; }
    bool ok = fwrite(o.out.data(), 1, o.out.size(), f) == o.out.size();
    fclose(f);
    if (!ok) err = "short write to " + path;
    return ok;
}

}  // namespace rs
