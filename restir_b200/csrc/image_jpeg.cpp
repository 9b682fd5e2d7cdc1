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
