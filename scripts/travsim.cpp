// travsim.cpp -- offline model of how a warp walks the traced tree (development aid, CPU only; not part of the product).
//
// Reads the traced BVH2 (RSTR_SCENE_TRACED_NODES / _TRIS dumps) and a camera, generates the primary rays of every
// 8x4-pixel warp tile (pixel-centre ray + a jittered ray, like k_gbuffer_restir_a) and replays different traversal
// schemes in lockstep to count WARP-level steps and lane utilisation:
//   per-lane while-while (the shipped scheme), per-lane walk of a collapsed 4- or 8-wide tree, and warp-packet walks.
// Build: g++ -O2 -fopenmp -o /tmp/travsim scripts/travsim.cpp ; run: /tmp/travsim dump_dir W H
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

struct Node { float lmin[3], lmax[3], rmin[3], rmax[3]; int left, right, pad[2]; };
struct Tri { float v[3][3]; int matId, prim, pad; };
struct Cam { float pos[3], right[3], up[3], view[3], aspect, tanFovY; };

static std::vector<Node> nodes;
static std::vector<Tri> tris;
static int root;

template <typename T>
static std::vector<T> readAll(const std::string& p) {
    FILE* f = fopen(p.c_str(), "rb");
    if (!f) { perror(p.c_str()); exit(1); }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<T> v(n / sizeof(T));
    if (fread(v.data(), sizeof(T), v.size(), f) != v.size()) exit(1);
    fclose(f);
    return v;
}

struct Ray { float o[3], d[3], inv[3], oi[3]; };
static Ray makeRay(const float* o, const float* d) {
    Ray r;
    for (int a = 0; a < 3; a++) {
        r.o[a] = o[a]; r.d[a] = d[a];
        float dd = fabsf(d[a]) < 1e-20f ? copysignf(1e-20f, d[a]) : d[a];
        r.inv[a] = 1.f / dd; r.oi[a] = -o[a] * r.inv[a];
    }
    return r;
}
static inline bool slab(const Ray& r, const float* lo, const float* hi, float limit, float& t) {
    float t0 = 0.f, t1 = limit;
    for (int a = 0; a < 3; a++) {
        float x0 = lo[a] * r.inv[a] + r.oi[a], x1 = hi[a] * r.inv[a] + r.oi[a];
        t0 = fmaxf(t0, fminf(x0, x1)); t1 = fminf(t1, fmaxf(x0, x1));
    }
    t = t0;
    return t0 <= t1;
}
static inline bool triHit(const Ray& r, const Tri& t, float& dist) {
    float e1[3], e2[3], p[3], s[3], q[3];
    for (int a = 0; a < 3; a++) { e1[a] = t.v[1][a] - t.v[0][a]; e2[a] = t.v[2][a] - t.v[0][a]; }
    p[0] = r.d[1] * e2[2] - r.d[2] * e2[1]; p[1] = r.d[2] * e2[0] - r.d[0] * e2[2]; p[2] = r.d[0] * e2[1] - r.d[1] * e2[0];
    float det = e1[0] * p[0] + e1[1] * p[1] + e1[2] * p[2];
    if (fabsf(det) < 1.19e-7f) return false;
    float inv = 1.f / det;
    for (int a = 0; a < 3; a++) s[a] = r.o[a] - t.v[0][a];
    float u = (s[0] * p[0] + s[1] * p[1] + s[2] * p[2]) * inv;
    if (u < 0 || u > 1) return false;
    q[0] = s[1] * e1[2] - s[2] * e1[1]; q[1] = s[2] * e1[0] - s[0] * e1[2]; q[2] = s[0] * e1[1] - s[1] * e1[0];
    float v = (r.d[0] * q[0] + r.d[1] * q[1] + r.d[2] * q[2]) * inv;
    if (v < 0 || u + v > 1) return false;
    dist = (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]) * inv;
    return dist > 0;
}
static inline int leafFirst(int ref) { return ref & 0x07ffffff; }
static inline int leafCount(int ref) { return ((ref >> 27) & 7) + 1; }

// ---------------------------------------------------------------- wide tree (collapse of the BVH2)
struct WNode { int n; float lo[8][3], hi[8][3]; int child[8]; };
static std::vector<WNode> wnodes;
static float area(const float* lo, const float* hi) { float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2]; return x * y + y * z + z * x; }
static int buildWide(int ref2, int width) {
    // ref2: internal node of the BVH2
    struct C { float lo[3], hi[3]; int ref; };
    std::vector<C> cs;
    auto push = [&](const float* lo, const float* hi, int ref) { C c; memcpy(c.lo, lo, 12); memcpy(c.hi, hi, 12); c.ref = ref; cs.push_back(c); };
    const Node& n = nodes[ref2];
    push(n.lmin, n.lmax, n.left); push(n.rmin, n.rmax, n.right);
    while ((int)cs.size() < width) {
        int best = -1; float ba = -1;
        for (int i = 0; i < (int)cs.size(); i++) if (cs[i].ref >= 0) { float a = area(cs[i].lo, cs[i].hi); if (a > ba) { ba = a; best = i; } }
        if (best < 0) break;
        const Node& m = nodes[cs[best].ref];
        cs.erase(cs.begin() + best);
        push(m.lmin, m.lmax, m.left); push(m.rmin, m.rmax, m.right);
    }
    int id = (int)wnodes.size();
    wnodes.push_back(WNode());
    WNode w; w.n = (int)cs.size();
    for (int i = 0; i < w.n; i++) { memcpy(w.lo[i], cs[i].lo, 12); memcpy(w.hi[i], cs[i].hi, 12); w.child[i] = cs[i].ref >= 0 ? buildWide(cs[i].ref, width) : cs[i].ref; }
    wnodes[id] = w;
    return id;
}

// ---------------------------------------------------------------- per-lane state machines
struct Lane2 {     // BVH2 while-while, one or two rays (pair walk) per lane
    Ray r[2]; int nr; float best[2];
    int stack[128]; float stackT[128]; int sp; int cur; bool done;
    long nodeSteps, triSteps;
};
static const int DONE = 0x7fffffff;
static thread_local bool g_anyhit = false;
static thread_local const float* g_seg = nullptr;    // any-hit: per-lane segment lengths
static thread_local int g_laneIdx = 0;
static void laneBegin(Lane2& L) {
    L.sp = 0; L.cur = root; L.done = false; L.nodeSteps = L.triSteps = 0;
    for (int i = 0; i < L.nr; i++) L.best[i] = 3e38f;
}
static inline float laneLimit(const Lane2& L) { float m = L.best[0]; for (int i = 1; i < L.nr; i++) m = fmaxf(m, L.best[i]); return m; }
static int lanePop(Lane2& L) {
    float lim = laneLimit(L);
    while (L.sp > 0) { --L.sp; if (L.stackT[L.sp] <= lim) return L.stack[L.sp]; }
    return DONE;
}
// one internal-node step; returns false if the lane is not at an internal node
static bool laneNodeStep(Lane2& L) {
    if (L.cur < 0 || L.cur == DONE) return false;
    const Node& n = nodes[L.cur];
    L.nodeSteps++;
    bool hL = false, hR = false; float tL = 3e38f, tR = 3e38f;
    for (int i = 0; i < L.nr; i++) {
        float t;
        if (slab(L.r[i], n.lmin, n.lmax, L.best[i], t)) { hL = true; tL = fminf(tL, t); }
        if (slab(L.r[i], n.rmin, n.rmax, L.best[i], t)) { hR = true; tR = fminf(tR, t); }
    }
    if (hL && hR) {
        bool ln = tL <= tR;
        L.stack[L.sp] = ln ? n.right : n.left; L.stackT[L.sp] = ln ? tR : tL; L.sp++;
        L.cur = ln ? n.left : n.right;
    } else if (hL) L.cur = n.left;
    else if (hR) L.cur = n.right;
    else L.cur = lanePop(L);
    return true;
}
// leaf: returns number of triangles (the lane processes them all), advances
static int laneLeaf(Lane2& L) {
    if (L.cur >= 0) return 0;     // DONE is positive
    int f = leafFirst(L.cur), c = leafCount(L.cur);
    for (int i = 0; i < c; i++)
        for (int k = 0; k < L.nr; k++) { float d; if (triHit(L.r[k], tris[f + i], d) && d < L.best[k]) { if (g_anyhit) { L.triSteps += i + 1; L.cur = DONE; L.sp = 0; return i + 1; } L.best[k] = d; } }
    L.triSteps += c;
    L.cur = lanePop(L);
    return c;
}

struct Stats { double warpNodeSteps = 0, warpTriSteps = 0, laneNodeSteps = 0, laneTriSteps = 0, warps = 0, maxLaneNode = 0; };

// while-while in lockstep: inner loop runs while ANY lane is at an internal node; then every lane at a leaf processes it
static void simWhileWhile(Lane2* L, int n, Stats& s) {
    for (int i = 0; i < n; i++) { laneBegin(L[i]); if (g_seg) L[i].best[0] = g_seg[i]; }
    for (;;) {
        bool any = false;
        for (;;) {
            bool stepped = false;
            for (int i = 0; i < n; i++) stepped |= laneNodeStep(L[i]);
            if (!stepped) break;
            s.warpNodeSteps++; any = true;
        }
        int mx = 0;
        for (int i = 0; i < n; i++) mx = std::max(mx, laneLeaf(L[i]));
        s.warpTriSteps += mx;
        if (mx) any = true;
        if (!any) break;
    }
    double mxn = 0;
    for (int i = 0; i < n; i++) { s.laneNodeSteps += L[i].nodeSteps; s.laneTriSteps += L[i].triSteps; mxn = std::max(mxn, (double)L[i].nodeSteps); }
    s.maxLaneNode += mxn;
    s.warps++;
}

// if-if in lockstep: every iteration each lane does ONE node step or ONE leaf; an iteration costs a node step when any lane
// was at a node plus max-triangles when any lane was at a leaf
static void simIfIf(Lane2* L, int n, Stats& s) {
    for (int i = 0; i < n; i++) { laneBegin(L[i]); if (g_seg) L[i].best[0] = g_seg[i]; }
    for (;;) {
        bool anyNode = false; int mx = 0;
        for (int i = 0; i < n; i++) {
            if (L[i].cur == DONE) continue;
            if (L[i].cur >= 0) { laneNodeStep(L[i]); anyNode = true; }
            else mx = std::max(mx, laneLeaf(L[i]));
        }
        if (!anyNode && !mx) break;
        s.warpNodeSteps += anyNode; s.warpTriSteps += mx;
    }
    double mxn = 0;
    for (int i = 0; i < n; i++) { s.laneNodeSteps += L[i].nodeSteps; s.laneTriSteps += L[i].triSteps; mxn = std::max(mxn, (double)L[i].nodeSteps); }
    s.maxLaneNode += mxn; s.warps++;
}

// any-hit variants: a lane stops at the first triangle hit closer than its segment length (best[0] = length, never updated)


// warp packet: ONE stack for the warp; a child is entered when any lane's ray(s) hit it; order by the min entry distance
static void simPacket(Lane2* L, int n, Stats& s) {
    for (int i = 0; i < n; i++) { laneBegin(L[i]); if (g_seg) L[i].best[0] = g_seg[i]; }
    static thread_local int stack[256]; static thread_local float stackT[256];
    int sp = 0, cur = root;
    auto limitAll = [&]() { float m = 0; for (int i = 0; i < n; i++) m = fmaxf(m, laneLimit(L[i])); return m; };
    for (;;) {
        while (cur >= 0 && cur != DONE) {
            const Node& nd = nodes[cur];
            s.warpNodeSteps++;
            bool hL = false, hR = false; float tL = 3e38f, tR = 3e38f;
            for (int i = 0; i < n; i++)
                for (int k = 0; k < L[i].nr; k++) {
                    float t;
                    if (slab(L[i].r[k], nd.lmin, nd.lmax, L[i].best[k], t)) { hL = true; tL = fminf(tL, t); }
                    if (slab(L[i].r[k], nd.rmin, nd.rmax, L[i].best[k], t)) { hR = true; tR = fminf(tR, t); }
                }
            if (hL && hR) { bool ln = tL <= tR; stack[sp] = ln ? nd.right : nd.left; stackT[sp] = ln ? tR : tL; sp++; cur = ln ? nd.left : nd.right; }
            else if (hL) cur = nd.left;
            else if (hR) cur = nd.right;
            else { cur = DONE; float lim = limitAll(); while (sp > 0) { --sp; if (stackT[sp] <= lim) { cur = stack[sp]; break; } } }
        }
        if (cur == DONE) break;
        int f = leafFirst(cur), c = leafCount(cur);
        for (int j = 0; j < c; j++)
            for (int i = 0; i < n; i++)
                for (int k = 0; k < L[i].nr; k++) { float d; if (triHit(L[i].r[k], tris[f + j], d) && d < L[i].best[k]) L[i].best[k] = d; }
        s.warpTriSteps += c;
        cur = DONE; float lim = limitAll(); while (sp > 0) { --sp; if (stackT[sp] <= lim) { cur = stack[sp]; break; } }
    }
    s.laneNodeSteps += s.warpNodeSteps * 0; s.warps++;
}


// warp packet driven by the TILE FRUSTUM: a node's child is entered when its box touches the frustum of the 8x4-pixel tile
// (4 side planes through the eye, p-vertex test) and is not beyond every lane's limit; leaf children are optionally confirmed
// by the lanes' own slab tests before their triangles are offered.  laneNodeSteps counts the confirm tests.
struct Frustum { float n[4][3], view[3], o[3]; };
static inline bool frustumBox(const Frustum& F, const float* lo, const float* hi, float limit, float& nearD) {
    for (int p = 0; p < 4; p++) {
        float d = 0;
        for (int a = 0; a < 3; a++) d += F.n[p][a] * ((F.n[p][a] > 0 ? hi[a] : lo[a]) - F.o[a]);
        if (d < -1e-4f) return false;
    }
    float fd = 0, nd = 0;
    for (int a = 0; a < 3; a++) { fd += F.view[a] * ((F.view[a] > 0 ? hi[a] : lo[a]) - F.o[a]); nd += F.view[a] * ((F.view[a] > 0 ? lo[a] : hi[a]) - F.o[a]); }
    if (fd < 0) return false;
    nearD = fmaxf(nd, 0.f);
    return nearD <= limit;
}
static void simFrustum(Lane2* L, int n, const Frustum& F, bool confirm, Stats& s) {
    for (int i = 0; i < n; i++) laneBegin(L[i]);
    static thread_local int stack[256]; static thread_local float stackT[256];
    int sp = 0, cur = root;
    auto limitAll = [&]() { float m = 0; for (int i = 0; i < n; i++) m = fmaxf(m, laneLimit(L[i])); return m; };
    auto confirmLeaf = [&](const float* lo, const float* hi) {
        s.laneNodeSteps++;
        for (int i = 0; i < n; i++)
            for (int k = 0; k < L[i].nr; k++) { float t; if (slab(L[i].r[k], lo, hi, L[i].best[k], t)) return true; }
        return false;
    };
    for (;;) {
        while (cur >= 0 && cur != DONE) {
            const Node& nd = nodes[cur];
            s.warpNodeSteps++;
            float lim = limitAll();
            float tL = 0, tR = 0;
            bool hL = frustumBox(F, nd.lmin, nd.lmax, lim, tL), hR = frustumBox(F, nd.rmin, nd.rmax, lim, tR);
            if (confirm) {
                if (hL && nd.left < 0) hL = confirmLeaf(nd.lmin, nd.lmax);
                if (hR && nd.right < 0) hR = confirmLeaf(nd.rmin, nd.rmax);
            }
            if (hL && hR) { bool ln = tL <= tR; stack[sp] = ln ? nd.right : nd.left; stackT[sp] = ln ? tR : tL; sp++; cur = ln ? nd.left : nd.right; }
            else if (hL) cur = nd.left;
            else if (hR) cur = nd.right;
            else { cur = DONE; while (sp > 0) { --sp; if (stackT[sp] <= lim) { cur = stack[sp]; break; } } }
        }
        if (cur == DONE) break;
        int f = leafFirst(cur), c = leafCount(cur);
        for (int j = 0; j < c; j++)
            for (int i = 0; i < n; i++)
                for (int k = 0; k < L[i].nr; k++) { float d; if (triHit(L[i].r[k], tris[f + j], d) && d < L[i].best[k]) L[i].best[k] = d; }
        s.warpTriSteps += c;
        cur = DONE; float lim = limitAll(); while (sp > 0) { --sp; if (stackT[sp] <= lim) { cur = stack[sp]; break; } }
    }
    s.warps++;
}

// per-lane wide-tree walk (sorted children), lockstep while-while
struct LaneW { Ray r[2]; int nr; float best[2]; int stack[256]; float stackT[256]; int sp, cur; long nodeSteps, triSteps; };
static bool laneWNode(LaneW& L) {
    if (L.cur < 0 || L.cur == DONE) return false;
    const WNode& w = wnodes[L.cur];
    L.nodeSteps++;
    int idx[8]; float tt[8]; int m = 0;
    for (int c = 0; c < w.n; c++) {
        bool h = false; float tm = 3e38f;
        for (int k = 0; k < L.nr; k++) { float t; if (slab(L.r[k], w.lo[c], w.hi[c], L.best[k], t)) { h = true; tm = fminf(tm, t); } }
        if (h) { idx[m] = c; tt[m] = tm; m++; }
    }
    for (int i = 1; i < m; i++) for (int j = i; j > 0 && tt[j] < tt[j - 1]; j--) { std::swap(tt[j], tt[j - 1]); std::swap(idx[j], idx[j - 1]); }
    for (int i = m - 1; i >= 1; i--) { L.stack[L.sp] = w.child[idx[i]]; L.stackT[L.sp] = tt[i]; L.sp++; }
    if (m) L.cur = w.child[idx[0]];
    else {
        L.cur = DONE; float lim = L.best[0]; for (int k = 1; k < L.nr; k++) lim = fmaxf(lim, L.best[k]);
        while (L.sp > 0) { --L.sp; if (L.stackT[L.sp] <= lim) { L.cur = L.stack[L.sp]; break; } }
    }
    return true;
}
static int laneWLeaf(LaneW& L) {
    if (L.cur >= 0) return 0;
    int f = leafFirst(L.cur), c = leafCount(L.cur);
    for (int i = 0; i < c; i++) for (int k = 0; k < L.nr; k++) { float d; if (triHit(L.r[k], tris[f + i], d) && d < L.best[k]) L.best[k] = d; }
    L.triSteps += c;
    L.cur = DONE; float lim = L.best[0]; for (int k = 1; k < L.nr; k++) lim = fmaxf(lim, L.best[k]);
    while (L.sp > 0) { --L.sp; if (L.stackT[L.sp] <= lim) { L.cur = L.stack[L.sp]; break; } }
    return c;
}
static void simWide(LaneW* L, int n, int wroot, Stats& s) {
    for (int i = 0; i < n; i++) { L[i].sp = 0; L[i].cur = wroot; L[i].nodeSteps = L[i].triSteps = 0; for (int k = 0; k < L[i].nr; k++) L[i].best[k] = 3e38f; }
    for (;;) {
        bool any = false;
        for (;;) { bool st = false; for (int i = 0; i < n; i++) st |= laneWNode(L[i]); if (!st) break; s.warpNodeSteps++; any = true; }
        int mx = 0; for (int i = 0; i < n; i++) mx = std::max(mx, laneWLeaf(L[i]));
        s.warpTriSteps += mx; if (mx) any = true;
        if (!any) break;
    }
    double mxn = 0;
    for (int i = 0; i < n; i++) { s.laneNodeSteps += L[i].nodeSteps; s.laneTriSteps += L[i].triSteps; mxn = std::max(mxn, (double)L[i].nodeSteps); }
    s.maxLaneNode += mxn; s.warps++;
}

static uint32_t rngState = 12345;
static float frand(uint32_t& s) { s = s * 1664525u + 1013904223u; return (s >> 8) * (1.f / 16777216.f); }

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: travsim dump_dir W H [stride]\n"); return 1; }
    std::string dir = argv[1];
    int W = atoi(argv[2]), H = atoi(argv[3]);
    int stride = argc > 4 ? atoi(argv[4]) : 4;      // simulate every stride-th warp tile in each direction
    nodes = readAll<Node>(dir + "/nodes.bin");
    tris = readAll<Tri>(dir + "/tris.bin");
    std::vector<float> camv = readAll<float>(dir + "/cam.bin");     // pos right up view aspect tan root
    Cam cam; memcpy(&cam, camv.data(), sizeof(Cam));
    root = (int)camv[14];
    printf("nodes %zu tris %zu root %d  %dx%d\n", nodes.size(), tris.size(), root, W, H);
    int w4 = buildWide(root, 4); size_t n4 = wnodes.size();
    int w8 = buildWide(root, 8);
    printf("wide nodes: bvh4 %zu, bvh8 %zu\n", n4, wnodes.size() - n4);

    auto makeCamRay = [&](int x, int y, float rx, float ry) {
        float ux = 1.f - 2.f * ((x + rx) / W), uy = 1.f - 2.f * ((y + ry) / H);
        float dx = ux * cam.aspect * cam.tanFovY, dy = uy * cam.tanFovY;
        float d[3];
        for (int a = 0; a < 3; a++) d[a] = cam.right[a] * dx + cam.up[a] * dy + cam.view[a];
        float l = 1.f / sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        for (int a = 0; a < 3; a++) d[a] *= l;
        return makeRay(cam.pos, d);
    };
    const int TW = 8, TH = 4;
    Stats sSingle, sPair, sPacket, sPacketPair, sW4, sW8, sW4pair, sW8pair, sSingle16, sPairSub, sIfSingle, sIfPair, sShWW, sShIf, sShPacket, sFr, sFrC;
    std::vector<int> emit;
    for (int i = 0; i < (int)tris.size(); i++) if (tris[i].matId >= 6) emit.push_back(i);
    printf("emitters %zu\n", emit.size());
    int tilesX = W / TW, tilesY = H / TH;
#pragma omp parallel
    {
        Stats a, b, c, d, e, f, g, h, i16, ia, ib, sw, si, sp_, fr, frc, d2_;
        std::vector<Lane2> L(32); std::vector<LaneW> LW(32);
#pragma omp for schedule(dynamic, 8)
        for (int ty = 0; ty < tilesY; ty += stride)
            for (int tx = (ty / stride) % stride; tx < tilesX; tx += stride) {
                uint32_t seed = (uint32_t)(ty * 7919 + tx) * 2654435761u + 1u;
                Ray rc[32], rj[32];
                for (int l = 0; l < 32; l++) {
                    int x = tx * TW + (l & 7), y = ty * TH + (l >> 3);
                    rc[l] = makeCamRay(x, y, .5f, .5f);
                    rj[l] = makeCamRay(x, y, frand(seed), frand(seed));
                }
                for (int l = 0; l < 32; l++) { L[l].nr = 1; L[l].r[0] = rc[l]; }
                simWhileWhile(L.data(), 32, a);
                simPacket(L.data(), 32, c);
                for (int l = 0; l < 32; l++) { L[l].nr = 2; L[l].r[0] = rc[l]; L[l].r[1] = rj[l]; }
                simWhileWhile(L.data(), 32, b);
                simPacket(L.data(), 32, d);
                {
                    Frustum F;
                    memcpy(F.o, cam.pos, 12); memcpy(F.view, cam.view, 12);
                    Ray c00 = makeCamRay(tx * TW, ty * TH, 0, 0), c10 = makeCamRay(tx * TW + TW, ty * TH, 0, 0), c01 = makeCamRay(tx * TW, ty * TH + TH, 0, 0), c11 = makeCamRay(tx * TW + TW, ty * TH + TH, 0, 0);
                    const Ray* cs[5] = {&c00, &c10, &c11, &c01, &c00};
                    Ray cc = makeCamRay(tx * TW + TW / 2, ty * TH + TH / 2, 0, 0);
                    for (int p = 0; p < 4; p++) {
                        const float* u = cs[p]->d; const float* v = cs[p + 1]->d;
                        float nn[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
                        float l = 1.f / sqrtf(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);
                        float sgn = (nn[0] * cc.d[0] + nn[1] * cc.d[1] + nn[2] * cc.d[2]) > 0 ? l : -l;
                        for (int a2 = 0; a2 < 3; a2++) F.n[p][a2] = nn[a2] * sgn;
                    }
                    simFrustum(L.data(), 32, F, false, fr);
                    float chk[32]; for (int l = 0; l < 32; l++) chk[l] = L[l].best[0] + L[l].best[1];
                    simFrustum(L.data(), 32, F, true, frc);
                    simPacket(L.data(), 32, d2_);
                    for (int l = 0; l < 32; l++) if (chk[l] != L[l].best[0] + L[l].best[1]) { fprintf(stderr, "frustum walk result differs\n"); exit(2); }
                }
                simIfIf(L.data(), 32, ib);
                for (int l = 0; l < 32; l++) { L[l].nr = 1; L[l].r[0] = rc[l]; }
                simIfIf(L.data(), 32, ia);
                // shadow rays: from each lane's centre hit to a light picked by a crude 32-candidate RIS (cos cos / d^2)
                {
                    float hitT[32];
                    for (int l = 0; l < 32; l++) hitT[l] = L[l].best[0];
                    int nsh = 0;
                    for (int l = 0; l < 32; l++) {
                        L[l].nr = 1;
                        float x[3]; bool ok = hitT[l] < 1e30f;
                        for (int a2 = 0; a2 < 3; a2++) x[a2] = rc[l].o[a2] + rc[l].d[a2] * hitT[l];
                        float wsum = 0; float bestp[3] = {0, 0, 0};
                        if (ok && !emit.empty()) for (int cnd = 0; cnd < 32; cnd++) {
                            const Tri& t = tris[emit[(size_t)(frand(seed) * emit.size()) % emit.size()]];
                            float u = frand(seed), v = frand(seed); if (u + v > 1) { u = 1 - u; v = 1 - v; }
                            float p[3], e1[3], e2[3], nl[3];
                            for (int a2 = 0; a2 < 3; a2++) { e1[a2] = t.v[1][a2] - t.v[0][a2]; e2[a2] = t.v[2][a2] - t.v[0][a2]; p[a2] = t.v[0][a2] + u * e1[a2] + v * e2[a2]; }
                            nl[0] = e1[1] * e2[2] - e1[2] * e2[1]; nl[1] = e1[2] * e2[0] - e1[0] * e2[2]; nl[2] = e1[0] * e2[1] - e1[1] * e2[0];
                            float dv[3] = {p[0] - x[0], p[1] - x[1], p[2] - x[2]};
                            float d2 = dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2];
                            float cl = -(nl[0] * dv[0] + nl[1] * dv[1] + nl[2] * dv[2]);
                            float w = (cl > 0 && dv[1] > 0) ? cl * dv[1] / (d2 * d2) : 0.f;
                            wsum += w;
                            if (w > 0 && frand(seed) * wsum < w) memcpy(bestp, p, 12);
                        }
                        if (wsum > 0) {
                            float dv[3] = {bestp[0] - x[0], bestp[1] - x[1], bestp[2] - x[2]};
                            float len = sqrtf(dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2]);
                            for (int a2 = 0; a2 < 3; a2++) dv[a2] /= len;
                            float o[3] = {x[0] + dv[0] * 1e-3f, x[1] + dv[1] * 1e-3f, x[2] + dv[2] * 1e-3f};
                            L[l].r[0] = makeRay(o, dv);
                            L[l].best[0] = len - 2e-3f; nsh++;
                        } else { float o[3] = {0, 1000, 0}, dv[3] = {0, 1, 0}; L[l].r[0] = makeRay(o, dv); L[l].best[0] = 1.f; }
                    }
                    // laneBegin resets best -> keep the segment lengths
                    float seg[32]; for (int l = 0; l < 32; l++) seg[l] = L[l].best[0];
                    auto withSeg = [&](auto fn, Stats& st) {
                        // run fn after restoring the segment lengths (laneBegin sets best = inf; emulate by a pre-pass)
                        g_anyhit = true; fn(st); g_anyhit = false;
                    };
                    (void)withSeg; (void)seg; (void)nsh;
                    g_anyhit = true;
                    g_seg = seg; simWhileWhile(L.data(), 32, sw); simIfIf(L.data(), 32, si); simPacket(L.data(), 32, sp_); g_seg = nullptr;
                    g_anyhit = false;
                }
                for (int l = 0; l < 32; l++) { LW[l].nr = 1; LW[l].r[0] = rc[l]; }
                simWide(LW.data(), 32, w4, e);
                simWide(LW.data(), 32, w8, f);
                for (int l = 0; l < 32; l++) { LW[l].nr = 2; LW[l].r[0] = rc[l]; LW[l].r[1] = rj[l]; }
                simWide(LW.data(), 32, w4, g);
                simWide(LW.data(), 32, w8, h);
            }
#pragma omp critical
        {
            auto add = [](Stats& t, const Stats& s) { t.warpNodeSteps += s.warpNodeSteps; t.warpTriSteps += s.warpTriSteps; t.laneNodeSteps += s.laneNodeSteps; t.laneTriSteps += s.laneTriSteps; t.warps += s.warps; t.maxLaneNode += s.maxLaneNode; };
            add(sFr, fr); add(sFrC, frc); add(sIfSingle, ia); add(sIfPair, ib); add(sShWW, sw); add(sShIf, si); add(sShPacket, sp_);
            add(sSingle, a); add(sPair, b); add(sPacket, c); add(sPacketPair, d); add(sW4, e); add(sW8, f); add(sW4pair, g); add(sW8pair, h);
        }
    }
    auto show = [](const char* name, const Stats& s) {
        printf("%-34s warps %7.0f | per warp: node steps %8.1f tri steps %7.1f | per lane: node %7.1f tri %6.1f | lane util node %5.1f/32 tri %5.1f/32 | max-lane node %7.1f\n", name, s.warps,
               s.warpNodeSteps / s.warps, s.warpTriSteps / s.warps, s.laneNodeSteps / s.warps / 32, s.laneTriSteps / s.warps / 32,
               s.laneNodeSteps / std::max(1.0, s.warpNodeSteps), s.laneTriSteps / std::max(1.0, s.warpTriSteps), s.maxLaneNode / s.warps);
    };
    show("bvh2 while-while, centre ray", sSingle);
    show("bvh2 while-while, pair walk", sPair);
    show("bvh2 warp packet, centre ray", sPacket);
    show("bvh2 warp packet, pair", sPacketPair);
    show("bvh2 frustum packet, pair", sFr);
    show("bvh2 frustum packet + leaf confirm", sFrC);
    printf("  leaf confirm tests per warp: %.1f\n", sFrC.laneNodeSteps / sFrC.warps);
    show("bvh4 while-while, centre ray", sW4);
    show("bvh8 while-while, centre ray", sW8);
    show("bvh4 while-while, pair walk", sW4pair);
    show("bvh8 while-while, pair walk", sW8pair);
    show("bvh2 if-if, centre ray", sIfSingle);
    show("bvh2 if-if, pair walk", sIfPair);
    show("shadow rays: while-while", sShWW);
    show("shadow rays: if-if", sShIf);
    show("shadow rays: packet", sShPacket);
    return 0;
}
