/*
 * restir_oracle.cpp -- CPU ORACLE for the ReSTIR DI hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A from-scratch, dependency-free (no glm, no thrust) restatement of the reference's
 * algorithm, written so that every float operation happens in the same order as in the
 * reference's sources when those are compiled WITHOUT fused multiply-add.  Build with
 *     g++ -O2 -fopenmp -ffp-contract=off -fno-fast-math
 * Parity pins: oracle/_ref/libref_harness.so (the reference's own headers compiled with g++,
 * see oracle/Makefile + tests/test_oracle_vs_ref.py) and the fixtures under tests/golden/.
 * The reference ships no tests / golden vectors of its own.
 *
 * Two deliberate deviations from a literal transcription:
 *   (i)  sample2D/3D/4D draw x,y,z,w left-to-right (nvcc device order, sampler.h:51-61);
 *   (ii) spatial reuse is a true two-phase pass (restir.cu:192-196 uses __syncthreads() as if it
 *        were a grid barrier; the racy read is not reproducible).
 * Float->int conversions that the reference executes ON THE GPU use CUDA semantics
 * (saturating, NaN->0); the ones it executes on the host use x86 semantics (INT_MIN).
 *
 * All citations are relative to /root/reference/src.
 */
#include "restir_oracle.h"

#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

/* ---------------------------------------------------------------- vector math (glm 0.9.6.3 semantics) */
struct V2 { float x, y; };
struct V3 { float x, y, z; float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); } };

inline V3 v3(float a) { return {a, a, a}; }
inline V3 v3(float a, float b, float c) { return {a, b, c}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }   /* type_vec3.inl:707 true division */
inline V3 operator/(V3 a, V3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
/* glm::min/max: func_common.inl:409,430  "x < y ? x : y" / "x > y ? x : y" (NaN semantics matter) */
inline float gmin(float x, float y) { return x < y ? x : y; }
inline float gmax(float x, float y) { return x > y ? x : y; }
inline V3 gmin(V3 a, V3 b) { return {gmin(a.x, b.x), gmin(a.y, b.y), gmin(a.z, b.z)}; }
inline V3 gmax(V3 a, V3 b) { return {gmax(a.x, b.x), gmax(a.y, b.y), gmax(a.z, b.z)}; }
/* func_geometric.inl:64-72  tmp = x*y; tmp.x + tmp.y + tmp.z */
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
/* func_geometric.inl:134 */
inline V3 cross(V3 x, V3 y) { return {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y}; }
inline float length(V3 v) { return sqrtf(dot(v, v)); }                          /* :95 */
inline V3 normalize(V3 v) { return v * (1.f / sqrtf(dot(v, v))); }              /* :154 + func_exponential.inl:150 */
inline float mixf(float x, float y, float a) { return x + a * (y - x); }        /* func_common.inl:97-105 */
inline V3 mix(V3 x, V3 y, float a) { return x + (y - x) * a; }                  /* a * (y - x): commutative per component */
inline V3 mix(V3 x, V3 y, V3 a) { return x + a * (y - x); }

const float Pi = 3.1415926535897932384626422832795028841971f;                   /* mathUtil.h:10 */
const float GlmPi = float(3.14159265358979323846264338327950288);               /* glm::pi<float>() */

inline float radians(float deg) { return deg * float(0.01745329251994329576923690768489); } /* func_trigonometric.inl:45 */

/* CUDA cvt.rzi.s32.f32: saturating, NaN -> 0 (the reference runs these conversions on the GPU) */
inline int f2i_cuda(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.f) return INT_MAX;
    if (f <= -2147483648.f) return INT_MIN;
    return (int)f;
}
/* x86 cvttss2si: out-of-range / NaN -> INT_MIN (the reference's host BVH builder, bvh.cpp:83) */
inline int f2i_x86(float f) {
    if (!(f > -2147483904.f && f < 2147483648.f)) return INT_MIN;
    return (int)f;
}

inline float luminance(V3 c) { return dot(c, v3(.2126f, .7152f, .0722f)); }     /* mathUtil.h:119-123 */
inline bool isNanOrInf(float x) { return std::isnan(x) || std::isinf(x); }      /* mathUtil.h:56 */
inline bool hasNanOrInf(V3 v) { return isNanOrInf(v.x) || isNanOrInf(v.y) || isNanOrInf(v.z); }
inline float satDot(V3 a, V3 b) { return gmax(dot(a, b), 0.f); }                /* mathUtil.h:64 */
inline float absDot(V3 a, V3 b) { return fabsf(dot(a, b)); }                    /* mathUtil.h:68 */
inline float pow5(float x) { float x2 = x * x; return x2 * x2 * x; }            /* mathUtil.h:72 */

/* mathUtil.h:190-198 */
inline uint32_t utilhash(uint32_t a) {
    a = (a + 0x7ed55d16) + (a << 12);
    a = (a ^ 0xc761c23c) ^ (a >> 19);
    a = (a + 0x165667b1) + (a << 5);
    a = (a + 0xd3a2646c) ^ (a << 9);
    a = (a + 0xfd7046c5) + (a << 3);
    a = (a ^ 0xb55a4f09) ^ (a >> 16);
    return a;
}

/* ---------------------------------------------------------------- RNG: thrust::minstd_rand + uniform_real_distribution<float>
 * sampler.h:39-48; thrust/random/detail/linear_congruential_engine.inl:45-63 (seed, operator());
 * thrust/random/detail/uniform_real_distribution.inl:63-74.  Thrust 2.x as shipped with CUDA 12.9. */
struct Rng {
    uint32_t x;
    Rng(int looper, int index) {
        /* sampler.h:42  int h = utilhash((1<<31)|(dim<<22)|iter) ^ utilhash(index), dim = 0 */
        uint32_t h = utilhash((1u << 31) | (uint32_t)looper) ^ utilhash((uint32_t)index);
        const uint32_t m = 2147483647u;
        x = h % m;
        if (x == 0) x = 1;         /* c == 0 && s % m == 0  ->  1 % m */
    }
    float next() {
        x = (uint32_t)(((uint64_t)x * 48271ull) % 2147483647ull);
        float r = (float)(x - 1u);                      /* urng() - min, min = 1 */
        r /= (1.f + (float)(2147483646u - 1u));         /* = 2^31 */
        return (r * (1.f - 0.f)) + 0.f;
    }
};

/* ---------------------------------------------------------------- AABB (bvh.h:15-161) */
struct AABB {
    V3 pMin, pMax;
    AABB() : pMin(v3(FLT_MAX)), pMax(v3(-FLT_MAX)) {}
    AABB(V3 a, V3 b) : pMin(a), pMax(b) {}
    AABB(V3 va, V3 vb, V3 vc) : pMin(gmin(gmin(va, vb), vc)), pMax(gmax(gmax(va, vb), vc)) {}   /* :20 */
    AABB grow(V3 p) const { return AABB(gmin(pMin, p), gmax(pMax, p)); }                        /* :26 */
    AABB grow(const AABB& r) const { return AABB(gmin(pMin, r.pMin), gmax(pMax, r.pMax)); }     /* :30 */
    V3 center() const { return (pMin + pMax) * .5f; }                                           /* :47 */
    float surfaceArea() const {                                                                 /* :51 */
        V3 s = pMax - pMin;
        return 2.f * (s.x * s.y + s.y * s.z + s.z * s.x);
    }
    int longestAxis() const {                                                                   /* :59 */
        V3 s = pMax - pMin;
        if (s.x < s.y) return s.y > s.z ? 1 : 2;
        return s.x > s.z ? 0 : 2;
    }
};

struct Ray { V3 origin, direction; };

inline bool between(float x, float lo, float hi) { return x >= lo && x <= hi; }  /* mathUtil.h:31 */
inline bool distMinMax(float a1, float a2, float b1, float b2, float& tMin) {    /* bvh.h:69 */
    tMin = fminf(a1, a2);
    float tMax = fmaxf(b1, b2);
    return tMax >= 0.f && tMax >= tMin;
}
inline bool distMaxMin(float a1, float a2, float b1, float b2, float& tMin) {    /* bvh.h:75 */
    tMin = fmaxf(a1, a2);
    float tMax = fminf(b1, b2);
    return tMax >= 0.f && tMax >= tMin;
}

/* bvh.h:85-157, verbatim control flow */
inline bool aabbIntersect(const AABB& b, const Ray& ray, float& tMin) {
    const float Eps = 1e-6f;
    V3 ori = ray.origin, dir = ray.direction;
    const V3 pMin = b.pMin, pMax = b.pMax;
    if (fabsf(dir.x) > 1.f - Eps) {
        if (between(ori.y, pMin.y, pMax.y) && between(ori.z, pMin.z, pMax.z)) {
            float inv = 1.f / dir.x;
            float t1 = (pMin.x - ori.x) * inv, t2 = (pMax.x - ori.x) * inv;
            return distMinMax(t1, t2, t1, t2, tMin);
        }
        return false;
    } else if (fabsf(dir.y) > 1.f - Eps) {
        if (between(ori.z, pMin.z, pMax.z) && between(ori.x, pMin.x, pMax.x)) {
            float inv = 1.f / dir.y;
            float t1 = (pMin.y - ori.y) * inv, t2 = (pMax.y - ori.y) * inv;
            return distMinMax(t1, t2, t1, t2, tMin);
        }
        return false;
    } else if (fabsf(dir.z) > 1.f - Eps) {
        if (between(ori.x, pMin.x, pMax.x) && between(ori.y, pMin.y, pMax.y)) {
            float inv = 1.f / dir.z;
            float t1 = (pMin.z - ori.z) * inv, t2 = (pMax.z - ori.z) * inv;
            return distMinMax(t1, t2, t1, t2, tMin);
        }
        return false;
    }
    V3 dirInv = {1.f / dir.x, 1.f / dir.y, 1.f / dir.z};
    V3 t1 = (pMin - ori) * dirInv;
    V3 t2 = (pMax - ori) * dirInv;
    V3 tNear = gmin(t1, t2);
    V3 tFar = gmax(t1, t2);
    V3 tDist = tFar - tNear;
    float yz = tFar.z - tNear.y;
    float zx = tFar.x - tNear.z;
    float xy = tFar.y - tNear.x;
    if (fabsf(dir.x) < Eps && tDist.y + tDist.z > yz) return distMaxMin(tNear.y, tNear.z, tFar.y, tFar.z, tMin);
    if (fabsf(dir.y) < Eps && tDist.z + tDist.x > zx) return distMaxMin(tNear.z, tNear.x, tFar.z, tFar.x, tMin);
    if (fabsf(dir.z) < Eps && tDist.x + tDist.y > xy) return distMaxMin(tNear.x, tNear.y, tFar.x, tFar.y, tMin);
    if (tDist.y + tDist.z > yz && tDist.z + tDist.x > zx && tDist.x + tDist.y > xy) {
        return distMaxMin(fmaxf(tNear.x, tNear.y), tNear.z, fminf(tFar.x, tFar.y), tFar.z, tMin);
    }
    return false;
}

/* intersections.h:17-53 */
inline bool intersectTriangle(const Ray& ray, V3 v0, V3 v1, V3 v2, V2& bary, float& dist) {
    V3 e01 = v1 - v0, e02 = v2 - v0;
    V3 ori = ray.origin, dir = ray.direction;
    V3 p = cross(dir, e02);
    float det = dot(p, e01);
    if (fabsf(det) < FLT_EPSILON) return false;
    V3 v0ToOri = ori - v0;
    if (det < 0.f) { det = -det; v0ToOri = -v0ToOri; }
    bary.x = dot(v0ToOri, p);
    if (bary.x < 0.f || bary.x > det) return false;
    V3 perp = cross(v0ToOri, e01);
    bary.y = dot(dir, perp);
    if (bary.y < 0.f || bary.x + bary.y > det) return false;
    float detInv = 1.f / det;
    dist = dot(e02, perp) * detInv;
    bary.x *= detInv; bary.y *= detInv;
    return dist > 0.f;
}

struct MTNode { int prim, box, miss; };               /* bvh.h:163-171 */
struct Alias { float prob; int failId; };              /* sampler.h:63-67 */

/* reservoir, restir.h:7-11,29-117; lightId is oracle-side metadata only (never influences arithmetic) */
struct Sample { V3 Li, wi; float dist; };
struct Resv {
    Sample s; int M; float w; int lightId;
    Resv() { s.Li = v3(0.f); s.wi = v3(0.f); s.dist = 0.f; M = 0; w = 0.f; lightId = -1; }
    void update(const Sample& ns, int nid, float nw, float r) {       /* :38 */
        w += nw; M++;
        if (r * w < nw) { s = ns; lightId = nid; }
    }
    bool invalid() const { return isNanOrInf(w) || w < 0.f; }        /* :51 */
    void checkValidity() { if (invalid()) { w = 0.f; M = 0; } }      /* :55, clear() :46 */
    void merge(const Resv& rhs, float r) {                            /* :61 */
        w += rhs.w; M += rhs.M;
        if (r * w < rhs.w) { s = rhs.s; lightId = rhs.lightId; }
    }
    void clamp(int val) {                                             /* :88 */
        if (M > val) { w *= (float)val / M; M = val; }
    }
    void preClampedMerge(int cap, Resv rhs, float r) {                /* :96 */
        if (M > 0) rhs.clamp((cap - 1) * M);
        merge(rhs, r);
    }
};
struct ResvPacked { float Li[3], wi[3], dist; int M; float w; };       /* 36 B reference layout */

} // namespace

/* ================================================================ scene */
struct OrcScene {
    int T = 0;
    std::vector<V3> vertices, normals;
    std::vector<V2> texcoords;
    std::vector<int> materialIds;
    std::vector<OrcMaterial> materials;
    int bvhSize = 0, bvhDepth = 0;
    std::vector<AABB> boxes;
    std::vector<MTNode> nodes[6];
    std::vector<int> lightPrimIds;
    std::vector<V3> lightUnitRadiance;
    std::vector<float> lightPower;
    std::vector<Alias> alias;
    float sumAll = 0.f, sumLightPowerInv = 0.f;
    /* textures (image.h:7-39, linear RGB float) and the environment map (scene.cpp:136-152) */
    struct Tex { int w = 0, h = 0; std::vector<V3> data; };
    std::vector<Tex> textures;
    int envMapTexId = -1;
    std::vector<Alias> envAlias;
    float envSumAll = 0.f;
};

namespace {

/* sampler.h:79-121  DiscreteSampler1D ctor: LIFO stacks, fp32 normalisation */
void buildAlias(std::vector<float> values, std::vector<Alias>& table, float& sumAll) {
    sumAll = 0.f;
    for (float v : values) sumAll += v;
    float sumInv = (float)values.size() / sumAll;
    for (float& v : values) v *= sumInv;
    size_t n = values.size();
    table.assign(n, Alias{0.f, 0});
    std::vector<Alias> gt(n * 2), ls(n * 2);
    int topGt = 0, topLs = 0;
    for (int i = 0; i < (int)n; i++) {
        float v = values[i];
        (v > 1.f ? gt[topGt++] : ls[topLs++]) = Alias{v, i};
    }
    while (topGt && topLs) {
        Alias g = gt[--topGt];
        Alias l = ls[--topLs];
        table[l.failId] = Alias{l.prob, g.failId};
        g.prob -= (1.f - l.prob);
        (g.prob > 1.f ? gt[topGt++] : ls[topLs++]) = g;
    }
    for (int i = topGt - 1; i >= 0; i--) table[gt[i].failId] = gt[i];
    for (int i = topLs - 1; i >= 0; i--) table[ls[i].failId] = ls[i];
}

/* bvh.cpp:10-131 */
struct PrimInfo { int primId; AABB bound; V3 center; };
struct NodeInfo { bool isLeaf; int primIdOrSize; };

void buildBVH(OrcScene& sc) {
    const int numPrims = sc.T;
    const int bvhSize = numPrims * 2 - 1;
    std::vector<PrimInfo> prim(numPrims);
    std::vector<NodeInfo> info(bvhSize);
    sc.boxes.assign(bvhSize, AABB());
    for (int i = 0; i < numPrims; i++) {
        prim[i].primId = i;
        prim[i].bound = AABB(sc.vertices[i * 3], sc.vertices[i * 3 + 1], sc.vertices[i * 3 + 2]);
        prim[i].center = prim[i].bound.center();
    }
    struct Build { int offset, start, end; };
    std::vector<Build> stack(bvhSize);
    int top = 0;
    stack[top++] = {0, 0, numPrims - 1};
    const int NB = 16;
    int depth = 0;
    std::vector<PrimInfo> temp;
    while (top) {
        depth = depth > top ? depth : top;
        top--;
        int offset = stack[top].offset, start = stack[top].start, end = stack[top].end;
        int n = end - start + 1;
        int nodeSize = n * 2 - 1;
        bool isLeaf = nodeSize == 1;
        info[offset] = {isLeaf, isLeaf ? prim[start].primId : nodeSize};
        AABB nodeBound, centerBound;
        for (int i = start; i <= end; i++) {
            nodeBound = nodeBound.grow(prim[i].bound);
            centerBound = centerBound.grow(prim[i].center);
        }
        sc.boxes[offset] = nodeBound;
        if (isLeaf) continue;
        int axis = centerBound.longestAxis();
        if (nodeSize == 2) {                                        /* bvh.cpp:65 (sic: 2 is unreachable, sizes are odd) */
            if (prim[start].center[axis] > prim[end].center[axis]) std::swap(prim[start], prim[end]);
            sc.boxes[offset + 1] = prim[start].bound;
            sc.boxes[offset + 2] = prim[end].bound;
            info[offset + 1] = {true, prim[start].primId};
            info[offset + 2] = {true, prim[end].primId};
        }
        AABB bucketBounds[NB];
        int bucketCounts[NB];
        memset(bucketCounts, 0, sizeof(bucketCounts));
        float dimMin = centerBound.pMin[axis], dimMax = centerBound.pMax[axis];
        auto bucketOf = [&](float c) {
            int b = f2i_x86((c - dimMin) / (dimMax - dimMin) * NB);
            return b < 0 ? 0 : (b > NB - 1 ? NB - 1 : b);         /* glm::clamp = min(max(x,lo),hi) */
        };
        for (int i = start; i <= end; i++) {
            int bid = bucketOf(prim[i].center[axis]);
            bucketBounds[bid] = bucketBounds[bid].grow(prim[i].bound);
            bucketCounts[bid]++;
        }
        AABB lB[NB], rB[NB];
        int countPrefix[NB];
        lB[0] = bucketBounds[0];
        rB[NB - 1] = bucketBounds[NB - 1];
        countPrefix[0] = bucketCounts[0];
        for (int i = 1, j = NB - 2; i < NB; i++, j--) {             /* bvh.cpp:96-100 (sic: not prefix unions) */
            lB[i] = lB[i].grow(bucketBounds[i - 1]);
            rB[j] = rB[j].grow(bucketBounds[j + 1]);
            countPrefix[i] = countPrefix[i - 1] + bucketCounts[i];
        }
        float minSAH = FLT_MAX;
        int divBucket = 0;
        for (int i = 0; i < NB - 1; i++) {
            float SAH = mixf(lB[i].surfaceArea(), rB[i + 1].surfaceArea(), (float)countPrefix[i] / n);
            if (SAH < minSAH) { minSAH = SAH; divBucket = i; }
        }
        temp.assign(prim.begin() + start, prim.begin() + start + n);
        int divPrim = start, divEnd = end;
        for (int i = 0; i < n; i++) {
            int bid = bucketOf(temp[i].center[axis]);
            (bid <= divBucket ? prim[divPrim++] : prim[divEnd--]) = temp[i];
        }
        divPrim = divPrim - 1;
        divPrim = divPrim < start ? start : (divPrim > end - 1 ? end - 1 : divPrim);
        int lSize = 2 * (divPrim - start + 1) - 1;
        stack[top++] = {offset + 1 + lSize, divPrim + 1, end};
        stack[top++] = {offset + 1, start, divPrim};
    }
    sc.bvhSize = bvhSize;
    sc.bvhDepth = depth;
    /* bvh.cpp:133-201 buildMTBVH */
    std::vector<int> st(bvhSize);
    for (int i = 0; i < 6; i++) {
        auto& nodes = sc.nodes[i];
        nodes.assign(bvhSize, MTNode{0, 0, 0});
        int stTop = 0, idNew = 0;
        st[stTop++] = 0;
        while (stTop) {
            int orig = st[--stTop];
            bool leaf = info[orig].isLeaf;
            int nodeSize = leaf ? 1 : info[orig].primIdOrSize;
            nodes[idNew] = {leaf ? info[orig].primIdOrSize : -1, orig, idNew + nodeSize};
            idNew++;
            if (leaf) continue;
            bool leftLeaf = info[orig + 1].isLeaf;
            int leftSize = leftLeaf ? 1 : info[orig + 1].primIdOrSize;
            int left = orig + 1, right = orig + 1 + leftSize;
            int dim = i / 2;
            bool lesser = i & 1;
            if ((sc.boxes[left].center()[dim] < sc.boxes[right].center()[dim]) ^ lesser) std::swap(left, right);
            st[stTop++] = right;
            st[stTop++] = left;
        }
    }
}

inline float triangleArea(V3 v0, V3 v1, V3 v2) { return length(cross(v1 - v0, v2 - v0)) * .5f; }   /* mathUtil.h:86 */
inline V3 triangleNormal(V3 v0, V3 v1, V3 v2) { return normalize(cross(v1 - v0, v2 - v0)); }       /* mathUtil.h:90 */

struct Isect { int primId, matId; V3 pos, norm; V2 uv; V3 wo; };

/* scene.h:101-119 */
inline int mtbvhId(V3 dir) {
    V3 a = {fabsf(dir.x), fabsf(dir.y), fabsf(dir.z)};
    if (a.x > a.y) {
        if (a.x > a.z) return dir.x > 0 ? 0 : 1;
        return dir.z > 0 ? 4 : 5;
    }
    if (a.y > a.z) return dir.y > 0 ? 2 : 3;
    return dir.z > 0 ? 4 : 5;
}

struct TraceStats { uint64_t nodes = 0, tris = 0, rays = 0; };

/* scene.h:245-284 + getIntersecGeomInfo :135-151 */
void sceneIntersect(const OrcScene& sc, const Ray& ray, Isect& is, TraceStats* st = nullptr) {
    float closestDist = FLT_MAX;
    int closestPrim = -1;
    V2 closestBary = {0.f, 0.f};
    const MTNode* nodes = sc.nodes[mtbvhId(-ray.direction)].data();
    int node = 0;
    uint64_t nv = 0, nt = 0;
    while (node != sc.bvhSize) {
        float bd;
        nv++;
        bool hit = aabbIntersect(sc.boxes[nodes[node].box], ray, bd);
        if (hit && bd < closestDist) {
            int prim = nodes[node].prim;
            if (prim != -1) {
                float d; V2 b;
                nt++;
                bool h = intersectTriangle(ray, sc.vertices[prim * 3], sc.vertices[prim * 3 + 1], sc.vertices[prim * 3 + 2], b, d);
                if (h && d < closestDist) { closestDist = d; closestBary = b; closestPrim = prim; }
            }
            node++;
        } else {
            node = nodes[node].miss;
        }
    }
    if (st) { st->nodes += nv; st->tris += nt; st->rays += 1; }
    if (closestPrim != -1) {
        int p = closestPrim;
        V3 va = sc.vertices[p * 3], vb = sc.vertices[p * 3 + 1], vc = sc.vertices[p * 3 + 2];
        V3 na = sc.normals[p * 3], nb = sc.normals[p * 3 + 1], nc = sc.normals[p * 3 + 2];
        V2 ta = sc.texcoords[p * 3], tb = sc.texcoords[p * 3 + 1], tc = sc.texcoords[p * 3 + 2];
        float bx = closestBary.x, by = closestBary.y, bz = 1.f - bx - by;
        is.pos = vb * bx + vc * by + va * bz;
        is.norm = normalize(nb * bx + nc * by + na * bz);
        is.uv = {tb.x * bx + tc.x * by + ta.x * bz, tb.y * bx + tc.y * by + ta.y * bz};
        is.matId = sc.materialIds[p];
    }
    is.primId = closestPrim;
}

/* scene.h:286-316 + intersections.h:12-14 (makeOffsetedRay) + scene.h:165-173 */
struct ShadowStats { uint64_t rays = 0, nodes = 0, maxNodes = 0; };
bool sceneOccluded(const OrcScene& sc, V3 x, V3 y, ShadowStats* st = nullptr) {
    const float Eps = 1e-4f;
    uint64_t visited = 0;
    struct Tally { ShadowStats* st; uint64_t& v; ~Tally() { if (st) { st->rays++; st->nodes += v; if (v > st->maxNodes) st->maxNodes = v; } } } tally{st, visited};
    V3 dir = y - x;
    float dist = length(dir);
    dir = dir / dist;
    Ray ray = {x + dir * 1e-5f, dir};
    dist -= Eps * 2.f;
    const MTNode* nodes = sc.nodes[mtbvhId(-ray.direction)].data();
    int node = 0;
    while (node != sc.bvhSize) {
        float bd;
        visited++;
        bool hit = aabbIntersect(sc.boxes[nodes[node].box], ray, bd);
        if (hit && bd < dist) {
            int prim = nodes[node].prim;
            if (prim != -1) {
                float d; V2 b;
                bool h = intersectTriangle(ray, sc.vertices[prim * 3], sc.vertices[prim * 3 + 1], sc.vertices[prim * 3 + 2], b, d);
                if (h && d < dist) return true;
            }
            node++;
        } else {
            node = nodes[node].miss;
        }
    }
    return false;
}


/* ---------------------------------------------------------------- textures (image.h:41-74, scene.h:68-99, mathUtil.h:134-155) */
inline float fractf(float x) { return x - floorf(x); }                            /* func_common.inl:332 */
inline V3 linearSample(const OrcScene::Tex& t, V2 uv) {                          /* image.h:41-74 */
    const int width = t.w, height = t.h;
    uv = {fractf(uv.x), fractf(uv.y)};
    float fx = uv.x * ((float)width - FLT_MIN) + .5f;
    float fy = uv.y * ((float)height - FLT_MIN) + .5f;
    int ix = f2i_cuda(fractf(fx) > .5f ? fx : fx - 1);
    if (ix < 0) ix += width;
    int iy = f2i_cuda(fractf(fy) > .5f ? fy : fy - 1);
    if (iy < 0) iy += height;
    int ux = ix + 1;
    if (ux >= width) ux -= width;
    int uy = iy + 1;
    if (uy >= height) uy -= height;
    float lx = fractf(fx + .5f);
    float ly = fractf(fy + .5f);
    V3 c1 = mix(t.data[iy * width + ix], t.data[iy * width + ux], lx);
    V3 c2 = mix(t.data[uy * width + ix], t.data[uy * width + ux], lx);
    return mix(c1, c2, ly);
}
inline V3 proceduralTexture(V2 uv) {                                              /* scene.h:68-76 */
    Rng rng(0, 0);
    uint32_t seed = (uint32_t)(f2i_cuda(uv.x * 1024) * 1024 + f2i_cuda(uv.y * 1024));
    rng.x = seed % 2147483647u;
    if (rng.x == 0) rng.x = 1;
    float rx = rng.next();
    float ry = rng.next();
    const float PiTwo = 6.2831853071795864769252867665590057683943f;
    float f = (sinf(uv.x * 10.f * PiTwo + rx * PiTwo) + 1.f) * .5f;
    float g = (sinf(uv.y * 10.f * PiTwo + ry * PiTwo) + 1.f) * .5f;
    return v3(f * g);
}
inline V3 localToWorld(V3 n, V3 v) {                                              /* mathUtil.h:146-155 */
    V3 t = (fabsf(n.y) > 0.9999f) ? v3(0.f, 0.f, 1.f) : v3(0.f, 1.f, 0.f);
    V3 b = normalize(cross(n, t));
    t = cross(b, n);
    /* mat3(t, b, n) * v: type_mat3x3.inl operator*(mat, vec) = m[0][i]*v.x + m[1][i]*v.y + m[2][i]*v.z */
    V3 r = {t.x * v.x + b.x * v.y + n.x * v.z, t.y * v.x + b.y * v.y + n.y * v.z, t.z * v.x + b.z * v.y + n.z * v.z};
    return normalize(r);
}
inline V2 toPlane(V3 v) {                                                         /* mathUtil.h:139-144 (PiInv = 1.f / Pi unparenthesised) */
    return {fractf(atan2f(v.z, v.x) * 1.f / Pi * .5f + 1.f), atan2f(sqrtf(v.x * v.x + v.z * v.z), v.y) * 1.f / Pi};
}
inline V3 toSphere(V2 v) {                                                        /* mathUtil.h:134-137 */
    const float PiTwo = 6.2831853071795864769252867665590057683943f;
    v = {v.x * PiTwo, v.y * Pi};
    return {cosf(v.x) * sinf(v.y), cosf(v.y), sinf(v.x) * sinf(v.y)};
}
/* scene.h:78-99: the material with its maps applied; may replace is.norm (normal map) */
inline OrcMaterial texturedMaterial(const OrcScene& sc, Isect& is) {
    OrcMaterial mat = sc.materials[is.matId];
    if (mat.baseColorMapId != -1) {
        V3 c = mat.baseColorMapId == -2 ? proceduralTexture(is.uv) : linearSample(sc.textures[mat.baseColorMapId], is.uv);
        mat.baseColor[0] = c.x; mat.baseColor[1] = c.y; mat.baseColor[2] = c.z;
    }
    if (mat.metallicMapId > -1) mat.metallic = linearSample(sc.textures[mat.metallicMapId], is.uv).x;
    if (mat.roughnessMapId > -1) mat.roughness = linearSample(sc.textures[mat.roughnessMapId], is.uv).x;
    if (mat.normalMapId != -1) {
        V3 mapped = linearSample(sc.textures[mat.normalMapId], is.uv);
        V3 localNorm = normalize(v3(mapped.x * 1.f - 0.5f, mapped.y * 1.f - 0.5f, mapped.z * 1.f - 0.5f));
        is.norm = localToWorld(is.norm, localNorm);
    }
    return mat;
}
inline V3 envMapLookup(const OrcScene& sc, V3 dir) { return linearSample(sc.textures[sc.envMapTexId], toPlane(dir)); }
inline int envAliasSample(const OrcScene& sc, float r1, float r2) {               /* sampler.h:203-207 on envMapSampler */
    int len = (int)sc.envAlias.size();
    int pass = f2i_cuda((float)len * r1);
    pass = pass < len - 1 ? pass : len - 1;
    Alias d = sc.envAlias[pass];
    return (r2 < d.prob) ? pass : d.failId;
}
/* scene.h:364-375 (and the unoccluded part of :377-392): pixel of the map -> radiance, direction, pdf */
inline float sampleEnvironmentMap(const OrcScene& sc, float r1, float r2, V3& radiance, V3& wi, int& pixId) {
    const OrcScene::Tex& env = sc.textures[sc.envMapTexId];
    pixId = envAliasSample(sc, r1, r2);
    int y = pixId / env.w;
    int x = pixId - y * env.w;
    radiance = env.data[pixId];
    wi = toSphere({(.5f + x) / env.w, (.5f + y) / env.h});
    return luminance(radiance) * sc.sumLightPowerInv * env.w * env.h * 1.f / Pi * 1.f / Pi * .5f;
}

/* sampler.h:203-207 (device lookup; the float->int runs on the GPU) */
inline int aliasSample(const OrcScene& sc, float r1, float r2) {
    int len = (int)sc.alias.size();
    int pass = f2i_cuda((float)len * r1);
    pass = pass < len - 1 ? pass : len - 1;
    Alias d = sc.alias[pass];
    return (r2 < d.prob) ? pass : d.failId;
}

/* scene.h:394-425; r = (x,y,z,w).  lightIdOut for the environment map (:400-403) is (L-1) + pixel id. */
float sampleDirectLightNoVisibility(const OrcScene& sc, V3 pos, const float r[4], V3& radiance, V3& wi, float& dist, int& lightIdOut) {
    if (sc.alias.empty()) return -1.f;
    int lightId = aliasSample(sc, r[0], r[1]);
    lightIdOut = lightId;
    if (lightId == (int)sc.alias.size() - 1 && !sc.envAlias.empty()) {
        dist = 1e10f;
        int pix;
        float pdf = sampleEnvironmentMap(sc, r[2], r[3], radiance, wi, pix);
        lightIdOut = lightId + pix;
        return pdf;
    }
    int prim = sc.lightPrimIds[lightId];
    V3 v0 = sc.vertices[prim * 3], v1 = sc.vertices[prim * 3 + 1], v2 = sc.vertices[prim * 3 + 2];
    /* mathUtil.h:94-100 sampleTriangleUniform(v0,v1,v2, ru = r.z, rv = r.w) */
    float sr = sqrtf(r[3]);
    float u = 1.f - sr;
    float v = r[2] * sr;
    V3 sampled = v1 * u + v2 * v + v0 * (1.f - u - v);
    V3 normal = triangleNormal(v0, v1, v2);
    V3 posToSampled = sampled - pos;
    if (dot(normal, posToSampled) > -1e-6f) return -1.f;           /* SCENE_LIGHT_SINGLE_SIDED */
    float area = triangleArea(v0, v1, v2);
    radiance = sc.lightUnitRadiance[lightId];
    wi = normalize(posToSampled);
    dist = length(posToSampled);
    float power = luminance(radiance) / (area * 2.f * GlmPi);
    /* mathUtil.h:182-185 pdfAreaToSolidAngle(pdf, x = pos, y = sampled, ny = normal) */
    float pdf = power * sc.sumLightPowerInv;
    V3 yx = pos - sampled;
    return pdf * dot(yx, yx) / absDot(normal, normalize(yx));
}

/* scene.h:427-459 (PTDirect's sampler: occlusion BEFORE the facing test) */
float sampleDirectLight(const OrcScene& sc, V3 pos, const float r[4], V3& radiance, V3& wi) {
    if (sc.alias.empty()) return -1.f;
    int lightId = aliasSample(sc, r[0], r[1]);
    if (lightId == (int)sc.alias.size() - 1 && !sc.envAlias.empty()) {    /* :433-435 -> sampleEnvironmentMap :377-392 */
        int pix;
        float pdf = sampleEnvironmentMap(sc, r[2], r[3], radiance, wi, pix);
        if (sceneOccluded(sc, pos, pos + wi * 1e6f)) return -1.f;
        return pdf;
    }
    int prim = sc.lightPrimIds[lightId];
    V3 v0 = sc.vertices[prim * 3], v1 = sc.vertices[prim * 3 + 1], v2 = sc.vertices[prim * 3 + 2];
    float sr = sqrtf(r[3]);
    float u = 1.f - sr;
    float v = r[2] * sr;
    V3 sampled = v1 * u + v2 * v + v0 * (1.f - u - v);
    if (sceneOccluded(sc, pos, sampled)) return -1.f;
    V3 normal = triangleNormal(v0, v1, v2);
    V3 posToSampled = sampled - pos;
    if (dot(normal, posToSampled) > -1e-6f) return -1.f;
    float area = triangleArea(v0, v1, v2);
    radiance = sc.lightUnitRadiance[lightId];
    wi = normalize(posToSampled);
    float power = luminance(radiance) / (area * 2.f * GlmPi);
    float pdf = power * sc.sumLightPowerInv;
    V3 yx = pos - sampled;
    return pdf * dot(yx, yx) / absDot(normal, normalize(yx));
}

/* material.h:218-228 with :122 (lambertian), :171-186 (metallic workflow), :137 (dielectric) */
inline float schlickG(float c, float alpha) { float a = alpha * .5f; return c / (c * (1.f - a) + a); }  /* :62 */
inline float GTR2Distrib(float c, float alpha) {                                                         /* :71 */
    if (c < 1e-6f) return 0.f;
    float aa = alpha * alpha;
    float denom = c * c * (aa - 1.f) + 1.f;
    denom = denom * denom * Pi;
    return aa / denom;
}
V3 materialBSDF(const OrcMaterial& m, V3 baseColor, V3 n, V3 wo, V3 wi) {
    switch (m.type) {
    case 0: return baseColor * 1.f / Pi;                               /* "baseColor * PiInv", PiInv = 1.f / Pi unparenthesised */
    case 1: {
        float alpha = m.roughness * m.roughness;
        V3 h = normalize(wo + wi);
        float cosO = dot(n, wo), cosI = dot(n, wi);
        if (cosI * cosO < 1e-7f) return v3(0.f);
        V3 f0 = mix(v3(.08f), baseColor, m.metallic);
        V3 f = mix(f0, v3(1.f), pow5(1.f - dot(h, wo)));               /* fresnelSchlick :39 */
        float g = schlickG(fabsf(cosO), alpha) * schlickG(fabsf(cosI), alpha);   /* smithG :67 */
        float d = GTR2Distrib(dot(n, h), alpha);
        return mix(baseColor * 1.f / Pi * (1.f - m.metallic), v3(g * d / (4.f * cosI * cosO)), f);
    }
    default: return v3(0.f);
    }
}

/* sceneStructs.h:69-86 Camera::sample (r.z, r.w unused); gbuffer.cu:11-23 uses the same maths with r = (.5,.5) */
Ray cameraRay(const OrcCamera& c, int x, int y, float rx, float ry, float tanFovY) {
    float aspect = (float)c.resolution[0] / c.resolution[1];
    V2 pixelSize = {1.f / (float)c.resolution[0], 1.f / (float)c.resolution[1]};
    V2 scr = {(float)x * pixelSize.x, (float)y * pixelSize.y};
    V2 ruv = {scr.x + pixelSize.x * rx, scr.y + pixelSize.y * ry};
    ruv = {1.f - ruv.x * 2.f, 1.f - ruv.y * 2.f};
    V3 pFocus = v3(ruv.x * aspect * tanFovY, ruv.y * 1.f * tanFovY, 1.f) * c.focalDist;
    V3 dir = pFocus - v3(0.f);
    V3 right = {c.right[0], c.right[1], c.right[2]}, up = {c.up[0], c.up[1], c.up[2]}, view = {c.view[0], c.view[1], c.view[2]};
    /* mat3(right,up,view) * dir : type_mat3x3.inl:487 */
    V3 d = {right.x * dir.x + up.x * dir.y + view.x * dir.z,
            right.y * dir.x + up.y * dir.y + view.y * dir.z,
            right.z * dir.x + up.z * dir.y + view.z * dir.z};
    Ray ray;
    ray.direction = normalize(d);
    V3 pos = {c.position[0], c.position[1], c.position[2]};
    ray.origin = pos + right * 0.f + up * 0.f;
    return ray;
}

/* sceneStructs.h:23-46 getRasterUV / getRasterCoord */
void rasterCoord(const OrcCamera& c, V3 pos, int& ox, int& oy) {
    V3 cp = {c.position[0], c.position[1], c.position[2]};
    V3 view = {c.view[0], c.view[1], c.view[2]};
    V3 dir = normalize(pos - cp);
    float d = 1.f / dot(dir, view);
    V3 q = dir * d;
    const float* m = c.rotationMatInv;
    V3 p = {m[0] * q.x + m[3] * q.y + m[6] * q.z,
            m[1] * q.x + m[4] * q.y + m[7] * q.z,
            m[2] * q.x + m[5] * q.y + m[8] * q.z};
    float aspect = (float)c.resolution[0] / c.resolution[1];
    float tanFovY = tanf(radians(c.fov[1]));
    p = p / v3(aspect * tanFovY, 1.f * tanFovY, 1.f);
    V2 ndc = {-p.x, -p.y};
    ndc = {ndc.x * .5f + .5f, ndc.y * .5f + .5f};
    ox = f2i_cuda((float)c.resolution[0] * ndc.x);
    oy = f2i_cuda((float)c.resolution[1] * ndc.y);
}

} // namespace

/* ================================================================ frame state */
struct OrcFrame {
    const OrcScene* sc;
    int w, h;
    std::vector<V3> albedo, normal[2], radiance;
    std::vector<int> motion, matId[2];
    std::vector<float> depth[2];
    int frameIdx = 0;
    OrcCamera lastCamera;
    bool haveLast = false;
    std::vector<Resv> resv, lastResv, temp;   /* devDirectReservoir, devLastDirectReservoir, devDirectTemp (restir.cu:8-10) */
    std::vector<Resv> unbNext;                /* unbiased mode: results of the running spatial pass */
    bool first = true;
    std::vector<ResvPacked> exportBuf;
    std::vector<int> exportIds;
    TraceStats stats;
    ShadowStats shadow;
    /* per-pixel state carried across the two phases of spatial reuse */
    struct Carry { uint32_t rng; int status; Resv r; V3 n, wo, direct, pos; OrcMaterial mat; };
    std::vector<Carry> carry;
};

extern "C" {

static int g_threads = 0;
void orc_set_threads(int n) {
    g_threads = n;
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    else omp_set_num_threads(omp_get_num_procs());
#endif
}
int orc_get_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

OrcScene* orc_scene_create(int numTris, const float* vertices, const float* normals, const float* texcoords,
                           const int* materialIds, int numMaterials, const OrcMaterial* materials) {
    OrcScene* sc = new OrcScene;
    sc->T = numTris;
    sc->vertices.resize(numTris * 3); sc->normals.resize(numTris * 3); sc->texcoords.resize(numTris * 3);
    memcpy(sc->vertices.data(), vertices, sizeof(float) * 9 * numTris);
    memcpy(sc->normals.data(), normals, sizeof(float) * 9 * numTris);
    if (texcoords) memcpy(sc->texcoords.data(), texcoords, sizeof(float) * 6 * numTris);
    else memset(sc->texcoords.data(), 0, sizeof(float) * 6 * numTris);
    sc->materialIds.assign(materialIds, materialIds + numTris);
    sc->materials.assign(materials, materials + numMaterials);
    /* scene.cpp:159-190 light list (vertices are already in world space) */
    for (int p = 0; p < numTris; p++) {
        const OrcMaterial& m = sc->materials[sc->materialIds[p]];
        if (m.type != 4) continue;
        V3 radianceUnitArea = {m.baseColor[0], m.baseColor[1], m.baseColor[2]};
        float powerUnitArea = luminance(radianceUnitArea) * 2.f * GlmPi;
        float area = triangleArea(sc->vertices[p * 3], sc->vertices[p * 3 + 1], sc->vertices[p * 3 + 2]);
        sc->lightPrimIds.push_back(p);
        sc->lightUnitRadiance.push_back(radianceUnitArea);
        sc->lightPower.push_back(powerUnitArea * area);
    }
    if (!sc->lightPower.empty()) {
        buildAlias(sc->lightPower, sc->alias, sc->sumAll);            /* scene.cpp:154 */
        sc->sumLightPowerInv = 1.f / sc->sumAll;                      /* scene.cpp:493 */
    }
    buildBVH(*sc);                                                    /* scene.cpp:199 */
    return sc;
}
/* Scene::addTexture results + Scene::createLightSampler (scene.cpp:136-157): call once, right after orc_scene_create */
int orc_scene_set_textures(OrcScene* sc, int numTextures, const int* widths, const int* heights, const float* const* rgb, int envMapTexId) {
    sc->textures.resize(numTextures);
    for (int t = 0; t < numTextures; t++) {
        OrcScene::Tex& tx = sc->textures[t];
        tx.w = widths[t]; tx.h = heights[t];
        tx.data.resize((size_t)tx.w * tx.h);
        memcpy(tx.data.data(), rgb[t], sizeof(V3) * tx.data.size());
    }
    for (const OrcMaterial& m : sc->materials) {
        const int ids[4] = {m.baseColorMapId, m.metallicMapId, m.roughnessMapId, m.normalMapId};
        for (int k = 0; k < 4; k++)
            if (ids[k] >= numTextures || ids[k] < (k == 0 ? -2 : -1)) return -1;
    }
    if (envMapTexId >= numTextures) return -1;
    sc->envMapTexId = envMapTexId < 0 ? -1 : envMapTexId;
    if (sc->envMapTexId >= 0) {
        const OrcScene::Tex& env = sc->textures[envMapTexId];
        std::vector<float> pdf((size_t)env.w * env.h);
        for (int i = 0; i < env.h; i++)
            for (int j = 0; j < env.w; j++) {
                int idx = i * env.w + j;
                pdf[idx] = luminance(env.data[idx]) * sinf((.5f + i) / env.h * Pi);       /* scene.cpp:144 */
            }
        buildAlias(pdf, sc->envAlias, sc->envSumAll);
        sc->lightPower.push_back(sc->envSumAll);                                        /* scene.cpp:151 */
        buildAlias(sc->lightPower, sc->alias, sc->sumAll);                              /* scene.cpp:154 */
        sc->sumLightPowerInv = 1.f / sc->sumAll;
    }
    return 0;
}
const void* orc_scene_env_alias(const OrcScene* s, int* lengthOut, float* sumAllOut) {
    if (lengthOut) *lengthOut = (int)s->envAlias.size();
    if (sumAllOut) *sumAllOut = s->envSumAll;
    return s->envAlias.data();
}
void orc_scene_destroy(OrcScene* s) { delete s; }
int orc_scene_bvh_size(const OrcScene* s) { return s->bvhSize; }
int orc_scene_bvh_depth(const OrcScene* s) { return s->bvhDepth; }
const float* orc_scene_boxes(const OrcScene* s) { return (const float*)s->boxes.data(); }
const int* orc_scene_mtbvh(const OrcScene* s, int i) { return (const int*)s->nodes[i].data(); }
int orc_scene_num_lights(const OrcScene* s) { return (int)s->alias.size(); }    /* lightSampler.length: emissive triangles (+1 for an environment map) */
const int* orc_scene_light_prim_ids(const OrcScene* s) { return s->lightPrimIds.data(); }
const float* orc_scene_light_radiance(const OrcScene* s) { return (const float*)s->lightUnitRadiance.data(); }
const void* orc_scene_alias_table(const OrcScene* s) { return s->alias.data(); }
float orc_scene_sum_light_power(const OrcScene* s) { return s->sumAll; }

/* sceneStructs.h:88-102 Camera::update (viewProjection is not used by the hot path and is left untouched) */
void orc_camera_update(OrcCamera* c) {
    float yaw = radians(c->rotation[0]), pitch = radians(c->rotation[1]);
    V3 view;
    view.x = cosf(yaw) * cosf(pitch);
    view.z = sinf(yaw) * cosf(pitch);
    view.y = sinf(pitch);
    view = normalize(view);
    V3 right = normalize(cross(view, v3(0.f, 1.f, 0.f)));
    V3 up = normalize(cross(right, view));
    c->view[0] = view.x; c->view[1] = view.y; c->view[2] = view.z;
    c->right[0] = right.x; c->right[1] = right.y; c->right[2] = right.z;
    c->up[0] = up.x; c->up[1] = up.y; c->up[2] = up.z;
    /* glm::inverse(mat3(right, up, view)) : type_mat3x3.inl:37-58, m[col][row] */
    float m[3][3] = {{right.x, right.y, right.z}, {up.x, up.y, up.z}, {view.x, view.y, view.z}};
    float ood = 1.f / (+m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2])
                       - m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2])
                       + m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]));
    float inv[3][3];
    inv[0][0] = +(m[1][1] * m[2][2] - m[2][1] * m[1][2]) * ood;
    inv[1][0] = -(m[1][0] * m[2][2] - m[2][0] * m[1][2]) * ood;
    inv[2][0] = +(m[1][0] * m[2][1] - m[2][0] * m[1][1]) * ood;
    inv[0][1] = -(m[0][1] * m[2][2] - m[2][1] * m[0][2]) * ood;
    inv[1][1] = +(m[0][0] * m[2][2] - m[2][0] * m[0][2]) * ood;
    inv[2][1] = -(m[0][0] * m[2][1] - m[2][0] * m[0][1]) * ood;
    inv[0][2] = +(m[0][1] * m[1][2] - m[1][1] * m[0][2]) * ood;
    inv[1][2] = -(m[0][0] * m[1][2] - m[1][0] * m[0][2]) * ood;
    inv[2][2] = +(m[0][0] * m[1][1] - m[1][0] * m[0][1]) * ood;
    for (int col = 0; col < 3; col++) for (int row = 0; row < 3; row++) c->rotationMatInv[col * 3 + row] = inv[col][row];
}

OrcFrame* orc_frame_create(const OrcScene* sc, int w, int h) {
    OrcFrame* f = new OrcFrame;
    f->sc = sc; f->w = w; f->h = h;
    size_t P = (size_t)w * h;
    f->albedo.assign(P, v3(0.f)); f->radiance.assign(P, v3(0.f));
    for (int i = 0; i < 2; i++) { f->normal[i].assign(P, v3(0.f)); f->matId[i].assign(P, 0); f->depth[i].assign(P, 0.f); }
    f->motion.assign(P, 0);
    f->resv.assign(P, Resv()); f->lastResv.assign(P, Resv()); f->temp.assign(P, Resv());   /* restir.cu:482-489 memset 0 */
    f->carry.resize(P);
    memset(&f->lastCamera, 0, sizeof(OrcCamera));
    return f;
}
void orc_frame_destroy(OrcFrame* f) { delete f; }
void orc_frame_reset(OrcFrame* f) { f->first = true; }

/* gbuffer.cu:3-73.  lastCamera is uninitialised in the reference before the first GBuffer::update
 * (gbuffer.h:56); here the first render uses the current camera instead (documented deviation). */
void orc_gbuffer_render(OrcFrame* f, const OrcCamera* cam) {
    const OrcScene& sc = *f->sc;
    const int W = cam->resolution[0], H = cam->resolution[1];
    const OrcCamera lastCam = f->haveLast ? f->lastCamera : *cam;
    const float tanFovY = tanf(radians(cam->fov[1]));
    const int cur = f->frameIdx;
    TraceStats total;
#pragma omp parallel
    {
        TraceStats st;
#pragma omp for schedule(dynamic, 4)
        for (int y = 0; y < H; y++) {
            for (int x = 0; x < W; x++) {
                int idx = y * W + x;
                Ray ray = cameraRay(*cam, x, y, .5f, .5f, tanFovY);
                Isect is;
                sceneIntersect(sc, ray, is, &st);
                if (is.primId != -1) {
                    int matId = is.matId;
                    if (sc.materials[is.matId].type == 4) matId = -2;  /* :29-31 */
                    const OrcMaterial m = texturedMaterial(sc, is);    /* :37 (may replace is.norm) */
                    f->albedo[idx] = {m.baseColor[0], m.baseColor[1], m.baseColor[2]};
                    f->normal[cur][idx] = is.norm;
                    f->matId[cur][idx] = matId;
                    f->depth[cur][idx] = length(ray.origin - is.pos);  /* glm::distance(pos, origin) = length(origin - pos) */
                    int lx, ly;
                    rasterCoord(lastCam, is.pos, lx, ly);
                    f->motion[idx] = (lx >= 0 && lx < f->w && ly >= 0 && ly < f->h) ? ly * W + lx : -1;
                } else {
                    f->albedo[idx] = sc.envMapTexId >= 0 ? envMapLookup(sc, ray.direction) : v3(0.f);   /* :58-63 */
                    f->normal[cur][idx] = v3(0.f);
                    f->matId[cur][idx] = -1;
                    f->depth[cur][idx] = 1.f;
                    f->motion[idx] = 0;
                }
            }
        }
#pragma omp critical
        { total.nodes += st.nodes; total.tris += st.tris; total.rays += st.rays; }
    }
    f->stats = total;
}

void orc_gbuffer_update(OrcFrame* f, const OrcCamera* cam) {   /* gbuffer.cu:75-78 */
    f->lastCamera = *cam;
    f->haveLast = true;
    f->frameIdx ^= 1;
}

/* shadow-ray traversal statistics of the last restir_direct: rays, total node visits, max node visits of one ray */
void orc_last_shadow_stats(OrcFrame* f, uint64_t* rays, uint64_t* nodes, uint64_t* maxNodes) { *rays = f->shadow.rays; *nodes = f->shadow.nodes; *maxNodes = f->shadow.maxNodes; }
void orc_last_trace_stats(OrcFrame* f, uint64_t* n, uint64_t* t, uint64_t* r) { *n = f->stats.nodes; *t = f->stats.tris; *r = f->stats.rays; }

/* restir.cu:20-45 */
static Resv findTemporalNeighbor(const OrcFrame* f, const std::vector<Resv>& in, int idx) {
    const int cur = f->frameIdx, last = cur ^ 1;
    int primId = f->matId[cur][idx];
    int lastIdx = f->motion[idx];
    bool diff = false;
    if (lastIdx < 0) diff = true;
    else if (primId <= -1) diff = true;
    else if (f->matId[last][lastIdx] != primId) diff = true;
    else {
        V3 norm = f->normal[cur][idx], lastNorm = f->normal[last][lastIdx];
        float depth = f->depth[cur][idx], pdepth = f->depth[last][lastIdx];
        if (absDot(norm, lastNorm) < .9f || fabsf(pdepth - depth) > depth * .1f) diff = true;
    }
    return diff ? Resv() : in[lastIdx];
}

/* restir.cu:47-85; mathUtil.h:128-132 toConcentricDisk */
static Resv findSpatialNeighborDisk(const OrcFrame* f, const std::vector<Resv>& buf, int x, int y, float rx, float ry, float radius) {
    const int cur = f->frameIdx;
    int idx = y * f->w + x;
    float rr = sqrtf(rx);
    float theta = ry * Pi * 2.0f;
    V2 p = {cosf(theta) * rr * radius, sinf(theta) * rr * radius};
    int px = f2i_cuda((float)x + .5f + p.x);
    int py = f2i_cuda((float)y + .5f + p.y);
    int pidx = py * f->w + px;
    bool diff = false;
    if (px < 0 || px >= f->w || py < 0 || py >= f->h || (px == x && py == y)) diff = true;
    else if (f->matId[cur][pidx] != f->matId[cur][idx]) diff = true;
    else {
        V3 norm = f->normal[cur][idx], pnorm = f->normal[cur][pidx];
        if (dot(norm, pnorm) < .9f) diff = true;
        float depth = f->depth[cur][idx], pdepth = f->depth[cur][pidx];
        if (fabsf(depth - pdepth) > depth * .1f) diff = true;
    }
    return diff ? Resv() : buf[pidx];
}
/* the same tests, returning the neighbour's pixel index (-1: rejected) -- unbiased mode */
static int spatialNeighbourIndex(const OrcFrame* f, int x, int y, float rx, float ry, float radius) {
    const int cur = f->frameIdx;
    int idx = y * f->w + x;
    float rr = sqrtf(rx);
    float theta = ry * Pi * 2.0f;
    V2 p = {cosf(theta) * rr * radius, sinf(theta) * rr * radius};
    int px = f2i_cuda((float)x + .5f + p.x);
    int py = f2i_cuda((float)y + .5f + p.y);
    int pidx = py * f->w + px;
    if (px < 0 || px >= f->w || py < 0 || py >= f->h || (px == x && py == y)) return -1;
    if (f->matId[cur][pidx] != f->matId[cur][idx]) return -1;
    V3 norm = f->normal[cur][idx], pnorm = f->normal[cur][pidx];
    bool diff = dot(norm, pnorm) < .9f;
    float depth = f->depth[cur][idx], pdepth = f->depth[cur][pidx];
    if (fabsf(depth - pdepth) > depth * .1f) diff = true;
    return diff ? -1 : pidx;
}

/* restir.cu:111-231 as two phases; frame bookkeeping restir.cu:418-446 */
/* ---- unbiased mode (OrcParams::unbiased): restatement of the PRODUCT's additional mode, restir_b200/csrc/kernels.cu "unbiased reuse".
 * A reservoir holds the light point y in s.wi and the contribution weight W in s.dist; the reference has no such mode. */
struct UnbPoint { V3 pos, n, wo; OrcMaterial mat; };
static V3 unbIntegrand(const OrcScene& sc, const UnbPoint& q, V3 y, int lightId) {
    if (lightId < 0) return v3(0.f);
    int prim = sc.lightPrimIds[lightId];
    V3 v0 = sc.vertices[prim * 3], v1 = sc.vertices[prim * 3 + 1], v2 = sc.vertices[prim * 3 + 2];
    V3 nl = triangleNormal(v0, v1, v2), Le = sc.lightUnitRadiance[lightId];
    V3 pts = y - q.pos;
    float cl = dot(nl, pts);
    if (cl > -1e-6f) return v3(0.f);
    float d2 = dot(pts, pts);
    float dist = sqrtf(d2);
    V3 wi = pts * (1.f / dist);
    V3 g = Le * materialBSDF(q.mat, v3(1.f), q.n, q.wo, wi) * satDot(q.n, wi);
    return g * (fabsf(dot(nl, wi)) / d2);
}
static float unbTarget(const OrcScene& sc, const UnbPoint& q, V3 y, int lightId) {
    float t = luminance(unbIntegrand(sc, q, y, lightId));
    return (isNanOrInf(t) || t < 0.f) ? 0.f : t;
}
static void unbFinalize(const OrcScene& sc, const UnbPoint& q, Resv& R) {
    float ph = R.lightId >= 0 ? unbTarget(sc, q, R.s.wi, R.lightId) : 0.f;
    R.s.dist = (ph > 0.f && R.M > 0) ? R.w / ((float)R.M * ph) : 0.f;
    if (isNanOrInf(R.s.dist)) R.s.dist = 0.f;
}
static void unbMerge(const OrcScene& sc, const UnbPoint& q, Resv& R, const Resv& N, float rnd) {
    float m = N.lightId >= 0 ? unbTarget(sc, q, N.s.wi, N.lightId) * N.s.dist * (float)N.M : 0.f;
    R.w += m; R.M += N.M;
    if (rnd * R.w < m) { R.s.wi = N.s.wi; R.s.Li = N.s.Li; R.lightId = N.lightId; }
}

void orc_restir_direct(OrcFrame* f, const OrcCamera* cam, const OrcParams* prm, int looper, int iter) {
    const OrcScene& sc = *f->sc;
    const bool unbiased = prm->unbiased != 0;
    const int W = cam->resolution[0], H = cam->resolution[1];
    const float tanFovY = tanf(radians(cam->fov[1]));
    const bool first = f->first;
    const int reuse = prm->reuse;
    std::vector<Resv>& out = f->resv;        /* reservoirOut */
    std::vector<Resv>& in = f->lastResv;     /* reservoirIn  */
    std::vector<Resv>& tmp = f->temp;        /* reservoirTemp */
    /* ---- phase A: :119-192 */
    ShadowStats shTotal;
#pragma omp parallel
    {
    ShadowStats sh;
#pragma omp for schedule(dynamic, 4)
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            int index = y * W + x;
            OrcFrame::Carry& c = f->carry[index];
            Rng rng(looper, index);
            float r4[4];
            for (int k = 0; k < 4; k++) r4[k] = rng.next();
            Ray ray = cameraRay(*cam, x, y, r4[0], r4[1], tanFovY);
            Isect is;
            sceneIntersect(sc, ray, is);
            if (is.primId == -1) {                                             /* :133-138 */
                c.status = 0;
                c.direct = sc.envMapTexId >= 0 ? envMapLookup(sc, ray.direction) : v3(0.f);
                continue;
            }
            const OrcMaterial material = texturedMaterial(sc, is);            /* :140 */
            const V3 baseColor = v3(1.f);                                      /* :141 */
            if (material.type == 4) { c.status = 1; continue; }               /* :143-146 */
            is.wo = -ray.direction;
            bool deltaBSDF = material.type == 2;
            if (!deltaBSDF && dot(is.norm, is.wo) < 0.f) is.norm = -is.norm;
            Resv reservoir;
            for (int i = 0; i < prm->numCandidates; i++) {                     /* :156-169 */
                for (int k = 0; k < 4; k++) r4[k] = rng.next();
                V3 Li = v3(0.f), wi = v3(0.f); float dist = 0.f; int lid = -1;
                float p = sampleDirectLightNoVisibility(sc, is.pos, r4, Li, wi, dist, lid);
                V3 g = Li * materialBSDF(material, baseColor, is.norm, is.wo, wi) * satDot(is.norm, wi);
                float weight = luminance(g / p);
                if (isNanOrInf(weight) || p <= 0.f) weight = 0.f;
                reservoir.update(Sample{Li, wi, dist}, lid, weight, rng.next());
            }
            if (unbiased) {
                UnbPoint q{is.pos, is.norm, is.wo, material};
                if (reservoir.lightId >= 0) reservoir.s.wi = is.pos + reservoir.s.wi * reservoir.s.dist;      /* the light point */
                if (!first && (reuse & 1)) {
                    Resv temporal = findTemporalNeighbor(f, in, index);
                    if (!temporal.invalid()) {
                        float rnd = rng.next();
                        if (reservoir.M > 0) temporal.clamp((prm->temporalCap - 1) * reservoir.M);
                        unbMerge(sc, q, reservoir, temporal, rnd);
                    }
                }
                reservoir.checkValidity();
                unbFinalize(sc, q, reservoir);
                out[index] = reservoir;
                if (reuse & 2) tmp[index] = reservoir;
                c.status = 2; c.rng = rng.x; c.r = reservoir; c.n = is.norm; c.wo = is.wo; c.mat = material; c.pos = is.pos;
                continue;
            }
            Sample s = reservoir.s;
            if (sceneOccluded(sc, is.pos, is.pos + s.wi * s.dist, &sh)) reservoir.w = 0.f;   /* :172-176 */
            if (!first && (reuse & 1)) {                                       /* :180-185 */
                Resv temporal = findTemporalNeighbor(f, in, index);
                if (!temporal.invalid()) reservoir.preClampedMerge(prm->temporalCap, temporal, rng.next());
            }
            Resv tempReservoir = reservoir;                                    /* :188 */
            if (reuse & 2) {
                reservoir.checkValidity();                                     /* :191-192 */
                tmp[index] = reservoir;
            }
            tempReservoir.checkValidity();                                     /* :211-212 */
            out[index] = tempReservoir;
            c.status = 2; c.rng = rng.x; c.r = reservoir; c.n = is.norm; c.wo = is.wo; c.mat = material;
        }
    }
#pragma omp critical
    { shTotal.rays += sh.rays; shTotal.nodes += sh.nodes; if (sh.maxNodes > shTotal.maxNodes) shTotal.maxNodes = sh.maxNodes; }
    }
    f->shadow = shTotal;
    /* ---- phase B: :196-230.  Passes 2..n follow the commented-out block :201-209: the reservoir of the previous pass
     * is published through reservoirTemp for every shaded pixel (no checkValidity), a grid barrier, a fresh 5-neighbour
     * aggregate, and `if (!aggregate.invalid()) reservoir.preClampedMerge<4>(aggregate, sample1D(rng))`. */
    const int passes = (reuse & 2) ? (prm->spatialPasses < 1 ? 1 : prm->spatialPasses) : 0;
    for (int pass = 1; pass <= passes; pass++) {
        if (pass > 1) {
#pragma omp parallel for schedule(static)
            for (int i = 0; i < W * H; i++)
                if (f->carry[i].status == 2) tmp[i] = f->carry[i].r;              /* :203 */
        }
        if (unbiased) f->unbNext.assign((size_t)W * H, Resv());
#pragma omp parallel for schedule(dynamic, 4)
        for (int y = 0; y < H; y++) {
            for (int x = 0; x < W; x++) {
                OrcFrame::Carry& c = f->carry[y * W + x];
                if (c.status != 2) continue;
                Rng rng(0, 0); rng.x = c.rng;
                if (unbiased) {
                    /* the product's k_restir_b_unb: neighbours read the reservoirs published in tmp; results go to a second
                     * buffer (next) so that every pixel of a pass sees the previous stage */
                    const UnbPoint q{c.pos, c.n, c.wo, c.mat};
                    const Rng rng0 = rng;
                    const Resv own = tmp[y * W + x];
                    const int capM = pass == 1 ? 0x7fffffff : 3 * own.M;
                    Resv S = own;
                    S.w = own.lightId >= 0 ? unbTarget(sc, q, own.s.wi, own.lightId) * own.s.dist * (float)own.M : 0.f;
                    for (int i = 0; i < prm->numSpatial; i++) {
                        float rx = rng.next(), ry = rng.next();
                        int pidx = spatialNeighbourIndex(f, x, y, rx, ry, prm->spatialRadius);
                        float rnd = rng.next();
                        if (pidx < 0) continue;
                        Resv N = tmp[pidx];
                        if (N.invalid()) continue;
                        if (N.M > capM) N.M = capM;
                        unbMerge(sc, q, S, N, rnd);
                    }
                    float Z = 0.f;
                    const float phq = S.lightId >= 0 ? unbTarget(sc, q, S.s.wi, S.lightId) : 0.f;
                    if (phq > 0.f) {
                        Z = (float)own.M;
                        Rng r2 = rng0;
                        for (int i = 0; i < prm->numSpatial; i++) {
                            float rx = r2.next(), ry = r2.next();
                            r2.next();
                            int pidx = spatialNeighbourIndex(f, x, y, rx, ry, prm->spatialRadius);
                            if (pidx < 0) continue;
                            Resv N = tmp[pidx];
                            if (N.invalid()) continue;
                            if (N.M > capM) N.M = capM;
                            const OrcFrame::Carry& cn = f->carry[pidx];
                            UnbPoint pn{cn.pos, cn.n, cn.wo, cn.mat};
                            if (cn.status != 2) pn.mat.type = 2;                  /* a pixel phase A did not shade: target 0 */
                            if (unbTarget(sc, pn, S.s.wi, S.lightId) > 0.f) Z += (float)N.M;
                        }
                    }
                    const float Wn = (phq > 0.f && Z > 0.f) ? S.w / (Z * phq) : 0.f;
                    S.s.dist = isNanOrInf(Wn) ? 0.f : Wn;
                    S.w = S.s.dist * phq * (float)S.M;
                    f->unbNext[y * W + x] = S;
                    c.rng = rng.x;
                    continue;
                }
                Resv agg;                                                          /* :87-100 */
                for (int i = 0; i < prm->numSpatial; i++) {
                    float rx = rng.next(), ry = rng.next();
                    Resv sp = findSpatialNeighborDisk(f, tmp, x, y, rx, ry, prm->spatialRadius);
                    if (!sp.invalid()) agg.merge(sp, rng.next());
                }
                if (pass == 1) {
                    if (!agg.invalid() && !c.r.invalid()) c.r.merge(agg, rng.next());        /* :197-199 */
                } else {
                    if (!agg.invalid()) c.r.preClampedMerge(4, agg, rng.next());             /* :206-208 */
                }
                c.rng = rng.x;
            }
        }
        if (unbiased) {
#pragma omp parallel for schedule(static)
            for (int i = 0; i < W * H; i++)
                if (f->carry[i].status == 2) f->carry[i].r = f->unbNext[i];
        }
    }
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            int index = y * W + x;
            OrcFrame::Carry& c = f->carry[index];
            V3 direct = v3(0.f);
            if (c.status == 0) direct = c.direct;
            if (c.status == 1) direct = v3(1.f);
            if (c.status == 2 && unbiased) {
                const Resv& R = c.r;
                const UnbPoint q{c.pos, c.n, c.wo, c.mat};
                if (R.lightId >= 0 && R.s.dist > 0.f && !sceneOccluded(sc, c.pos, R.s.wi)) direct = unbIntegrand(sc, q, R.s.wi, R.lightId) * R.s.dist;
                if (hasNanOrInf(direct)) direct = v3(0.f);
            } else if (c.status == 2) {
                Resv reservoir = c.r;
                const OrcMaterial& material = c.mat;
                Sample s = reservoir.s;                                        /* :216-222 */
                if (!reservoir.invalid()) {
                    V3 LiBSDF = s.Li * materialBSDF(material, v3(1.f), c.n, c.wo, s.wi);
                    direct = LiBSDF / luminance(LiBSDF) * reservoir.w / (float)reservoir.M;
                }
                if (hasNanOrInf(direct)) direct = v3(0.f);
            }
            direct = direct * f->albedo[index];                                /* :229-230 */
            f->radiance[index] = (f->radiance[index] * (float)iter + direct) / (float)(iter + 1);
        }
    }
    std::swap(f->resv, f->lastResv);                                           /* :434 */
    f->first = false;
}

/* pathtrace.cu:279-328 */
void orc_pathtrace_direct(OrcFrame* f, const OrcCamera* cam, int looper, int iter) {
    const OrcScene& sc = *f->sc;
    const int W = cam->resolution[0], H = cam->resolution[1];
    const float tanFovY = tanf(radians(cam->fov[1]));
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            int index = y * W + x;
            V3 direct = v3(0.f);
            Rng rng(looper, index);
            float r4[4];
            for (int k = 0; k < 4; k++) r4[k] = rng.next();
            Ray ray = cameraRay(*cam, x, y, r4[0], r4[1], tanFovY);
            Isect is;
            sceneIntersect(sc, ray, is);
            if (is.primId == -1) {
                if (sc.envMapTexId >= 0) direct = envMapLookup(sc, ray.direction);      /* :295-300 */
            } else {
                const OrcMaterial material = texturedMaterial(sc, is);                  /* :302 */
                V3 baseColor = {material.baseColor[0], material.baseColor[1], material.baseColor[2]};
                if (material.type == 4) direct = baseColor;
                else {
                    is.wo = -ray.direction;
                    bool deltaBSDF = material.type == 2;
                    if (!deltaBSDF && dot(is.norm, is.wo) < 0.f) is.norm = -is.norm;
                    if (!deltaBSDF) {
                        for (int k = 0; k < 4; k++) r4[k] = rng.next();
                        V3 Li = v3(0.f), wi = v3(0.f);
                        float lightPdf = sampleDirectLight(sc, is.pos, r4, Li, wi);
                        if (lightPdf > 0.f)
                            direct = Li * materialBSDF(material, baseColor, is.norm, is.wo, wi) * satDot(is.norm, wi) / lightPdf;
                    }
                }
            }
            f->radiance[index] = (f->radiance[index] * (float)iter + direct) / (float)(iter + 1);
        }
    }
}

const void* orc_frame_buffer(OrcFrame* f, int which) {
    const int cur = f->frameIdx;
    switch (which) {
    case ORC_BUF_ALBEDO: return f->albedo.data();
    case ORC_BUF_NORMAL: return f->normal[cur].data();
    case ORC_BUF_MATID: return f->matId[cur].data();
    case ORC_BUF_DEPTH: return f->depth[cur].data();
    case ORC_BUF_MOTION: return f->motion.data();
    case ORC_BUF_RADIANCE: return f->radiance.data();
    case ORC_BUF_RESERVOIR:
    case ORC_BUF_RESERVOIR_TEMP:
    case ORC_BUF_LIGHT_INDEX: {
        /* after the swap at restir.cu:434 the buffer written by the last frame is lastResv */
        const std::vector<Resv>& src = which == ORC_BUF_RESERVOIR_TEMP ? f->temp : f->lastResv;
        size_t P = src.size();
        if (which == ORC_BUF_LIGHT_INDEX) {
            f->exportIds.resize(P);
            for (size_t i = 0; i < P; i++) f->exportIds[i] = src[i].lightId;
            return f->exportIds.data();
        }
        f->exportBuf.resize(P);
        for (size_t i = 0; i < P; i++) {
            ResvPacked& o = f->exportBuf[i];
            const Resv& r = src[i];
            o.Li[0] = r.s.Li.x; o.Li[1] = r.s.Li.y; o.Li[2] = r.s.Li.z;
            o.wi[0] = r.s.wi.x; o.wi[1] = r.s.wi.y; o.wi[2] = r.s.wi.z;
            o.dist = r.s.dist; o.M = r.M; o.w = r.w;
        }
        return f->exportBuf.data();
    }
    }
    return nullptr;
}

/* sampler.h:79-121 exposed directly (known-answer tests) */
float orc_alias_build(int n, const float* values, void* outTable) {
    std::vector<float> v(values, values + n);
    std::vector<Alias> t;
    float sum = 0.f;
    buildAlias(v, t, sum);
    memcpy(outTable, t.data(), sizeof(Alias) * n);
    return sum;
}

void orc_rng_draws(int looper, int index, int n, float* out) {
    Rng rng(looper, index);
    for (int i = 0; i < n; i++) out[i] = rng.next();
}

int orc_intersect(const OrcScene* sc, const float* o, const float* d, float* out8, int* outMatId) {
    Ray ray = {{o[0], o[1], o[2]}, {d[0], d[1], d[2]}};
    Isect is;
    sceneIntersect(*sc, ray, is);
    if (is.primId != -1) {
        out8[0] = is.pos.x; out8[1] = is.pos.y; out8[2] = is.pos.z;
        out8[3] = is.norm.x; out8[4] = is.norm.y; out8[5] = is.norm.z;
        out8[6] = is.uv.x; out8[7] = is.uv.y;
        *outMatId = is.matId;
    }
    return is.primId;
}
int orc_occluded(const OrcScene* sc, const float* x, const float* y) {
    return sceneOccluded(*sc, {x[0], x[1], x[2]}, {y[0], y[1], y[2]}) ? 1 : 0;
}

} // extern "C"

/* ================================================================ ReSTIR GI (restir.cu:232-416, 448-476; material.h:122-256)
 * ReSTIRIndirectKernel as the reference ships it: commented out at its call site (main.cpp:168) and with Settings::traceDepth = 0
 * (common.cpp:3) a no-op until the GUI raises the depth.  Restated with the depth as an argument.  One deviation: the kernel jumps
 * to WriteSample past the initialisation of primMaterial / primWo / primSampleDelta when the jittered ray leaves the scene or
 * hits an emitter (undefined behaviour if a history sample is then shaded with them); such a pixel writes indirect = 0 here. */
struct BsdfSample { V3 dir, bsdf; float pdf; unsigned type; };
enum { BS_Diffuse = 1, BS_Glossy = 2, BS_Specular = 4, BS_Reflection = 16, BS_Transmission = 32, BS_Invalid = 1 << 15 };   /* material.h:15-24 */

static inline V3 matVec(V3 c0, V3 c1, V3 c2, V3 v) {                 /* type_mat3x3.inl operator*(mat3, vec3) */
    return {c0.x * v.x + c1.x * v.y + c2.x * v.z, c0.y * v.x + c1.y * v.y + c2.y * v.z, c0.z * v.x + c1.z * v.y + c2.z * v.z};
}
static inline V3 reflectG(V3 I, V3 N) { return I - N * dot(N, I) * 2.f; }                     /* func_geometric.inl:176 */
static inline V3 sampleHemisphereCosine(V3 n, float rx, float ry) {                          /* mathUtil.h:128-132, 157-161 */
    float r = sqrtf(rx), theta = ry * Pi * 2.0f;
    V2 d = {cosf(theta) * r, sinf(theta) * r};
    float z = sqrtf(1.f - (d.x * d.x + d.y * d.y));
    return localToWorld(n, v3(d.x, d.y, z));
}
static inline float fresnelExact(float cosIn, float ior) {                                    /* material.h:43-60 (the #if tests a misspelt macro: this branch) */
    if (cosIn < 0) { ior = 1.f / ior; cosIn = -cosIn; }
    float sinIn = sqrtf(1.f - cosIn * cosIn);
    float sinTr = sinIn / ior;
    if (sinTr >= 1.f) return 1.f;
    float cosTr = sqrtf(1.f - sinTr * sinTr);
    float a = (cosIn - ior * cosTr) / (cosIn + ior * cosTr), b = (ior * cosIn - cosTr) / (ior * cosIn + cosTr);
    return (a * a + b * b) * .5f;
}
static inline bool refractM(V3 n, V3 wi, float ior, V3& wt) {                                 /* mathUtil.h:163-180 */
    float cosIn = dot(n, wi);
    if (cosIn < 0) ior = 1.f / ior;
    float sin2In = gmax(0.f, 1.f - cosIn * cosIn);
    float sin2Tr = sin2In / (ior * ior);
    if (sin2Tr >= 1.f) return false;
    float cosTr = sqrtf(1.f - sin2Tr);
    if (cosIn < 0) cosTr = -cosTr;
    wt = normalize(-wi / ior + n * (cosIn / ior - cosTr));
    return true;
}
static inline float GTR2Pdf(V3 n, V3 m, V3 wo, float alpha) {                                 /* material.h:82-85 */
    return GTR2Distrib(dot(n, m), alpha) * schlickG(dot(n, wo), alpha) * absDot(m, wo) / absDot(n, wo);
}
static V3 GTR2Sample(V3 n, V3 wo, float alpha, float rx, float ry) {                          /* material.h:93-112 */
    V3 t0 = (fabsf(n.y) > 0.9999f) ? v3(0.f, 0.f, 1.f) : v3(0.f, 1.f, 0.f);                   /* localRefMatrix, mathUtil.h:146-151 */
    V3 b0 = normalize(cross(n, t0));
    t0 = cross(b0, n);
    const V3 m0 = t0, m1 = b0, m2 = n;                                                        /* columns */
    /* glm::inverse(mat3), type_mat3x3.inl:37-57 (m[c][r]) */
    const float m00 = m0.x, m01 = m0.y, m02 = m0.z, m10 = m1.x, m11 = m1.y, m12 = m1.z, m20 = m2.x, m21 = m2.y, m22 = m2.z;
    float ood = 1.f / (+m00 * (m11 * m22 - m21 * m12) - m10 * (m01 * m22 - m21 * m02) + m20 * (m01 * m12 - m11 * m02));
    V3 i0, i1, i2;                                                                            /* columns of the inverse */
    i0.x = +(m11 * m22 - m21 * m12) * ood; i1.x = -(m10 * m22 - m20 * m12) * ood; i2.x = +(m10 * m21 - m20 * m11) * ood;
    i0.y = -(m01 * m22 - m21 * m02) * ood; i1.y = +(m00 * m22 - m20 * m02) * ood; i2.y = -(m00 * m21 - m20 * m01) * ood;
    i0.z = +(m01 * m12 - m11 * m02) * ood; i1.z = -(m00 * m12 - m10 * m02) * ood; i2.z = +(m00 * m11 - m10 * m01) * ood;
    V3 vh = normalize(matVec(i0, i1, i2, wo) * v3(alpha, alpha, 1.f));
    float lenSq = vh.x * vh.x + vh.y * vh.y;
    V3 t = lenSq > 0.f ? v3(-vh.y, vh.x, 0.f) / sqrtf(lenSq) : v3(1.f, 0.f, 0.f);
    V3 b = cross(vh, t);
    float r = sqrtf(rx), theta = ry * Pi * 2.0f;                                              /* toConcentricDisk */
    V2 p = {cosf(theta) * r, sinf(theta) * r};
    float s = 0.5f * (vh.z + 1.f);
    p.y = (1.f - s) * sqrtf(1.f - p.x * p.x) + s * p.y;
    V3 h = t * p.x + b * p.y + vh * sqrtf(gmax(0.f, 1.f - (p.x * p.x + p.y * p.y)));
    h = v3(h.x * alpha, h.y * alpha, gmax(0.f, h.z));
    return normalize(matVec(m0, m1, m2, h));
}
static float materialPdf(const OrcMaterial& m, V3 n, V3 wo, V3 wi) {                          /* material.h:230-240 */
    switch (m.type) {
    case 0: return satDot(n, wi) * 1.f / Pi;
    case 1: {
        V3 h = normalize(wo + wi);
        return mixf(satDot(n, wi) * 1.f / Pi, GTR2Pdf(n, h, wo, m.roughness * m.roughness) / (4.f * absDot(h, wo)), 1.f / (2.f - m.metallic));
    }
    default: return 0.f;
    }
}
static void materialSample(const OrcMaterial& m, V3 baseColor, V3 n, V3 wo, float rx, float ry, float rz, BsdfSample& s) {   /* material.h:242-256 */
    switch (m.type) {
    case 0:                                                                                   /* :130-135 */
        s.dir = sampleHemisphereCosine(n, rx, ry);
        s.bsdf = baseColor * 1.f / Pi;
        s.pdf = satDot(n, s.dir) * 1.f / Pi;
        s.type = BS_Diffuse | BS_Reflection;
        return;
    case 1: {                                                                                 /* :197-216 */
        float alpha = m.roughness * m.roughness;
        if (rz > (1.f / (2.f - m.metallic))) s.dir = sampleHemisphereCosine(n, rx, ry);
        else {
            V3 h = GTR2Sample(n, wo, alpha, rx, ry);
            s.dir = -reflectG(wo, h);
        }
        if (dot(n, s.dir) < 0.f) s.type = BS_Invalid;
        else {
            s.bsdf = materialBSDF(m, baseColor, n, wo, s.dir);
            s.pdf = materialPdf(m, n, wo, s.dir);
            s.type = BS_Glossy | BS_Reflection;
        }
        return;
    }
    case 2: {                                                                                 /* :145-169 */
        float pdfRefl = fresnelExact(dot(n, wo), m.ior);
        s.bsdf = baseColor;
        if (rz < pdfRefl) {
            s.dir = reflectG(-wo, n);
            s.type = BS_Specular | BS_Reflection;
            s.pdf = 1.f;
        } else {
            if (!refractM(n, wo, m.ior, s.dir)) { s.type = BS_Invalid; return; }
            float eta = m.ior;
            if (dot(n, wo) < 0) eta = 1.f / eta;
            s.bsdf = s.bsdf / (eta * eta);
            s.type = BS_Specular | BS_Transmission;
            s.pdf = 1.f;
        }
        return;
    }
    default: s.type = BS_Invalid;
    }
}
static inline float powerHeuristic(float f, float g) { float f2 = f * f; return f2 / (f2 + g * g); }   /* mathUtil.h:81-84 */

struct IndSample { V3 Lo, xv, nv, xs, ns; IndSample() { Lo = xv = nv = xs = ns = v3(0.f); } bool invalid() const { return luminance(Lo) < 1e-8f; } };   /* restir.h:13-27 */
struct IndResv {                                                                              /* Reservoir<IndirectLiSample>, restir.h:29-117 */
    IndSample s; int M = 0; float w = 0.f;
    void update(const IndSample& ns, float nw, float r) { w += nw; M++; if (r * w < nw) s = ns; }
    bool invalid() const { return isNanOrInf(w) || w < 0.f; }
    void merge(const IndResv& rhs, float r) { w += rhs.w; M += rhs.M; if (r * w < rhs.w) s = rhs.s; }
    void clamp(int val) { if (M > val) { w *= (float)val / M; M = val; } }
};
struct OrcGI {
    OrcFrame* f;
    std::vector<IndResv> resv, lastResv;      /* devIndTemporalReservoir, devIndLastTemporalReservoir (restir.cu:12-13) */
    std::vector<V3> indirect;
    std::vector<float> exportBuf;
    bool first = true;
};

extern "C" {

OrcGI* orc_gi_create(OrcFrame* f) {
    OrcGI* g = new OrcGI;
    const size_t P = (size_t)f->w * f->h;
    g->f = f; g->resv.assign(P, IndResv()); g->lastResv.assign(P, IndResv()); g->indirect.assign(P, v3(0.f));
    return g;
}
void orc_gi_destroy(OrcGI* g) { delete g; }

/* ReSTIRIndirect (restir.cu:448-476) around ReSTIRIndirectKernel (:242-416); reuse bit 0 = temporal (common.h:36-43) */
void orc_restir_indirect(OrcGI* g, const OrcCamera* cam, int looper, int iter, int maxDepth, int reuse) {
    OrcFrame* f = g->f;
    const OrcScene& sc = *f->sc;
    const int W = cam->resolution[0], H = cam->resolution[1];
    const float tanFovY = tanf(radians(cam->fov[1]));
    const int cur = f->frameIdx, last = cur ^ 1;
    const bool first = g->first;
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const int index = y * W + x;
            IndSample smp;
            Rng rng(looper, index);
            float r4[4];
            for (int k = 0; k < 4; k++) r4[k] = rng.next();
            Ray ray = cameraRay(*cam, x, y, r4[0], r4[1], tanFovY);
            Isect is;
            sceneIntersect(sc, ray, is);
            bool shaded = false, primDelta = false;
            float primPdf = 1.f;
            V3 primWo = v3(0.f), primBase = v3(0.f);
            OrcMaterial primMat{};
            if (is.primId != -1) {
                OrcMaterial material = texturedMaterial(sc, is);
                if (material.type != 4) {
                    shaded = true;
                    V3 throughput = v3(1.f);
                    is.wo = -ray.direction;
                    primWo = -ray.direction;
                    primMat = material;
                    primBase = {material.baseColor[0], material.baseColor[1], material.baseColor[2]};
                    for (int depth = 1; depth <= maxDepth; depth++) {
                        const V3 baseColor = {material.baseColor[0], material.baseColor[1], material.baseColor[2]};
                        bool deltaBSDF = material.type == 2;
                        if (material.type != 2 && dot(is.norm, is.wo) < 0.f) is.norm = -is.norm;
                        if (!deltaBSDF && depth > 1) {                                        /* :290-301 */
                            for (int k = 0; k < 4; k++) r4[k] = rng.next();
                            V3 radiance = v3(0.f), wi = v3(0.f);
                            float lightPdf = sampleDirectLight(sc, is.pos, r4, radiance, wi);
                            if (lightPdf > 0.f) {
                                float bsdfPdf = materialPdf(material, is.norm, is.wo, wi);
                                smp.Lo = smp.Lo + throughput * materialBSDF(material, baseColor, is.norm, is.wo, wi) * radiance * satDot(is.norm, wi) / lightPdf *
                                                      powerHeuristic(lightPdf, bsdfPdf);
                            }
                        }
                        float r3[3];
                        for (int k = 0; k < 3; k++) r3[k] = rng.next();
                        BsdfSample bs;
                        bs.dir = bs.bsdf = v3(0.f); bs.pdf = 0.f; bs.type = 0;
                        materialSample(material, baseColor, is.norm, is.wo, r3[0], r3[1], r3[2], bs);
                        if (bs.type == BS_Invalid) break;
                        else if (bs.pdf < 1e-8f) break;
                        const bool deltaSample = (bs.type & BS_Specular) != 0;
                        if (depth > 1) throughput = throughput * (bs.bsdf / bs.pdf * (deltaSample ? 1.f : absDot(is.norm, bs.dir)));
                        else { primPdf = bs.pdf; primDelta = deltaSample; smp.xv = is.pos; smp.nv = is.norm; }
                        ray.origin = is.pos + bs.dir * 1e-5f; ray.direction = bs.dir;         /* makeOffsetedRay, intersections.h:13 */
                        const V3 curPos = is.pos;
                        sceneIntersect(sc, ray, is);
                        is.wo = -ray.direction;
                        if (is.primId == -1) {                                                /* :334-345 */
                            if (sc.envMapTexId >= 0) {
                                V3 radiance = envMapLookup(sc, ray.direction) * throughput;
                                float envPdf = luminance(envMapLookup(sc, ray.direction)) * sc.sumLightPowerInv * sc.textures[sc.envMapTexId].w * sc.textures[sc.envMapTexId].h * .5f;   /* scene.h:358-362 */
                                float weight = deltaSample ? 1.f : powerHeuristic(bs.pdf, envPdf);
                                smp.Lo = smp.Lo + radiance * weight;
                            }
                            break;
                        }
                        material = texturedMaterial(sc, is);
                        if (material.type == 4) {                                             /* :348-373 */
                            if (dot(is.norm, ray.direction) < 0.f) break;                     /* SCENE_LIGHT_SINGLE_SIDED, common.h:6 */
                            V3 radiance = {material.baseColor[0], material.baseColor[1], material.baseColor[2]};
                            const V3 v0 = sc.vertices[is.primId * 3], v1 = sc.vertices[is.primId * 3 + 1], v2 = sc.vertices[is.primId * 3 + 2];
                            float weight = 1.f;
                            if (!(deltaSample || depth == 1)) {
                                float area = length(cross(v1 - v0, v2 - v0)) * .5f;           /* getPrimitiveArea, scene.h:121-126 */
                                V3 yx = curPos - is.pos;                                      /* pdfAreaToSolidAngle, mathUtil.h:182-185 */
                                float lp = luminance(radiance) * sc.sumLightPowerInv * area * dot(yx, yx) / absDot(is.norm, normalize(yx));
                                weight = powerHeuristic(bs.pdf, lp);
                            }
                            smp.Lo = smp.Lo + radiance * throughput * weight;
                            if (depth == 1) { smp.xs = is.pos; smp.ns = is.norm; }
                            break;
                        }
                        if (depth == 1) { smp.xs = is.pos; smp.ns = is.norm; }
                    }
                }
            }
            /* WriteSample (:382-415) */
            IndResv reservoir;
            float sampleWeight = 0.f;
            if (!smp.invalid()) {
                sampleWeight = luminance(smp.Lo / primPdf);
                if (std::isnan(sampleWeight) || sampleWeight < 0.f) sampleWeight = 0.f;
            }
            reservoir.update(smp, sampleWeight, rng.next());
            if (!first && (reuse & 1)) {
                /* findTemporalNeighbor<IndirectReservoir>, restir.cu:20-45 */
                const int primId = f->matId[cur][index], lastIdx = f->motion[index];
                bool diff = false;
                if (lastIdx < 0) diff = true;
                else if (primId <= -1) diff = true;
                else if (f->matId[last][lastIdx] != primId) diff = true;
                else {
                    float depth = f->depth[cur][index], pdepth = f->depth[last][lastIdx];
                    if (absDot(f->normal[cur][index], f->normal[last][lastIdx]) < .9f || fabsf(pdepth - depth) > depth * .1f) diff = true;
                }
                IndResv temp = diff ? IndResv() : g->lastResv[lastIdx];
                if (!temp.invalid()) reservoir.merge(temp, rng.next());
            }
            V3 indirect = v3(0.f);
            const IndSample sample = reservoir.s;
            reservoir.clamp(20);
            if (shaded && !reservoir.invalid()) {
                V3 primWi = normalize(sample.xs - sample.xv);
                indirect = reservoir.s.Lo / luminance(reservoir.s.Lo) * reservoir.w / (float)reservoir.M;
                indirect = indirect * (materialBSDF(primMat, primBase, sample.nv, primWo, primWi) * (primDelta ? 1.f : satDot(sample.nv, primWi)));
            }
            if (hasNanOrInf(indirect)) indirect = v3(0.f);
            g->resv[index] = reservoir;
            g->indirect[index] = (g->indirect[index] * (float)iter + indirect) / (float)(iter + 1);
        }
    std::swap(g->resv, g->lastResv);                                                          /* restir.cu:463 */
    g->first = false;
}
const float* orc_gi_indirect(OrcGI* g) { return &g->indirect[0].x; }
/* the reservoirs written by the last call as P x 17 f32: Lo xv nv xs ns (15), M, weight */
const float* orc_gi_reservoirs(OrcGI* g) {
    const size_t P = g->lastResv.size();
    g->exportBuf.resize(P * 17);
    for (size_t i = 0; i < P; i++) {
        const IndResv& r = g->lastResv[i];
        const V3 v[5] = {r.s.Lo, r.s.xv, r.s.nv, r.s.xs, r.s.ns};
        for (int k = 0; k < 5; k++) { g->exportBuf[i * 17 + 3 * k] = v[k].x; g->exportBuf[i * 17 + 3 * k + 1] = v[k].y; g->exportBuf[i * 17 + 3 * k + 2] = v[k].z; }
        g->exportBuf[i * 17 + 15] = (float)r.M; g->exportBuf[i * 17 + 16] = r.w;
    }
    return g->exportBuf.data();
}

}  // extern "C"

/* ================================================================ image-space filters (denoiser.cu:25-567)
 * Restated only (a __global__ body cannot be compiled by g++; the harness has no counterpart): the pin for these is the
 * reference's own CUDA build, which links denoiser.cu unmodified (oracle/ref_headless_main.cpp, tests/test_ref_cuda.py). */
struct OrcDenoiser {
    OrcFrame* f;
    int kind;                       /* 1 LeveledEAWFilter, 2 SpatioTemporalFilter */
    float sigLumin, sigNormal, sigDepth;
    std::vector<V3> colorOut, tempColor, accumColor[2], accumMoment[2];
    std::vector<float> variance, tempVariance, filteredVariance;
    bool firstTime = true;
    int frameIdx = 0;
};

static const float Gaussian3x3[3][3] = {{.075f, .124f, .075f}, {.124f, .204f, .124f}, {.075f, .124f, .075f}};          /* denoiser.cu:11-15 */
static const float Gaussian5x5[5][5] = {{.0030f, .0133f, .0219f, .0133f, .0030f}, {.0133f, .0596f, .0983f, .0596f, .0133f},
                                        {.0219f, .0983f, .1621f, .0983f, .0219f}, {.0133f, .0596f, .0983f, .0596f, .0133f},
                                        {.0030f, .0133f, .0219f, .0133f, .0030f}};                                    /* :17-23 */

/* Camera::getPosition, sceneStructs.h:48-64 */
static V3 cameraPosition(const OrcCamera& c, int x, int y, float dist, float tanFovY) {
    Ray r = cameraRay(c, x, y, .5f, .5f, tanFovY);
    return r.origin + r.direction * dist;
}

/* waveletFilter, denoiser.cu:64-134 */
static void eawPass(const OrcFrame* f, const OrcCamera& cam, std::vector<V3>& out, const std::vector<V3>& in, float sigDepth, float sigNormal, float sigLumin, int level) {
    const int W = f->w, H = f->h, step = 1 << level, cur = f->frameIdx;
    const float tanFovY = tanf(radians(cam.fov[1]));
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const int p = y * W + x, idP = f->matId[cur][p];
            if (idP <= -1) { out[p] = in[p]; continue; }
            const V3 normP = f->normal[cur][p], colorP = in[p], posP = cameraPosition(cam, x, y, f->depth[cur][p], tanFovY);
            V3 sum = v3(0.f);
            float sumWeight = 0.f;
            for (int i = -2; i <= 2; i++)
                for (int j = -2; j <= 2; j++) {
                    const int qx = x + j * step, qy = y + i * step;
                    if (qx >= W || qy >= H || qx < 0 || qy < 0) continue;
                    const int q = qy * W + qx;
                    if (f->matId[cur][q] != idP) continue;
                    const V3 normQ = f->normal[cur][q], colorQ = in[q], posQ = cameraPosition(cam, qx, qy, f->depth[cur][q], tanFovY);
                    const float wColor = gmin(1.f, expf(-dot(colorP - colorQ, colorP - colorQ) / sigLumin));
                    const float wNorm = gmin(1.f, expf(-dot(normP - normQ, normP - normQ) / sigNormal));
                    const float wPos = gmin(1.f, expf(-dot(posP - posQ, posP - posQ) / sigDepth));
                    const float weight = wColor * wNorm * wPos * Gaussian5x5[i + 2][j + 2];
                    sum = sum + colorQ * weight;
                    sumWeight += weight;
                }
            out[p] = sumWeight == 0.f ? in[p] : sum / sumWeight;
        }
}

/* waveletFilter (SVGF form), denoiser.cu:139-216 */
static void eawSvgfPass(const OrcFrame* f, const OrcCamera& cam, std::vector<V3>& out, const std::vector<V3>& in, std::vector<float>& varOut,
                        const std::vector<float>& varIn, const std::vector<float>& varFiltered, float sigDepth, float sigNormal, float sigLumin, int level) {
    const int W = f->w, H = f->h, step = 1 << level, cur = f->frameIdx;
    const float tanFovY = tanf(radians(cam.fov[1]));
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const int p = y * W + x, idP = f->matId[cur][p];
            if (idP <= -1) { out[p] = in[p]; varOut[p] = varIn[p]; continue; }
            const V3 normP = f->normal[cur][p], colorP = in[p], posP = cameraPosition(cam, x, y, f->depth[cur][p], tanFovY);
            V3 sumColor = v3(0.f);
            float sumVariance = 0.f, sumWeight = 0.f, sumWeight2 = 0.f;
            for (int i = -2; i <= 2; i++)
                for (int j = -2; j <= 2; j++) {
                    const int qx = x + j * step, qy = y + i * step;
                    if (qx >= W || qy >= H || qx < 0 || qy < 0) continue;
                    const int q = qy * W + qx;
                    if (f->matId[cur][q] != idP) continue;
                    const V3 normQ = f->normal[cur][q], colorQ = in[q], posQ = cameraPosition(cam, qx, qy, f->depth[cur][q], tanFovY);
                    const float varQ = varIn[q];
                    const float wPos = expf(-dot(posP - posQ, posP - posQ) / sigDepth) + 1e-4f;
                    const float wNorm = powf(satDot(normP, normQ), sigNormal) + 1e-4f;
                    const float denom = sigLumin * sqrtf(gmax(varFiltered[q], 0.f)) + 1e-4f;
                    const float wColor = expf(-fabsf(luminance(colorP) - luminance(colorQ)) / denom) + 1e-4f;
                    const float weight = wColor * wNorm * wPos * Gaussian5x5[i + 2][j + 2];
                    const float weight2 = weight * weight;
                    sumColor = sumColor + colorQ * weight;
                    sumVariance += varQ * weight2;
                    sumWeight += weight;
                    sumWeight2 += weight2;
                }
            out[p] = sumWeight < FLT_EPSILON ? in[p] : sumColor / sumWeight;
            varOut[p] = sumWeight2 < FLT_EPSILON ? varIn[p] : sumVariance / sumWeight2;
        }
}

/* temporalAccumulate, denoiser.cu:250-305 */
static void temporalAccumulate(const OrcFrame* f, std::vector<V3>& colorOut, const std::vector<V3>& colorLast, std::vector<V3>& momentOut,
                               const std::vector<V3>& momentLast, const std::vector<V3>& colorIn, bool first) {
    const int P = f->w * f->h, cur = f->frameIdx;
    const float Alpha = .2f;
#pragma omp parallel for
    for (int p = 0; p < P; p++) {
        const int id = f->matId[cur][p], lastIdx = f->motion[p];
        bool diff = first;
        if (lastIdx < 0) diff = true;
        else if (id <= -1) diff = true;
        else if (f->matId[cur ^ 1][lastIdx] != id) diff = true;
        else if (fabsf(dot(f->normal[cur][p], f->normal[cur ^ 1][lastIdx])) < .1f) diff = true;
        const V3 color = colorIn[p];
        const float lum = luminance(color);
        if (diff) {
            colorOut[p] = color;
            momentOut[p] = v3(lum, lum * lum, 0.f);
        } else {
            const V3 lastColor = colorLast[lastIdx], lastMoment = momentLast[lastIdx];
            colorOut[p] = mix(lastColor, color, Alpha);
            momentOut[p] = v3(mixf(lastMoment.x, lum, Alpha), mixf(lastMoment.y, lum * lum, Alpha), lastMoment.z + 1.f);
        }
    }
}

/* estimateVariance, denoiser.cu:307-343 */
static void estimateVariance(std::vector<float>& variance, const std::vector<V3>& moment, int W, int H) {
#pragma omp parallel for
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const int p = y * W + x;
            const V3 m = moment[p];
            if (m.z > 3.5f) { variance[p] = m.y - m.x * m.x; continue; }
            float sx = 0.f, sy = 0.f;
            int num = 0;
            for (int i = -1; i <= 1; i++)
                for (int j = -1; j <= 1; j++) {
                    const int qx = x + j, qy = y + i;
                    if (qx < 0 || qx >= W || qy < 0 || qy >= H) continue;
                    sx += moment[qy * W + qx].x; sy += moment[qy * W + qx].y;
                    num++;
                }
            sx /= (float)num; sy /= (float)num;
            variance[p] = sy - sx * sx;
        }
}

/* filterVariance, denoiser.cu:345-370 */
static void filterVariance(std::vector<float>& out, const std::vector<float>& in, int W, int H) {
#pragma omp parallel for
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            float sum = 0.f, sumWeight = 0.f;
            for (int i = -1; i <= 1; i++)
                for (int j = -1; j <= 1; j++) {
                    const int qx = x + i, qy = y + j;
                    if (qx < 0 || qx >= W || qy < 0 || qy >= H) continue;
                    const float weight = Gaussian3x3[i + 1][j + 1];
                    sum += in[qy * W + qx] * weight;
                    sumWeight += weight;
                }
            out[y * W + x] = sum / sumWeight;
        }
}

extern "C" {

OrcDenoiser* orc_denoiser_create(OrcFrame* f, int kind) {
    OrcDenoiser* d = new OrcDenoiser;
    const size_t P = (size_t)f->w * f->h;
    d->f = f; d->kind = kind;
    d->colorOut.assign(P, v3(0.f)); d->tempColor.assign(P, v3(0.f));
    if (kind == 1) { d->sigLumin = 64.f; d->sigNormal = .2f; d->sigDepth = 1.f; }             /* denoiser.cu:455 */
    else {
        d->sigLumin = 4.f; d->sigNormal = 128.f; d->sigDepth = 1.f;                           /* denoiser.cu:488 */
        for (int i = 0; i < 2; i++) { d->accumColor[i].assign(P, v3(0.f)); d->accumMoment[i].assign(P, v3(0.f)); }
        d->variance.assign(P, 0.f); d->tempVariance.assign(P, 0.f); d->filteredVariance.assign(P, 0.f);
    }
    return d;
}
void orc_denoiser_destroy(OrcDenoiser* d) { delete d; }
void orc_denoiser_set_sigmas(OrcDenoiser* d, float sigLumin, float sigNormal, float sigDepth) { d->sigLumin = sigLumin; d->sigNormal = sigNormal; d->sigDepth = sigDepth; }

/* LeveledEAWFilter::filter (denoiser.cu:463-477) / SpatioTemporalFilter::filter (:537-564) on the frame's radiance */
void orc_denoiser_filter(OrcDenoiser* d, const OrcCamera* cam) {
    const OrcFrame* f = d->f;
    if (d->kind == 1) {
        eawPass(f, *cam, d->colorOut, f->radiance, d->sigDepth, d->sigNormal, d->sigLumin, 0);
        for (int level = 1; level <= 4; level++) {
            eawPass(f, *cam, d->tempColor, d->colorOut, d->sigDepth, d->sigNormal, d->sigLumin, level);
            std::swap(d->colorOut, d->tempColor);
        }
        return;
    }
    const int fi = d->frameIdx;
    temporalAccumulate(f, d->accumColor[fi], d->accumColor[fi ^ 1], d->accumMoment[fi], d->accumMoment[fi ^ 1], f->radiance, d->firstTime);
    d->firstTime = false;
    estimateVariance(d->variance, d->accumMoment[fi], f->w, f->h);
    auto wavelet = [&](std::vector<V3>& out, const std::vector<V3>& in, int level) {
        filterVariance(d->filteredVariance, d->variance, f->w, f->h);
        eawSvgfPass(f, *cam, out, in, d->tempVariance, d->variance, d->filteredVariance, d->sigDepth, d->sigNormal, d->sigLumin, level);
    };
    wavelet(d->colorOut, d->accumColor[fi], 0);
    std::swap(d->colorOut, d->accumColor[fi]);
    std::swap(d->tempVariance, d->variance);
    wavelet(d->colorOut, d->accumColor[fi], 1);
    std::swap(d->tempVariance, d->variance);
    for (int level = 2; level <= 4; level++) {
        wavelet(d->tempColor, d->colorOut, level);
        std::swap(d->tempColor, d->colorOut);
        std::swap(d->tempVariance, d->variance);
    }
}
void orc_denoiser_next_frame(OrcDenoiser* d) { d->frameIdx ^= 1; }                               /* denoiser.cu:566-568 */
/* modulateAlbedo (denoiser.cu:218-228, 405-411) on the filtered image */
void orc_denoiser_modulate_albedo(OrcDenoiser* d) {
    const size_t P = d->colorOut.size();
    for (size_t p = 0; p < P; p++) {
        V3 c = d->colorOut[p] / 1.f;
        c = c / (v3(1.f) - c + v3(1e-4f));                                                      /* Math::LDRToHDR, mathUtil.h:40-43 */
        d->colorOut[p] = c * gmax(d->f->albedo[p], v3(0.f));
    }
}
const float* orc_denoiser_color(OrcDenoiser* d) { return &d->colorOut[0].x; }
const float* orc_denoiser_variance(OrcDenoiser* d) { return d->variance.data(); }

}  // extern "C"
