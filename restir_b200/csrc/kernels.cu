// kernels.cu -- sm_100a kernels of the ReSTIR DI frame.
//
//   k_gbuffer     primary pixel-centre ray -> G-buffer                      (gbuffer.cu:3-73)
//   k_restir_a    jittered primary ray, RIS candidates, shadow ray, temporal reuse; in non-spatial modes
//                 also the final shade                                       (restir.cu:119-192, 211-230)
//   k_restir_b    spatial reuse as a true second pass + final shade         (restir.cu:196-230)
//   k_ptdirect    one-sample NEE reference image                            (pathtrace.cu:279-328)
//   k_tonemap     tone-map + gamma + 8-bit pack                             (pathtrace.cu:30-56)
//   k_export_*    device-internal planes -> reference host layouts
//   k_restir_indirect*  ReSTIR GI                                            (restir.cu:242-416; gi_kernels.inl)
//
// Compiled with -fmad=false: every fp32 operation that feeds a discrete decision (hit selection, reservoir
// selection, similarity tests, pixel truncation) is a plain IEEE mul/add/div/sqrt in the reference's order,
// so results are bit-comparable with the CPU oracle.  Traversal walks ONE packed tree in the reference's
// per-ray child order (bvh.cpp:184-188 + scene.h:101-119) with the reference's exact box predicate
// (bvh.h:85-157), which makes the visited-leaf sequence, tie-breaking and pruning identical to the
// reference's 6-copy threaded MTBVH walk.
#include "kernels.h"
#include "camera_dev.h"

namespace rs {

// ------------------------------------------------------------------------------------------------ rays
struct RayT {
    f3 o, d, inv;
    int flags;     // 0: general path. bits0-1: axis with |d| > 1-1e-6 (1 x, 2 y, 3 z); bit2/3/4: |d.x|,|d.y|,|d.z| < 1e-6
    int dim;       // major axis of -d (scene.h:101-119)
    int lesser;    // ordering parity: 1 when -d[dim] <= 0
};

RS_D RayT makeRayT(f3 o, f3 d) {
    RayT r;
    r.o = o; r.d = d;
    r.inv = mk3(1.f / d.x, 1.f / d.y, 1.f / d.z);
    const float Eps = 1e-6f;
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int big = ax > 1.f - Eps ? 1 : (ay > 1.f - Eps ? 2 : (az > 1.f - Eps ? 3 : 0));
    r.flags = big | (ax < Eps ? 4 : 0) | (ay < Eps ? 8 : 0) | (az < Eps ? 16 : 0);
    // getMTBVHId(-ray.direction)
    f3 nd = -d;
    int id;
    if (ax > ay) {
        if (ax > az) id = nd.x > 0 ? 0 : 1;
        else id = nd.z > 0 ? 4 : 5;
    } else {
        if (ay > az) id = nd.y > 0 ? 2 : 3;
        else id = nd.z > 0 ? 4 : 5;
    }
    r.dim = id >> 1;
    r.lesser = id & 1;
    return r;
}

RS_D bool distMinMax(float a1, float a2, float b1, float b2, float& tMin) {   // bvh.h:69
    tMin = fminf(a1, a2);
    float tMax = fmaxf(b1, b2);
    return tMax >= 0.f && tMax >= tMin;
}
RS_D bool distMaxMin(float a1, float a2, float b1, float b2, float& tMin) {   // bvh.h:75
    tMin = fmaxf(a1, a2);
    float tMax = fminf(b1, b2);
    return tMax >= 0.f && tMax >= tMin;
}
RS_D bool between(float x, float lo, float hi) { return x >= lo && x <= hi; }

// A ray with |d_a| < 1e-6 makes the reference drop axis a's slab altogether (bvh.h:137-147), so it walks every box
// of the sheet spanned by the other two axes -- tens of thousands of nodes for the 2-3 such pixels each frame has
// (image centre column / horizon row), a multi-millisecond single-thread tail.  A triangle can only be hit inside
// its box, so requiring the ray's a-coordinate over the box's [tMin,tMax] interval to overlap the box (padded for
// rounding) prunes that sheet without changing which triangles are found.
RS_D bool axisGuard(float o, float d, float lo, float hi, float t0, float t1) {
    float ta = fmaxf(t0, 0.f);
    float a0 = o + ta * d, a1 = o + t1 * d;
    float pad = 1e-5f * (fabsf(o) + fabsf(lo) + fabsf(hi) + 1.f);
    return fmaxf(a0, a1) >= lo - pad && fminf(a0, a1) <= hi + pad;
}

// bvh.h:85-157, all special cases (rare rays: axis-aligned or with a vanishing component)
__device__ __noinline__ bool boxHitSlow(const RayT& r, f3 pMin, f3 pMax, float& tMin) {
    f3 ori = r.o;
    int big = r.flags & 3;
    if (big == 1) {
        if (between(ori.y, pMin.y, pMax.y) && between(ori.z, pMin.z, pMax.z)) {
            float t1 = (pMin.x - ori.x) * r.inv.x, t2 = (pMax.x - ori.x) * r.inv.x;
            return distMinMax(t1, t2, t1, t2, tMin);
        }
        return false;
    } else if (big == 2) {
        if (between(ori.z, pMin.z, pMax.z) && between(ori.x, pMin.x, pMax.x)) {
            float t1 = (pMin.y - ori.y) * r.inv.y, t2 = (pMax.y - ori.y) * r.inv.y;
            return distMinMax(t1, t2, t1, t2, tMin);
        }
        return false;
    } else if (big == 3) {
        if (between(ori.x, pMin.x, pMax.x) && between(ori.y, pMin.y, pMax.y)) {
            float t1 = (pMin.z - ori.z) * r.inv.z, t2 = (pMax.z - ori.z) * r.inv.z;
            return distMinMax(t1, t2, t1, t2, tMin);
        }
        return false;
    }
    f3 t1 = (pMin - ori) * r.inv;
    f3 t2 = (pMax - ori) * r.inv;
    f3 tNear = gmin(t1, t2);
    f3 tFar = gmax(t1, t2);
    f3 tDist = tFar - tNear;
    float yz = tFar.z - tNear.y;
    float zx = tFar.x - tNear.z;
    float xy = tFar.y - tNear.x;
    if ((r.flags & 4) && tDist.y + tDist.z > yz)
        return distMaxMin(tNear.y, tNear.z, tFar.y, tFar.z, tMin) && axisGuard(ori.x, r.d.x, pMin.x, pMax.x, tMin, fminf(tFar.y, tFar.z));
    if ((r.flags & 8) && tDist.z + tDist.x > zx)
        return distMaxMin(tNear.z, tNear.x, tFar.z, tFar.x, tMin) && axisGuard(ori.y, r.d.y, pMin.y, pMax.y, tMin, fminf(tFar.z, tFar.x));
    if ((r.flags & 16) && tDist.x + tDist.y > xy)
        return distMaxMin(tNear.x, tNear.y, tFar.x, tFar.y, tMin) && axisGuard(ori.z, r.d.z, pMin.z, pMax.z, tMin, fminf(tFar.x, tFar.y));
    if (tDist.y + tDist.z > yz && tDist.z + tDist.x > zx && tDist.x + tDist.y > xy)
        return distMaxMin(fmaxf(tNear.x, tNear.y), tNear.z, fminf(tFar.x, tFar.y), tFar.z, tMin);
    return false;
}

// general path of bvh.h:124-156 (no special-case flag set: every 1/d is finite)
RS_D bool boxHit(const RayT& r, f3 pMin, f3 pMax, float& tMin) {
    if (r.flags != 0) return boxHitSlow(r, pMin, pMax, tMin);
    f3 t1 = (pMin - r.o) * r.inv;
    f3 t2 = (pMax - r.o) * r.inv;
    f3 tNear = gmin(t1, t2);
    f3 tFar = gmax(t1, t2);
    f3 tDist = tFar - tNear;
    float yz = tFar.z - tNear.y;
    float zx = tFar.x - tNear.z;
    float xy = tFar.y - tNear.x;
    bool pre = (tDist.y + tDist.z > yz) & (tDist.z + tDist.x > zx) & (tDist.x + tDist.y > xy);
    tMin = fmaxf(fmaxf(tNear.x, tNear.y), tNear.z);
    float tMax = fminf(fminf(tFar.x, tFar.y), tFar.z);
    return pre && tMax >= 0.f && tMax >= tMin;
}

// intersections.h:17-53
// Absolute error bound of the Moller-Trumbore distance of a hit: dist = dot(e02, cross(o - v0, e01)) / det, so the
// rounding error is about |e02| |o - v0| |e01| / |det| ulps; 2^-20 = 16 ulps of safety.
RS_D float triDistError(const RayT& r, f3 v0, f3 v1, f3 v2) {
    f3 e01 = v1 - v0, e02 = v2 - v0, vo = r.o - v0;
    float det = dot(cross(r.d, e02), e01);
    return 9.5367431640625e-7f * sqrtf(dot(e01, e01) * dot(e02, e02) * dot(vo, vo)) / fabsf(det);
}

RS_D bool triHit(const RayT& r, f3 v0, f3 v1, f3 v2, float& bx, float& by, float& dist) {
    f3 e01 = v1 - v0, e02 = v2 - v0;
    f3 p = cross(r.d, e02);
    float det = dot(p, e01);
    if (fabsf(det) < FLT_EPSILON) return false;
    f3 v0ToOri = r.o - v0;
    if (det < 0.f) { det = -det; v0ToOri = -v0ToOri; }
    bx = dot(v0ToOri, p);
    if (bx < 0.f || bx > det) return false;
    f3 perp = cross(v0ToOri, e01);
    by = dot(r.d, perp);
    if (by < 0.f || bx + by > det) return false;
    float detInv = 1.f / det;
    dist = dot(e02, perp) * detInv;
    bx *= detInv; by *= detInv;
    return dist > 0.f;
}

// 32-byte read-only load (sm_100: LDG.E.256): a 64-byte node / light record is two requests per lane instead of four, which
// is what bounds the gather-heavy kernels (k_shadow, k_candidates: L1 request throughput 86-93 % with 16-byte loads)
struct F8 { float4 lo, hi; };
RS_D F8 ldg256(const float4* p) {
    F8 r;
#ifdef RS_HOST_EMU      /* tests/emu: this translation unit compiled by g++ to check the kernels' arithmetic on the CPU; never in the product */
    r.lo = p[0]; r.hi = p[1];
#else
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w) : "l"(p));
#endif
    return r;
}

struct Tri { f3 v0, v1, v2; int matId; int prim; };   // prim = original primitive id

// triangle record fi of the leaf-ordered array
RS_D Tri loadTriFast(const DevScene& s, int fi) {
#ifdef RS_DEBUG_BOUNDS
    if (fi < 0 || fi >= s.numTris) { printf("loadTriFast: bad triangle %d (of %d)\n", fi, s.numTris); __trap(); }
#endif
    const float4* p = s.triGeom + 3 * (size_t)fi;
    float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    Tri t;
    t.v0 = mk3(a.x, a.y, a.z); t.v1 = mk3(a.w, b.x, b.y); t.v2 = mk3(b.z, b.w, c.x);
    t.matId = __float_as_int(c.y);
    t.prim = __float_as_int(c.z);
    return t;
}
// by original primitive id
RS_D Tri loadTri(const DevScene& s, int prim) { return loadTriFast(s, __ldg(s.primToFast + prim)); }

struct Hit { float t, bx, by; int prim; };

// Traversal stack: the first RS_SMEM_STACK entries of every thread live in shared memory, laid out [entry][thread] so
// that a lane always hits its own bank whatever its stack depth (divergent depths cost nothing); deeper entries
// spill to a per-thread local array.
#ifndef RS_SMEM_STACK
#define RS_SMEM_STACK 24       /* (tests/emu also builds with 4: every walk then spills, which exercises the local-array halves of the stacks) */
#endif
#define RS_BLOCK 128
#ifndef RS_MINB_GBUF
#define RS_MINB_GBUF 8      /* __launch_bounds__ minBlocks: 64 registers, 32 warps/SM (A/B on B200: -5 % vs uncapped 72) */
#endif
#ifndef RS_MINB_FUSED
#define RS_MINB_FUSED 6     /* fused G-buffer + phase A kernel: 80 registers */
#endif
#ifndef RS_MINB_RESTIR
#define RS_MINB_RESTIR 8    /* 64 registers instead of 94: -12 % on the 1M-triangle scene despite ~200 B of spills */
#endif
struct Stack {
    int* sRef;      // &smemRef[0][tid]
    float* sT;      // &smemT[0][tid]
    int lRef[RS_STACK_DEPTH - RS_SMEM_STACK];
    float lT[RS_STACK_DEPTH - RS_SMEM_STACK];
    RS_D void push(int sp, int ref, float t) {
#ifdef RS_DEBUG_BOUNDS
        if (sp < 0 || sp >= RS_STACK_DEPTH) { printf("stack overflow sp %d\n", sp); __trap(); }
#endif
        if (sp < RS_SMEM_STACK) { sRef[sp * RS_BLOCK] = ref; sT[sp * RS_BLOCK] = t; }
        else { lRef[sp - RS_SMEM_STACK] = ref; lT[sp - RS_SMEM_STACK] = t; }
    }
    RS_D void pushRef(int sp, int ref) {
        if (sp < RS_SMEM_STACK) sRef[sp * RS_BLOCK] = ref;
        else lRef[sp - RS_SMEM_STACK] = ref;
    }
    RS_D int ref(int sp) const { return sp < RS_SMEM_STACK ? sRef[sp * RS_BLOCK] : lRef[sp - RS_SMEM_STACK]; }
    RS_D float t(int sp) const { return sp < RS_SMEM_STACK ? sT[sp * RS_BLOCK] : lT[sp - RS_SMEM_STACK]; }
};
// any-hit rays only push node references: no distance plane
#define RS_DECLARE_REFSTACK(name)                                      \
    __shared__ int name##_ref[RS_SMEM_STACK][RS_BLOCK];                \
    Stack name;                                                        \
    name.sRef = &name##_ref[0][threadIdx.x];                           \
    name.sT = nullptr
#define RS_DECLARE_STACK(name)                                         \
    __shared__ int name##_ref[RS_SMEM_STACK][RS_BLOCK];                \
    __shared__ float name##_t[RS_SMEM_STACK][RS_BLOCK];                \
    Stack name;                                                        \
    name.sRef = &name##_ref[0][threadIdx.x];                           \
    name.sT = &name##_t[0][threadIdx.x]

// scene.h:245-278 on the packed tree.  Children are visited in the reference's order for this ray; the deferred
// child is re-tested against the current closest distance when popped (= the reference's test on arrival).
RS_D void traceClosestExact(const DevScene& s, const RayT& r, Hit& h, Stack& stack) {
    h.t = FLT_MAX; h.prim = -1; h.bx = 0.f; h.by = 0.f;
    float tRoot;
    if (!(boxHit(r, mk3(s.rootMin[0], s.rootMin[1], s.rootMin[2]), mk3(s.rootMax[0], s.rootMax[1], s.rootMax[2]), tRoot) && tRoot < h.t)) return;
    int sp = 0;
    int cur = s.rootRef;
    for (;;) {
        if (cur < 0) {
            int prim = ~cur;
            Tri t = loadTri(s, prim);
            float bx, by, d;
            if (triHit(r, t.v0, t.v1, t.v2, bx, by, d) && d < h.t) { h.t = d; h.bx = bx; h.by = by; h.prim = prim; }
        } else {
            const float4* np = s.nodes + 4 * (size_t)cur;
            float4 a = __ldg(np), b = __ldg(np + 1), c = __ldg(np + 2);
            int4 l = __ldg((const int4*)(np + 3));
            float tL = 0.f, tR = 0.f;
            bool hL = boxHit(r, mk3(a.x, a.y, a.z), mk3(a.w, b.x, b.y), tL);
            bool hR = boxHit(r, mk3(b.z, b.w, c.x), mk3(c.y, c.z, c.w), tR);
            bool swp = (((l.z >> r.dim) & 1) ^ r.lesser) != 0;
            int first = swp ? l.y : l.x, second = swp ? l.x : l.y;
            float tF = swp ? tR : tL, tS = swp ? tL : tR;
            bool goF = (swp ? hR : hL) && tF < h.t;
            bool goS = (swp ? hL : hR) && tS < h.t;
            if (goF) {
                if (goS) { stack.push(sp, second, tS); sp++; }
                cur = first;
                continue;
            }
            if (goS) { cur = second; continue; }
        }
        bool found = false;
        while (sp > 0) {
            --sp;
            if (stack.t(sp) < h.t) { cur = stack.ref(sp); found = true; break; }
        }
        if (!found) break;
    }
}

// scene.h:286-316.  Any-hit: the answer does not depend on the visiting order.
RS_D bool traceOccludedExact(const DevScene& s, const RayT& r, float dist, Stack& stack) {
    float tRoot;
    if (!(boxHit(r, mk3(s.rootMin[0], s.rootMin[1], s.rootMin[2]), mk3(s.rootMax[0], s.rootMax[1], s.rootMax[2]), tRoot) && tRoot < dist)) return false;
    int sp = 0;
    int cur = s.rootRef;
    for (;;) {
        if (cur < 0) {
            Tri t = loadTri(s, ~cur);
            float bx, by, d;
            if (triHit(r, t.v0, t.v1, t.v2, bx, by, d) && d < dist) return true;
        } else {
            const float4* np = s.nodes + 4 * (size_t)cur;
            float4 a = __ldg(np), b = __ldg(np + 1), c = __ldg(np + 2);
            int4 l = __ldg((const int4*)(np + 3));
            float tL = 0.f, tR = 0.f;
            bool goL = boxHit(r, mk3(a.x, a.y, a.z), mk3(a.w, b.x, b.y), tL) && tL < dist;
            bool goR = boxHit(r, mk3(b.z, b.w, c.x), mk3(c.y, c.z, c.w), tR) && tR < dist;
            if (goL) {
                if (goR) { stack.pushRef(sp, l.y); sp++; }
                cur = l.x;
                continue;
            }
            if (goR) { cur = l.y; continue; }
        }
        if (sp == 0) return false;
        cur = stack.ref(--sp);
    }
}

// ------------------------------------------------------------------------------------------------ traced tree
// Standard slab test on the padded boxes of the binned-SAH tree: t = p * inv + (-o * inv), explicit FMAs.
struct RayF { f3 inv, oi; };
RS_D RayF makeRayF(const RayT& r) {
    RayF f;
    // a zero component would give inf * 0 = NaN; 1e-20 keeps "origin inside the slab <=> always inside" intact
    float dx = fabsf(r.d.x) < 1e-20f ? copysignf(1e-20f, r.d.x) : r.d.x;
    float dy = fabsf(r.d.y) < 1e-20f ? copysignf(1e-20f, r.d.y) : r.d.y;
    float dz = fabsf(r.d.z) < 1e-20f ? copysignf(1e-20f, r.d.z) : r.d.z;
    f.inv = mk3(1.f / dx, 1.f / dy, 1.f / dz);
    f.oi = mk3(-r.o.x * f.inv.x, -r.o.y * f.inv.y, -r.o.z * f.inv.z);
    return f;
}
RS_D bool slabHit(const RayF& f, float lx, float ly, float lz, float hx, float hy, float hz, float tLimit, float& tEntry) {
    float x0 = __fmaf_rn(lx, f.inv.x, f.oi.x), x1 = __fmaf_rn(hx, f.inv.x, f.oi.x);
    float y0 = __fmaf_rn(ly, f.inv.y, f.oi.y), y1 = __fmaf_rn(hy, f.inv.y, f.oi.y);
    float z0 = __fmaf_rn(lz, f.inv.z, f.oi.z), z1 = __fmaf_rn(hz, f.inv.z, f.oi.z);
    tEntry = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));
    float tExit = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tLimit));
    return tEntry <= tExit;
}

// ---- what the reference's walk finds, without walking its tree ----
// The reference reaches a triangle only through nested boxes that end in the triangle's own AABB (1 triangle per leaf,
// bvh.cpp:25), each tested with the predicate of bvh.h:85-157 and pruned by "box distance < closest" (scene.h:260).
// Both the predicate and the box distance are monotone along the nesting (an enclosing box passes whenever the enclosed
// one does, and is entered no later), so the whole chain reduces to the leaf box: the reference finds triangle T iff
//   T is hit (intersections.h:17-53),  the predicate holds for AABB(T),  boxDistance(AABB(T)) < closest  and  d < closest,
// with candidates visited in the order of the ray's MTBVH ordering (bvh.cpp:156-193 -> DevScene::rank).  The distances d
// are computed with the reference's own arithmetic, so the order only matters between hits that can prune one another:
// hits whose distances differ by less than the Moller-Trumbore rounding error (boxDistance <= d up to that error).  Any
// CONSERVATIVE walk of any tree that offers every triangle near the ray to this criterion therefore finds the
// reference's hit; the walk keeps the best hit plus the hits inside a band behind it and replays the reference's visits of
// those in its order.  More than RS_MAX_TIES such hits leave the pixel undecided (fix-up kernel, reference-order walk).
// Two hits are "near" when their distances differ by no more than the sum of their Moller-Trumbore error bounds
// (triDistError) plus one ulp-scale term; RS_TIE_BAND only widens the traversal limit so that such hits BEHIND the best
// one are still visited.
#define RS_TIE_BAND 1e-4f
#define RS_MAX_TIES 3
#define RS_DONE 0x7fffffff
#define RS_WSTACK RS_PACKET_STACK

RS_D bool leafBox(const RayT& r, const Tri& t, float& tBox) {
    return boxHit(r, gmin(gmin(t.v0, t.v1), t.v2), gmax(gmax(t.v0, t.v1), t.v2), tBox);
}

// near-tie candidates of one ray: RS_MAX_TIES {triangle, distance, error bound} records per thread in shared memory, [entry][thread],
// followed by the ray's direction (three more rows): the walk needs it only at leaves, so it does not occupy registers
#define RS_TIE_ROWS (3 * RS_MAX_TIES + 3)
struct TieStore {
    float* base;       // this thread's column of [field: triangle, distance, error][entry][thread]; one pointer, 2 registers
    RS_D int& fi(int k) const { return *(int*)(base + k * RS_BLOCK); }
    RS_D float& d(int k) const { return base[(RS_MAX_TIES + k) * RS_BLOCK]; }
    RS_D float& e(int k) const { return base[(2 * RS_MAX_TIES + k) * RS_BLOCK]; }
    RS_D void setDir(f3 v) const { base[3 * RS_MAX_TIES * RS_BLOCK] = v.x; base[(3 * RS_MAX_TIES + 1) * RS_BLOCK] = v.y; base[(3 * RS_MAX_TIES + 2) * RS_BLOCK] = v.z; }
    RS_D f3 dir() const { return mk3(base[3 * RS_MAX_TIES * RS_BLOCK], base[(3 * RS_MAX_TIES + 1) * RS_BLOCK], base[(3 * RS_MAX_TIES + 2) * RS_BLOCK]); }
};
RS_D bool nearTie(float da, float ea, float db, float eb) { return fabsf(da - db) <= ea + eb + 1e-6f * fmaxf(da, db); }

// Running result of one closest-hit ray of a packet: 11 registers.  The slab test uses cinv / oi (explicit FMAs on the padded
// boxes of the traced tree); everything that decides which triangle is reported (triHit, leafBox) uses the ray's (o, d) with the
// reference's arithmetic: o is the packet's common origin (the camera position: warp-uniform, never in per-thread
// registers), d waits in the TieStore until a leaf is reached.
struct PRay {
    f3 cinv, oi;      // 1 / d with |d| kept >= 1e-20 (a zero component would give inf * 0 = NaN), -o / d
    float bestD;      // distance of the best hit so far (FLT_MAX: none)
    float bestErr;    // its error bound (triDistError)
    float limit;      // bestD widened by 2 * (RS_TIE_BAND * bestD + bestErr): nothing beyond matters.  < 0: no ray (lane outside the image)
    int bestFi;       // leaf-order index of the best hit's triangle, -1: none
    int nt;           // near-tie candidates in the TieStore; -1: more than RS_MAX_TIES (undecided)
};

RS_D PRay prayBegin(f3 o, f3 d, bool active, const TieStore& ts) {
    PRay p;
    ts.setDir(d);
    float dx = fabsf(d.x) < 1e-20f ? copysignf(1e-20f, d.x) : d.x;
    float dy = fabsf(d.y) < 1e-20f ? copysignf(1e-20f, d.y) : d.y;
    float dz = fabsf(d.z) < 1e-20f ? copysignf(1e-20f, d.z) : d.z;
    p.cinv = mk3(1.f / dx, 1.f / dy, 1.f / dz);
    p.oi = mk3(-o.x * p.cinv.x, -o.y * p.cinv.y, -o.z * p.cinv.z);
    p.bestD = FLT_MAX; p.bestErr = 0.f; p.limit = active ? FLT_MAX : -1.f;
    p.bestFi = -1; p.nt = 0;
    return p;
}

RS_D bool slabHitP(const PRay& p, float lx, float ly, float lz, float hx, float hy, float hz, float& tEntry) {
    float x0 = __fmaf_rn(lx, p.cinv.x, p.oi.x), x1 = __fmaf_rn(hx, p.cinv.x, p.oi.x);
    float y0 = __fmaf_rn(ly, p.cinv.y, p.oi.y), y1 = __fmaf_rn(hy, p.cinv.y, p.oi.y);
    float z0 = __fmaf_rn(lz, p.cinv.z, p.oi.z), z1 = __fmaf_rn(hz, p.cinv.z, p.oi.z);
    tEntry = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));
    float tExit = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), p.limit));
    return tEntry <= tExit;
}
// the same test when the caller already knows which plane of every slab the ray enters through (n*) and leaves through (f*):
// the same products, so the same tEntry / tExit bit for bit, without the six min / max that sort them
RS_D bool slabHitSorted(const PRay& p, float nx, float ny, float nz, float fx, float fy, float fz, float& tEntry) {
    float x0 = __fmaf_rn(nx, p.cinv.x, p.oi.x), x1 = __fmaf_rn(fx, p.cinv.x, p.oi.x);
    float y0 = __fmaf_rn(ny, p.cinv.y, p.oi.y), y1 = __fmaf_rn(fy, p.cinv.y, p.oi.y);
    float z0 = __fmaf_rn(nz, p.cinv.z, p.oi.z), z1 = __fmaf_rn(fz, p.cinv.z, p.oi.z);
    tEntry = fmaxf(fmaxf(x0, y0), fmaxf(z0, 0.f));
    float tExit = fminf(fminf(x1, y1), fminf(z1, p.limit));
    return tEntry <= tExit;
}

// intersections.h:17-53 on (o, d)
RS_D bool triHitOD(f3 o, f3 d, f3 v0, f3 v1, f3 v2, float& bx, float& by, float& dist) {
    RayT r;
    r.o = o; r.d = d;
    return triHit(r, v0, v1, v2, bx, by, dist);
}

// the part of a ray's running result that a newly accepted hit changes
struct BestState { float bestD, bestErr, limit; int bestFi, nt; };

// a triangle that is hit within the ray's limit (rare: a few times per ray): leaf-box criterion, then best / near-tie
// update.  Everything by value: a reference to the caller's PRay would pin it in local memory for the whole walk.
__device__ __noinline__ BestState prayAccept(f3 o, f3 dir, BestState p, const TieStore ts, f3 v0, f3 v1, f3 v2, int fi, float d) {
    RayT rt = makeRayT(o, dir);
    Tri t;
    t.v0 = v0; t.v1 = v1; t.v2 = v2;
    float tBox;
    if (!leafBox(rt, t, tBox)) return p;                             // the reference never sees this triangle
    const float err = triDistError(rt, v0, v1, v2);
    if (d < p.bestD) {
        // the stored candidates and the old best stay candidates when they are near the NEW best
        int n = 0;
        const bool oldNear = p.bestFi >= 0 && nearTie(d, err, p.bestD, p.bestErr);
        bool over = p.nt < 0 && oldNear;                             // candidates lost earlier lie behind the old best
        const int m = p.nt < 0 ? RS_MAX_TIES : p.nt;
        for (int k = 0; k < m; k++) {
            const float dk = ts.d(k), ek = ts.e(k);
            if (nearTie(d, err, dk, ek)) { ts.d(n) = dk; ts.e(n) = ek; ts.fi(n) = ts.fi(k); n++; }
        }
        if (oldNear) {
            if (n < RS_MAX_TIES) { ts.d(n) = p.bestD; ts.e(n) = p.bestErr; ts.fi(n) = p.bestFi; n++; }
            else over = true;
        }
        p.nt = over ? -1 : n;
        p.bestD = d; p.bestErr = err; p.bestFi = fi;
        p.limit = d + 2.f * (RS_TIE_BAND * d + err);
    } else if (nearTie(d, err, p.bestD, p.bestErr)) {
        if (p.nt >= 0 && p.nt < RS_MAX_TIES) { ts.d(p.nt) = d; ts.e(p.nt) = err; ts.fi(p.nt) = fi; p.nt++; }
        else p.nt = -1;
    }
    return p;
}

// one triangle of a visited leaf offered to a ray's running result
RS_D void prayOffer(PRay& p, f3 o, f3 dir, const TieStore& ts, const Tri& t, int fi) {
    float bx, by, d;
    if (!triHitOD(o, dir, t.v0, t.v1, t.v2, bx, by, d)) return;
    if (!(d <= p.limit)) return;
    BestState b;
    b.bestD = p.bestD; b.bestErr = p.bestErr; b.limit = p.limit; b.bestFi = p.bestFi; b.nt = p.nt;
    b = prayAccept(o, dir, b, ts, t.v0, t.v1, t.v2, fi, d);
    p.bestD = b.bestD; p.bestErr = b.bestErr; p.limit = b.limit; p.bestFi = b.bestFi; p.nt = b.nt;
}

// the reference's visits of the best hit and its near ties, in its order for this ray
__device__ __noinline__ Hit prayReplay(const DevScene& s, f3 o, f3 dir, int bestFi, int nt, const TieStore ts) {
    const RayT rt = makeRayT(o, dir);
    const int* rank = s.rank + (size_t)(2 * rt.dim + rt.lesser) * s.numTris;
    float cd[RS_MAX_TIES + 1], cbx[RS_MAX_TIES + 1], cby[RS_MAX_TIES + 1], ctb[RS_MAX_TIES + 1];
    int cprim[RS_MAX_TIES + 1], crk[RS_MAX_TIES + 1];
    Hit h;
    h.t = FLT_MAX; h.prim = -1; h.bx = 0.f; h.by = 0.f;
    const int n = nt + 1;
    for (int k = 0; k < n; k++) {
        const int fi = k == 0 ? bestFi : ts.fi(k - 1);
        Tri t = loadTriFast(s, fi);
        triHit(rt, t.v0, t.v1, t.v2, cbx[k], cby[k], cd[k]);
        leafBox(rt, t, ctb[k]);
        cprim[k] = t.prim;
        crk[k] = __ldg(rank + t.prim);
    }
    float closest = FLT_MAX;
    for (int k = 0; k < n; k++) {                          // n <= 4: selection by increasing rank
        int m = -1;
        for (int i = 0; i < n; i++)
            if (crk[i] >= 0 && (m < 0 || crk[i] < crk[m])) m = i;
        if (ctb[m] < closest && cd[m] < closest) {         // scene.h:260,267
            closest = cd[m];
            h.t = cd[m]; h.bx = cbx[m]; h.by = cby[m]; h.prim = cprim[m];
        }
        crk[m] = -1;
    }
    return h;
}

// false = undecided (more mutually near hits than RS_MAX_TIES + 1)
RS_D bool prayResolve(const DevScene& s, f3 o, f3 dir, const PRay& p, const TieStore& ts, Hit& h) {
    h.t = FLT_MAX; h.prim = -1; h.bx = 0.f; h.by = 0.f;
    if (p.bestFi < 0) return true;
    if (p.nt < 0) return false;
    if (p.nt > 0) {
        h = prayReplay(s, o, dir, p.bestFi, p.nt, ts);
        return true;
    }
    Tri t = loadTriFast(s, p.bestFi);
    triHitOD(o, dir, t.v0, t.v1, t.v2, h.bx, h.by, h.t);            // same arithmetic as during the walk: h.t == bestD
    h.prim = t.prim;
    return true;
}

// Closest hit of a WARP's rays over the traced tree as one packet: the 32 lanes hold the rays of an 8x4-pixel tile (one
// ray per lane, or two: the G-buffer's centre ray and the ReSTIR kernel's jittered ray of the same pixel, gbuffer.cu:11-23 /
// restir.cu:129), and the warp walks the tree ONCE with a single stack: a node is fetched once (same address in every
// lane), every lane slab-tests its own ray(s) against both children, a child is entered when ANY lane hits it, the
// nearer one (smallest entry distance over the warp) first, and every triangle of a visited leaf is offered to every
// lane's ray(s).  Each ray keeps its own result and its own limit, and what a ray finds does not depend on which
// conservative boxes were entered (see above), so the results are exactly those of 32 / 64 separate walks.  Primary
// rays of a tile are so coherent that the union of their walks is barely longer than the longest of them, while 32
// independent walks in lockstep cost about twice the longest (divergence between descending and leaf-testing lanes):
// scripts/travsim.cpp, config4 at 1080p: 106 node steps per warp against 191, with every lane busy instead of 10.6 of 32.
// The warp's stack lives in shared memory and is addressed through ONE 32-bit shared-space cursor (8-byte entries {node, entry
// distance bits}); every lane stores the same entry (same value, same address).
RS_D void wsPush(unsigned& cursor, int ref, unsigned key) {
#ifdef RS_HOST_EMU      /* tests/emu: a "shared-space address" is an offset from the emulation's anchor */
    int* p = (int*)emuSharedPtr(cursor);
    p[0] = ref; p[1] = (int)key;
#else
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(cursor), "r"(ref), "r"(key) : "memory");
#endif
    cursor += 8;
}
// next stacked node that is not beyond every lane's limit; the bottom entry is a sentinel {RS_DONE, distance 0} that always
// passes, so the cursor alone describes the stack (no base address to keep -- or to rebuild from %tid, S2R, in the walk)
RS_D int wsPop(unsigned& cursor, float wlimit) {
    for (;;) {
        cursor -= 8;
        int ref;
        unsigned key;
#ifdef RS_HOST_EMU
        const int* p = (const int*)emuSharedPtr(cursor);
        ref = p[0]; key = (unsigned)p[1];
#else
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(ref), "=r"(key) : "r"(cursor) : "memory");
#endif
        if (__uint_as_float(key) <= wlimit) return ref;
    }
}
// A lane's any-hit stack addressed through an OPAQUE 32-bit shared-space address: left to itself the compiler may rebuild the
// address from %tid inside the walk (S2R, a slow instruction: measured 13 % on k_shadow's per-lane loop).
RS_D unsigned stackAddr(const Stack& st) {
    unsigned a = (unsigned)__cvta_generic_to_shared(st.sRef);
    asm volatile("" : "+r"(a));
    return a;
}
RS_D void stackPush(unsigned addr, Stack& st, int sp, int ref) {
#ifdef RS_HOST_EMU
    if (sp < RS_SMEM_STACK) *(int*)emuSharedPtr(addr + (unsigned)sp * (RS_BLOCK * 4u)) = ref;
#else
    if (sp < RS_SMEM_STACK) asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr + (unsigned)sp * (RS_BLOCK * 4u)), "r"(ref) : "memory");
#endif
    else st.lRef[sp - RS_SMEM_STACK] = ref;
}
RS_D int stackRef(unsigned addr, const Stack& st, int sp) {
    if (sp < RS_SMEM_STACK) {
        int ref;
#ifdef RS_HOST_EMU
        ref = *(const int*)emuSharedPtr(addr + (unsigned)sp * (RS_BLOCK * 4u));
#else
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(ref) : "r"(addr + (unsigned)sp * (RS_BLOCK * 4u)) : "memory");
#endif
        return ref;
    }
    return st.lRef[sp - RS_SMEM_STACK];
}
template <bool TWO>
RS_D void packetWalk(const DevScene& s, f3 o, PRay& a, PRay& b, const TieStore& ta, const TieStore& tb, int2* wst) {
    const unsigned FULL = 0xffffffffu;
    unsigned wsp = (unsigned)__cvta_generic_to_shared(wst);
    wsPush(wsp, RS_DONE, 0u);
    float tA, tB;
    bool hit = slabHitP(a, s.fastRootMin[0], s.fastRootMin[1], s.fastRootMin[2], s.fastRootMax[0], s.fastRootMax[1], s.fastRootMax[2], tA);
    if (TWO) hit |= slabHitP(b, s.fastRootMin[0], s.fastRootMin[1], s.fastRootMin[2], s.fastRootMax[0], s.fastRootMax[1], s.fastRootMax[2], tB);
    int cur = __any_sync(FULL, hit) ? s.fastRoot : RS_DONE;
    // Rays of a tile nearly always share the signs of their direction components: then the entry / exit plane of every slab
    // is known per WARP (three uniform selects per box instead of six min / max per ray and box).  Lanes without a ray never
    // hit whatever the planes (limit < 0), so only the lanes with rays have to agree.
    int oct = (a.cinv.x < 0.f ? 1 : 0) | (a.cinv.y < 0.f ? 2 : 0) | (a.cinv.z < 0.f ? 4 : 0);
    const unsigned have = __ballot_sync(FULL, a.limit >= 0.f);
    const int oct0 = __shfl_sync(FULL, oct, have ? __ffs(have) - 1 : 0);
    bool same = a.limit < 0.f || oct == oct0;
    if (TWO) same = same && (b.limit < 0.f || ((b.cinv.x < 0.f ? 1 : 0) | (b.cinv.y < 0.f ? 2 : 0) | (b.cinv.z < 0.f ? 4 : 0)) == oct0);
    const bool sorted = __all_sync(FULL, same);
    const bool negX = (oct0 & 1) != 0, negY = (oct0 & 2) != 0, negZ = (oct0 & 4) != 0;
    float wlimit = FLT_MAX;                       // max of the lanes' limits (warp-uniform): culls popped entries
    for (;;) {
        while (cur >= 0 && cur != RS_DONE) {
            const float4* np = s.fastNodes + 4 * (size_t)cur;
            const F8 nA = ldg256(np), nB = ldg256(np + 2);
            const float4 n0 = nA.lo, n1 = nA.hi, n2 = nB.lo;
            const int2 l = make_int2(__float_as_int(nB.hi.x), __float_as_int(nB.hi.y));
            float tL, tR, t2;
            bool hL, hR;
            if (sorted) {
                const float lnx = negX ? n0.w : n0.x, lfx = negX ? n0.x : n0.w, lny = negY ? n1.x : n0.y, lfy = negY ? n0.y : n1.x;
                const float lnz = negZ ? n1.y : n0.z, lfz = negZ ? n0.z : n1.y;
                const float rnx = negX ? n2.y : n1.z, rfx = negX ? n1.z : n2.y, rny = negY ? n2.z : n1.w, rfy = negY ? n1.w : n2.z;
                const float rnz = negZ ? n2.w : n2.x, rfz = negZ ? n2.x : n2.w;
                hL = slabHitSorted(a, lnx, lny, lnz, lfx, lfy, lfz, tL);
                hR = slabHitSorted(a, rnx, rny, rnz, rfx, rfy, rfz, tR);
                if (!hL) tL = FLT_MAX;
                if (!hR) tR = FLT_MAX;
                if (TWO) {
                    if (slabHitSorted(b, lnx, lny, lnz, lfx, lfy, lfz, t2)) { hL = true; tL = fminf(tL, t2); }
                    if (slabHitSorted(b, rnx, rny, rnz, rfx, rfy, rfz, t2)) { hR = true; tR = fminf(tR, t2); }
                }
            } else {
                hL = slabHitP(a, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, tL);
                hR = slabHitP(a, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, tR);
                if (!hL) tL = FLT_MAX;
                if (!hR) tR = FLT_MAX;
                if (TWO) {
                    if (slabHitP(b, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, t2)) { hL = true; tL = fminf(tL, t2); }
                    if (slabHitP(b, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, t2)) { hR = true; tR = fminf(tR, t2); }
                }
            }
            const bool anyL = __any_sync(FULL, hL), anyR = __any_sync(FULL, hR);
            if (anyL && anyR) {
                // entry distances are >= 0, so their bit patterns order like the floats
                const unsigned kL = __reduce_min_sync(FULL, __float_as_uint(tL)), kR = __reduce_min_sync(FULL, __float_as_uint(tR));
                const bool leftNear = kL <= kR;
                wsPush(wsp, leftNear ? l.y : l.x, leftNear ? kR : kL);
                cur = leftNear ? l.x : l.y;
            } else if (anyL) cur = l.x;
            else if (anyR) cur = l.y;
            else cur = wsPop(wsp, wlimit);
        }
        if (cur == RS_DONE) return;
        {
            const int first = cur & 0x07ffffff, count = ((cur >> 27) & 7) + 1;
            const f3 da = ta.dir(), db = TWO ? tb.dir() : da;
            for (int i = 0; i < count; i++) {
                const Tri t = loadTriFast(s, first + i);
                prayOffer(a, o, da, ta, t, first + i);
                if (TWO) prayOffer(b, o, db, tb, t, first + i);
            }
        }
        const float lim = TWO ? fmaxf(a.limit, b.limit) : a.limit;
        wlimit = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(fmaxf(lim, 0.f))));
        cur = wsPop(wsp, wlimit);
    }
}

// shared memory of the packet walk for a block of RS_BLOCK threads: near-tie stores of NR rays per thread + one stack per warp
#define RS_DECLARE_PACKET(name, NR)                                                 \
    __shared__ float name##_ties[NR][RS_TIE_ROWS][RS_BLOCK];                            \
    __shared__ int2 name##_ws[RS_BLOCK / 32][RS_WSTACK + 1];                            \
    TieStore name##_ta, name##_tb;                                                  \
    name##_ta.base = &name##_ties[0][0][threadIdx.x];                               \
    name##_tb.base = &name##_ties[NR - 1][0][threadIdx.x];                          \
    int2* name##_wst = name##_ws[threadIdx.x >> 5]

// any hit: order does not matter, the criterion above is exact per triangle
RS_D int traceOccludedFast(const DevScene& s, const RayT& r, float dist, Stack& stack) {
    const RayF f = makeRayF(r);
    float t0;
    if (!slabHit(f, s.fastRootMin[0], s.fastRootMin[1], s.fastRootMin[2], s.fastRootMax[0], s.fastRootMax[1], s.fastRootMax[2], dist, t0)) return 0;
    int sp = 0;
    int cur = s.fastRoot;
    for (;;) {
        while (cur >= 0) {
            const float4* np = s.fastNodes + 4 * (size_t)cur;
            const F8 nA = ldg256(np), nB = ldg256(np + 2);
            const float4 a = nA.lo, b = nA.hi, c = nB.lo;
            const int2 l = make_int2(__float_as_int(nB.hi.x), __float_as_int(nB.hi.y));
            float tL, tR;
            bool hL = slabHit(f, a.x, a.y, a.z, a.w, b.x, b.y, dist, tL);
            bool hR = slabHit(f, b.z, b.w, c.x, c.y, c.z, c.w, dist, tR);
            if (hL && hR) {
                bool leftNear = tL <= tR;
                stack.pushRef(sp, leftNear ? l.y : l.x); sp++;
                cur = leftNear ? l.x : l.y;
            } else if (hL) cur = l.x;
            else if (hR) cur = l.y;
            else {
                if (sp == 0) return 0;
                cur = stack.ref(--sp);
            }
        }
        int first = cur & 0x07ffffff, count = ((cur >> 27) & 7) + 1;
        for (int i = 0; i < count; i++) {
            Tri t = loadTriFast(s, first + i);
            float bx, by, d, tb;
            if (triHit(r, t.v0, t.v1, t.v2, bx, by, d) && d < dist && leafBox(r, t, tb) && tb < dist) return 1;
        }
        if (sp == 0) return 0;
        cur = stack.ref(--sp);
    }
}

// DevScene::testOcclusion (scene.h:286-316): 1 occluded, 0 free, -1 undecided (never with EXACT)
template <bool EXACT>
RS_D int traceOccluded(const DevScene& s, f3 x, f3 y, Stack& stack) {
    const float Eps = 1e-4f;
    f3 dir = y - x;
    float dist = length(dir);
    if (!(dist > 0.f)) return 0;              // x == y (or NaN): dir is NaN and every reference box test fails
    dir = dir / dist;
    RayT r = makeRayT(x + dir * 1e-5f, dir);                                   // makeOffsetedRay, intersections.h:12
    dist -= Eps * 2.f;
    if (EXACT) return traceOccludedExact(s, r, dist, stack) ? 1 : 0;
    return traceOccludedFast(s, r, dist, stack);
}

// ------------------------------------------------------------------------------------------------ camera
// cameraRay / cameraOrigin: camera_dev.h

// Camera::getRasterCoord (sceneStructs.h:23-46); float->int is cvt.rzi (saturating, NaN -> 0) as in the reference's kernel
RS_D void rasterCoord(const CamDev& c, f3 pos, int& ox, int& oy) {
    f3 dir = normalize(pos - mk3(c.position[0], c.position[1], c.position[2]));
    float dd = 1.f / dot(dir, mk3(c.view[0], c.view[1], c.view[2]));
    f3 q = dir * dd;
    const float* m = c.rotInv;
    f3 p = mk3(m[0] * q.x + m[3] * q.y + m[6] * q.z, m[1] * q.x + m[4] * q.y + m[7] * q.z, m[2] * q.x + m[5] * q.y + m[8] * q.z);
    p = p / mk3(c.aspect * c.tanFovY, 1.f * c.tanFovY, 1.f);
    float nx = -p.x, ny = -p.y;
    nx = nx * .5f + .5f; ny = ny * .5f + .5f;
    ox = __float2int_rz(c.resX * nx);
    oy = __float2int_rz(c.resY * ny);
}

// ------------------------------------------------------------------------------------------------ shading
RS_D float schlickG(float c, float alpha) { float a = alpha * .5f; return c / (c * (1.f - a) + a); }   // material.h:62
RS_D float GTR2Distrib(float c, float alpha) {                                                          // material.h:71
    if (c < 1e-6f) return 0.f;
    float aa = alpha * alpha;
    float denom = c * c * (aa - 1.f) + 1.f;
    denom = denom * denom * RS_PI;
    return aa / denom;
}
// Material::BSDF (material.h:218-228; :122 lambertian, :171-186 metallic workflow, :137 dielectric)
// `diffuse` = baseColor * 1.f / Pi ("baseColor * PiInv", material.h:123,185): three IEEE divisions that do not depend on
// wi, hoisted out of the 32-candidate loop by the callers
RS_D f3 diffuseTerm(f3 baseColor) { return baseColor * 1.f / RS_PI; }
RS_D f3 materialBSDF(int type, float metallic, float roughness, f3 baseColor, f3 diffuse, f3 n, f3 wo, f3 wi) {
    if (type == 0) return diffuse;
    if (type == 1) {
        float alpha = roughness * roughness;
        f3 h = normalize(wo + wi);
        float cosO = dot(n, wo), cosI = dot(n, wi);
        if (cosI * cosO < 1e-7f) return mk3(0.f);
        f3 f0 = mix(mk3(.08f), baseColor, metallic);
        f3 f = mix(f0, mk3(1.f), pow5(1.f - dot(h, wo)));
        float g = schlickG(fabsf(cosO), alpha) * schlickG(fabsf(cosI), alpha);
        float d = GTR2Distrib(dot(n, h), alpha);
        return mix(diffuse * (1.f - metallic), mk3(g * d / (4.f * cosI * cosO)), f);
    }
    return mk3(0.f);
}


// ------------------------------------------------------------------------------------------------ textures
// image.h:41-74 linearSample on the float4-texel copy of texture `id`
RS_D float fractf_(float x) { return x - floorf(x); }                                                   // func_common.inl:332
RS_D f3 texLinear(const DevScene& s, int id, float u, float v) {
    int4 info = __ldg(s.texInfo + id);
    const int width = info.x, height = info.y;
    const float4* data = s.texData + info.z;
    u = fractf_(u); v = fractf_(v);
    float fx = u * ((float)width - 1.17549435e-38f) + .5f;
    float fy = v * ((float)height - 1.17549435e-38f) + .5f;
    int ix = __float2int_rz(fractf_(fx) > .5f ? fx : fx - 1);
    if (ix < 0) ix += width;
    int iy = __float2int_rz(fractf_(fy) > .5f ? fy : fy - 1);
    if (iy < 0) iy += height;
    int ux = ix + 1;
    if (ux >= width) ux -= width;
    int uy = iy + 1;
    if (uy >= height) uy -= height;
    float lx = fractf_(fx + .5f);
    float ly = fractf_(fy + .5f);
    float4 a = __ldg(data + iy * width + ix), b = __ldg(data + iy * width + ux);
    float4 c = __ldg(data + uy * width + ix), d = __ldg(data + uy * width + ux);
    f3 c1 = mix(mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), lx);
    f3 c2 = mix(mk3(c.x, c.y, c.z), mk3(d.x, d.y, d.z), lx);
    return mix(c1, c2, ly);
}
// scene.h:68-76; thrust::default_random_engine seeded with the texel hash, two uniform draws
RS_D f3 proceduralTex(float u, float v) {
    Rng rng;
    uint32_t seed = (uint32_t)(__float2int_rz(u * 1024) * 1024 + __float2int_rz(v * 1024));
    rng.x = seed % 2147483647u;
    if (rng.x == 0) rng.x = 1;
    float rx = rng.next();
    float ry = rng.next();
    const float PiTwo = 6.2831853071795864769252867665590057683943f;
    float f = (sinf(u * 10.f * PiTwo + rx * PiTwo) + 1.f) * .5f;
    float g = (sinf(v * 10.f * PiTwo + ry * PiTwo) + 1.f) * .5f;
    return mk3(f * g);
}
RS_D f3 localToWorld(f3 n, f3 v) {                                                                      // mathUtil.h:146-155
    f3 t = (fabsf(n.y) > 0.9999f) ? mk3(0.f, 0.f, 1.f) : mk3(0.f, 1.f, 0.f);
    f3 b = normalize(cross(n, t));
    t = cross(b, n);
    f3 r = mk3(t.x * v.x + b.x * v.y + n.x * v.z, t.y * v.x + b.y * v.y + n.y * v.z, t.z * v.x + b.z * v.y + n.z * v.z);
    return normalize(r);
}
// gbuffer.cu:60-61 / restir.cu:135: envMap->linearSample(Math::toPlane(dir)), mathUtil.h:139-144
RS_D f3 envLookup(const DevScene& s, f3 d) {
    float u = fractf_(atan2f(d.z, d.x) * 1.f / RS_PI * .5f + 1.f);
    float v = atan2f(sqrtf(d.x * d.x + d.z * d.z), d.y) * 1.f / RS_PI;
    return texLinear(s, s.envTex, u, v);
}
// DevScene::getTexturedMaterialAndSurface (scene.h:78-99) at the hit (prim, bary): material with its maps applied;
// a normal map replaces nrm
struct Surf { int type; f3 baseColor; float metallic, roughness; };
RS_D Surf texturedMaterial(const DevScene& s, int matId, int prim, float bx, float by, f3& nrm) {
    const RstrMaterial* m = s.materials + matId;
    Surf r;
    r.type = __ldg(&m->type);
    r.baseColor = mk3(__ldg(&m->baseColor[0]), __ldg(&m->baseColor[1]), __ldg(&m->baseColor[2]));
    r.metallic = __ldg(&m->metallic);
    r.roughness = __ldg(&m->roughness);
    if (!s.anyMaps) return r;
    const int bcMap = __ldg(&m->baseColorMapId), mMap = __ldg(&m->metallicMapId), rMap = __ldg(&m->roughnessMapId), nMap = __ldg(&m->normalMapId);
    if (bcMap == -1 && mMap <= -1 && rMap <= -1 && nMap == -1) return r;
    const float4* tp = s.triUV + 2 * (size_t)prim;
    float4 t01 = __ldg(tp), t2 = __ldg(tp + 1);
    float bz = 1.f - bx - by;
    float u = t01.z * bx + t2.x * by + t01.x * bz;                                                      // scene.h:150
    float v = t01.w * bx + t2.y * by + t01.y * bz;
    if (bcMap != -1) r.baseColor = bcMap == -2 ? proceduralTex(u, v) : texLinear(s, bcMap, u, v);
    if (mMap > -1) r.metallic = texLinear(s, mMap, u, v).x;
    if (rMap > -1) r.roughness = texLinear(s, rMap, u, v).x;
    if (nMap != -1) {
        f3 mapped = texLinear(s, nMap, u, v);
        f3 localNorm = normalize(mk3(mapped.x * 1.f - 0.5f, mapped.y * 1.f - 0.5f, mapped.z * 1.f - 0.5f));
        nrm = localToWorld(nrm, localNorm);
    }
    return r;
}

struct Resv {       // register-resident reservoir (restir.h:29-117); Li is implied by lightId
    f3 wi; float dist; float w; int M; int lightId;
};
RS_D Resv emptyResv() { Resv r; r.wi = mk3(0.f); r.dist = 0.f; r.w = 0.f; r.M = 0; r.lightId = -1; return r; }
RS_D bool resvInvalid(const Resv& r) { return isNanOrInf(r.w) || r.w < 0.f; }                           // :51
RS_D void resvCheck(Resv& r) { if (resvInvalid(r)) { r.w = 0.f; r.M = 0; } }                            // :55
RS_D void resvMerge(Resv& a, const Resv& b, float rnd) {                                                 // :61
    a.w += b.w; a.M += b.M;
    if (rnd * a.w < b.w) { a.wi = b.wi; a.dist = b.dist; a.lightId = b.lightId; }
}
RS_D void resvClamp(Resv& r, int val) { if (r.M > val) { r.w *= (float)val / r.M; r.M = val; } }        // :88
RS_D Resv loadResv(const ResvD* p) {
    const float4* q = (const float4*)p;
    float4 a = q[0], b = q[1];
    Resv r; r.wi = mk3(a.x, a.y, a.z); r.dist = a.w; r.w = b.x; r.M = __float_as_int(b.y); r.lightId = __float_as_int(b.z);
    return r;
}
RS_D void storeResv(ResvD* p, const Resv& r) {
    float4* q = (float4*)p;
    q[0] = make_float4(r.wi.x, r.wi.y, r.wi.z, r.dist);
    q[1] = make_float4(r.w, __int_as_float(r.M), __int_as_float(r.lightId), 0.f);
}
RS_D f3 lightLe(const DevScene& s, int lightId) {
    if (lightId < 0) return mk3(0.f);
    if (s.envTex >= 0 && lightId >= s.numLights - 1) {          // environment map: lightId = (L-1) + texel
        float4 e = __ldg(s.texData + __ldg(s.texInfo + s.envTex).z + (lightId - (s.numLights - 1)));
        return mk3(e.x, e.y, e.z);
    }
    float4 d = __ldg(s.lights + 4 * (size_t)lightId + 3);       // LightRec words 12..15 = {Le.x, Le.y, Le.z, pdfArea}
    return mk3(d.x, d.y, d.z);
}

// pixel owned by this thread: 128-thread blocks = 4 warps of 8x4 pixels, block tile 16x8
RS_D bool pixelOf(const FrameDev& f, int& x, int& y) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    y = f.rowLo + blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    return x < f.W && y < f.rowHi;
}
// pixels whose outcome depends on the reference's visiting order are deferred to the fix-up kernel
RS_D void enqueuePixel(const FrameDev& f, int x, int y) { f.queue[atomicAdd(f.queueCount, 1u)] = y * f.W + x; }
RS_D size_t planeIndex(const FrameDev& f, int x, int y) { return (size_t)(y - f.bufRow0) * f.W + x; }
// cost profile for the strip cuts: SM cycles every block (16x8 pixels) held its SM slot, summed per group of 8 rows
RS_D void accountBlock(const FrameDev& f, const unsigned int* t0) {
    __syncthreads();
    if (threadIdx.x == 0) {
        int y = f.rowLo + blockIdx.y * 8;
        if (y < f.H) atomicAdd(f.rowCost + (y >> 3), (unsigned long long)((unsigned int)clock64() - *t0));
    }
}
RS_D bool rowResident(const FrameDev& f, int y) { return y >= f.bufRow0 && y < f.bufRow0 + f.bufRows; }

// restir.cu:216-230: final shade of a reservoir and accumulation into the radiance image
RS_D void writeRadiance(const FrameDev& f, size_t li, f3 direct, int iter) {
    float4 am = f.albedoMotion[li];
    direct = direct * mk3(am.x, am.y, am.z);                                     // :229
    float* out = f.radiance + 3 * li;
    f3 prev = mk3(out[0], out[1], out[2]);
    f3 v = (prev * (float)iter + direct) / (float)(iter + 1);                    // :230
    out[0] = v.x; out[1] = v.y; out[2] = v.z;
}
RS_D f3 shadeReservoir(const DevScene& s, const Resv& r, int type, float metallic, float roughness, f3 n, f3 wo) {
    f3 direct = mk3(0.f);
    if (!resvInvalid(r)) {
        f3 LiBSDF = lightLe(s, r.lightId) * materialBSDF(type, metallic, roughness, mk3(1.f), diffuseTerm(mk3(1.f)), n, wo, r.wi);
        direct = LiBSDF / luminance(LiBSDF) * r.w / (float)r.M;
    }
    if (hasNanOrInf(direct)) direct = mk3(0.f);
    return direct;
}

// ------------------------------------------------------------------------------------------------ kernels
// gbuffer.cu:3-73 for one pixel; false = undecided, nothing written
RS_D void gbufferFinish(const DevScene& s, const FrameDev& f, const CamDev& lastCam, int x, int y, f3 o, f3 d, const Hit& h);

// with the reference-order walk of the reference tree (validation mode, fix-up kernels)
RS_D void gbufferPixelExact(const DevScene& s, const FrameDev& f, const CamDev& cam, const CamDev& lastCam, int x, int y, Stack& stack) {
    f3 o, d;
    cameraRay(cam, x, y, .5f, .5f, o, d);
    RayT r = makeRayT(o, d);
    Hit h;
    traceClosestExact(s, r, h, stack);
    gbufferFinish(s, f, lastCam, x, y, o, d, h);
}

// gbuffer.cu:28-72: what is stored for a pixel once its hit is known
RS_D void gbufferFinish(const DevScene& s, const FrameDev& f, const CamDev& lastCam, int x, int y, f3 o, f3 d, const Hit& h) {
    size_t li = planeIndex(f, x, y);
    if (h.prim >= 0) {
        Tri t = loadTri(s, h.prim);
        const float4* np = s.triNorm + 3 * (size_t)h.prim;
        float4 na4 = __ldg(np), nb4 = __ldg(np + 1), nc4 = __ldg(np + 2);
        f3 na = mk3(na4.x, na4.y, na4.z), nb = mk3(na4.w, nb4.x, nb4.y), nc = mk3(nb4.z, nb4.w, nc4.x);
        float bz = 1.f - h.bx - h.by;
        f3 pos = t.v1 * h.bx + t.v2 * h.by + t.v0 * bz;                              // scene.h:148
        f3 nrm = normalize(nb * h.bx + nc * h.by + na * bz);                         // scene.h:149
        Surf m = texturedMaterial(s, t.matId, h.prim, h.bx, h.by, nrm);              // gbuffer.cu:37
        int matId = m.type == 4 ? -2 : t.matId;                                      // gbuffer.cu:29-31
        f3 albedo = m.baseColor;
        float depth = length(o - pos);                                               // gbuffer.cu:44
        int lx, ly;
        rasterCoord(lastCam, pos, lx, ly);                                           // gbuffer.cu:49-55
        int motion = (lx >= 0 && lx < f.W && ly >= 0 && ly < f.H) ? ly * f.W + lx : -1;
        if (motion >= 0) {                       // per-frame bound on the temporal halo: max |row' - row|
            unsigned int dy = (unsigned int)abs(ly - y);
            if (dy > RS_PEEK(f.motionRows)) atomicMax(f.motionRows, dy);
        }
        f.geom[0][li] = make_float4(nrm.x, nrm.y, nrm.z, depth);
        f.matId[0][li] = matId;
        f.albedoMotion[li] = make_float4(albedo.x, albedo.y, albedo.z, __int_as_float(motion));
    } else {
        f.geom[0][li] = make_float4(0.f, 0.f, 0.f, 1.f);                             // gbuffer.cu:57-72
        f.matId[0][li] = -1;
        f3 albedo = s.envTex >= 0 ? envLookup(s, d) : mk3(0.f);                      // gbuffer.cu:58-63
        f.albedoMotion[li] = make_float4(albedo.x, albedo.y, albedo.z, __int_as_float(0));
    }
}

// traced tree: the warp's 32 centre rays walk as one packet
__global__ void __launch_bounds__(RS_BLOCK, RS_MINB_GBUF) k_gbuffer(const __grid_constant__ DevScene s, const __grid_constant__ FrameDev f,
                                                      const __grid_constant__ CamDev cam, const __grid_constant__ CamDev lastCam) {
    RS_DECLARE_PACKET(pk, 1);
    __shared__ unsigned int t0;
    if (f.rowCost && threadIdx.x == 0) t0 = (unsigned int)clock64();
    int x, y;
    const bool active = pixelOf(f, x, y);
    f3 o = mk3(0.f), d = mk3(0.f, 0.f, 1.f);
    if (active) cameraRay(cam, x, y, .5f, .5f, o, d);
    const f3 oc = cameraOrigin(cam);
    PRay a = prayBegin(oc, d, active, pk_ta);
    packetWalk<false>(s, oc, a, a, pk_ta, pk_ta, pk_wst);
    if (active) {
        Hit h;
        if (prayResolve(s, oc, d, a, pk_ta, h)) gbufferFinish(s, f, lastCam, x, y, o, d, h);
        else enqueuePixel(f, x, y);
    }
    if (f.rowCost) accountBlock(f, &t0);
}
// reference-order walk for every pixel (RS_TRAVERSAL_EXACT)
__global__ void __launch_bounds__(RS_BLOCK, RS_MINB_GBUF) k_gbuffer_exact(const __grid_constant__ DevScene s, const __grid_constant__ FrameDev f,
                                                            const __grid_constant__ CamDev cam, const __grid_constant__ CamDev lastCam) {
    RS_DECLARE_STACK(stack);
    __shared__ unsigned int t0;
    if (f.rowCost && threadIdx.x == 0) t0 = (unsigned int)clock64();
    int x, y;
    if (pixelOf(f, x, y)) gbufferPixelExact(s, f, cam, lastCam, x, y, stack);
    if (f.rowCost) accountBlock(f, &t0);
}
// recomputes the queued pixels with the reference-order walk
__global__ void __launch_bounds__(RS_BLOCK) k_gbuffer_fix(const __grid_constant__ DevScene s, const __grid_constant__ FrameDev f,
                                                          const __grid_constant__ CamDev cam, const __grid_constant__ CamDev lastCam) {
    RS_DECLARE_STACK(stack);
    const unsigned n = *f.queueCount;
    // entry i -> lane i / numWarps of warp i % numWarps: the first numWarps entries each get a warp of their own
    // (reference-order walks diverge completely, sharing a warp would serialise them)
    const unsigned numWarps = gridDim.x * (RS_BLOCK / 32), gw = blockIdx.x * (RS_BLOCK / 32) + (threadIdx.x >> 5);
    for (unsigned i = (threadIdx.x & 31) * numWarps + gw; i < n; i += 32 * numWarps) {
        int idx = f.queue[i];
        gbufferPixelExact(s, f, cam, lastCam, idx % f.W, idx / f.W, stack);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(s.fallbackRays + 0, n);
}

// scene.h:364-375: texel of the environment map drawn from envMapSampler -> radiance and direction of its centre
RS_D int envSample(const DevScene& s, float r1, float r2, f3& radiance, f3& wi) {
    int len = s.envLen;
    int pass = min(__float2int_rz((float)len * r1), len - 1);                       // sampler.h:204
    float2 e = __ldg(s.envAlias + pass);
    int pix = (r2 < e.x) ? pass : __float_as_int(e.y);
    int4 info = __ldg(s.texInfo + s.envTex);
    float4 c = __ldg(s.texData + info.z + pix), w = __ldg(s.envDir + pix);
    radiance = mk3(c.x, c.y, c.z);
    wi = mk3(w.x, w.y, w.z);
    return pix;
}
RS_D float envPdf(const DevScene& s, f3 radiance) {                                  // scene.h:373-374 (PiInv = 1.f / Pi)
    int4 info = __ldg(s.texInfo + s.envTex);
    return luminance(radiance) * s.sumLightPowerInv * (float)info.x * (float)info.y * 1.f / RS_PI * 1.f / RS_PI * .5f;
}
// scene.h:394-425 on the packed light record; returns the pdf (<= 0: invalid), fills wi / dist / lightId
RS_D float sampleLight(const DevScene& s, f3 pos, float r0, float r1, float r2, float r3, f3& Li, f3& wi, float& dist, int& lightId) {
    int len = s.numLights;
    int pass = min(__float2int_rz((float)len * r0), len - 1);                       // sampler.h:204
    float2 e = __ldg(s.alias + pass);
    lightId = (r1 < e.x) ? pass : __float_as_int(e.y);
    if (s.envTex >= 0 && lightId == len - 1) {                                       // scene.h:400-403
        dist = 1e10f;
        int pix = envSample(s, r2, r3, Li, wi);
        lightId = len - 1 + pix;
        return envPdf(s, Li);
    }
    const float4* lp = s.lights + 4 * (size_t)lightId;
    const F8 lA = ldg256(lp), lB = ldg256(lp + 2);
    const float4 a = lA.lo, b = lA.hi, c = lB.lo, d4 = lB.hi;
    f3 v0 = mk3(a.x, a.y, a.z), v1 = mk3(a.w, b.x, b.y), v2 = mk3(b.z, b.w, c.x);
    f3 n = mk3(c.y, c.z, c.w);
    float sr = sqrtf(r3);                                                            // mathUtil.h:94-100 (ru = r.z, rv = r.w)
    float u = 1.f - sr;
    float v = r2 * sr;
    f3 sampled = v1 * u + v2 * v + v0 * (1.f - u - v);
    f3 pts = sampled - pos;
    if (dot(n, pts) > -1e-6f) return -1.f;                                           // scene.h:415
    Li = mk3(d4.x, d4.y, d4.z);
    float len2 = dot(pts, pts);
    float sq = sqrtf(len2);
    wi = pts * (1.f / sq);                                                           // normalize(posToSampled)
    dist = sq;                                                                       // length(posToSampled)
    // pdfAreaToSolidAngle(pdfArea, pos, sampled, n): yx = pos - sampled = -pts exactly, so dot(yx,yx) = len2 and
    // normalize(yx) = -wi bit for bit; |dot(n, -wi)| = |dot(n, wi)|
    return d4.w * len2 / fabsf(dot(n, wi));
}

// restir.cu:20-45
RS_D Resv findTemporal(const FrameDev& f, size_t li, int idx) {
    int primId = f.matId[0][li];
    int lastIdx = __float_as_int(f.albedoMotion[li].w);
    if (lastIdx < 0 || primId <= -1) return emptyResv();
    int ly = lastIdx / f.W;
    if (!rowResident(f, ly)) { atomicAdd(f.haloMiss, 1u); return emptyResv(); }
    size_t lli = (size_t)lastIdx - (size_t)f.bufRow0 * f.W;
    if (f.matId[1][lli] != primId) return emptyResv();
    float4 g = f.geom[0][li], lg = f.geom[1][lli];
    if (absDot(mk3(g.x, g.y, g.z), mk3(lg.x, lg.y, lg.z)) < .9f || fabsf(lg.w - g.w) > g.w * .1f) return emptyResv();
    return loadResv(f.resvIn + lli);
}

// restir.cu:119-192 (+ :211-230 when spatial reuse is off) for one pixel; false = undecided, nothing written
template <bool EXACT, bool SPATIAL>
RS_D bool restirAAfterHit(const DevScene& s, const FrameDev& f, const RstrParams& prm, int iter, int first,
                          int x, int y, Stack& stack, Rng rng, f3 d, const Hit& h);

// the jittered primary ray of a pixel (restir.cu:126-129): seeds the pixel's RNG stream and draws sample4D
RS_D void jitteredRay(const FrameDev& f, const CamDev& cam, int looper, int x, int y, Rng& rng, f3& o, f3& d) {
    rng.seed(looper, y * f.W + x);
    float r0 = rng.next(), r1 = rng.next();
    rng.next(); rng.next();                                                          // r.z, r.w of sample4D are drawn and unused (restir.cu:129)
    cameraRay(cam, x, y, r0, r1, o, d);
}

// with the reference-order walk of the reference tree (validation mode, fix-up kernels)
template <bool SPATIAL>
RS_D void restirAPixelExact(const DevScene& s, const FrameDev& f, const CamDev& cam, const RstrParams& prm, int looper, int iter, int first,
                            int x, int y, Stack& stack) {
    Rng rng;
    f3 o, d;
    jitteredRay(f, cam, looper, x, y, rng, o, d);
    RayT ray = makeRayT(o, d);
    Hit h;
    traceClosestExact(s, ray, h, stack);
    restirAAfterHit<true, SPATIAL>(s, f, prm, iter, first, x, y, stack, rng, d, h);
}

// ---- the stages of restir.cu:133-192 (+ :211-230) once the jittered primary ray's hit is known.  The fused kernel runs
// them back to back in one thread (restirAAfterHit); the staged pipeline runs each in a kernel of its own over the
// compacted list of shaded pixels.
struct ShadePoint {
    f3 pos, nrm, wo;
    int matId, type;
    float metallic, roughness;
};
// restir.cu:133-153: 0 miss, 1 emitter, 2 shaded (ShadePoint filled, normal flipped toward wo)
RS_D int shadePointOf(const DevScene& s, const Hit& h, f3 d, ShadePoint& sp) {
    if (h.prim < 0) return 0;
    Tri t = loadTri(s, h.prim);
    const float4* np = s.triNorm + 3 * (size_t)h.prim;
    float4 na4 = __ldg(np), nb4 = __ldg(np + 1), nc4 = __ldg(np + 2);
    f3 na = mk3(na4.x, na4.y, na4.z), nb = mk3(na4.w, nb4.x, nb4.y), nc = mk3(nb4.z, nb4.w, nc4.x);
    float bz = 1.f - h.bx - h.by;
    sp.pos = t.v1 * h.bx + t.v2 * h.by + t.v0 * bz;
    sp.nrm = normalize(nb * h.bx + nc * h.by + na * bz);
    sp.matId = t.matId;
    Surf m = texturedMaterial(s, sp.matId, h.prim, h.bx, h.by, sp.nrm);              // restir.cu:140
    sp.type = m.type;
    sp.metallic = m.metallic;
    sp.roughness = m.roughness;
    if (sp.type == 4) return 1;
    sp.wo = -d;
    if (sp.type != 2 && dot(sp.nrm, sp.wo) < 0.f) sp.nrm = -sp.nrm;                  // restir.cu:150-153
    return 2;
}
// restir.cu:133-146 -> WriteRadiance for a pixel whose jittered ray missed (status 0) or hit an emitter (status 1)
template <bool SPATIAL>
RS_D void unshadedPixel(const DevScene& s, const FrameDev& f, size_t li, int status, f3 d, int iter) {
    f3 direct = status == 1 ? mk3(1.f) : (s.envTex >= 0 ? envLookup(s, d) : mk3(0.f));
    writeRadiance(f, li, direct, iter);
    if (SPATIAL) {
        float4* q = (float4*)(f.hit + li);
        q[0] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
        q[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
// restir.cu:155-169: the RIS reservoir over numCandidates light samples
RS_D Resv candidateLoop(const DevScene& s, const RstrParams& prm, const ShadePoint& sp, Rng& rng) {
    Resv R = emptyResv();
    const f3 diffuse = diffuseTerm(mk3(1.f));                                        // material.baseColor = 1 (restir.cu:141)
    const int nc = prm.numCandidates;
    for (int i = 0; i < nc; i++) {
        float c0 = rng.next(), c1 = rng.next(), c2 = rng.next(), c3 = rng.next();
        f3 Li = mk3(0.f), wi = mk3(0.f);
        float dist = 0.f;
        int lid = -1;
        float p = s.numLights > 0 ? sampleLight(s, sp.pos, c0, c1, c2, c3, Li, wi, dist, lid) : -1.f;
        f3 g = Li * materialBSDF(sp.type, sp.metallic, sp.roughness, mk3(1.f), diffuse, sp.nrm, sp.wo, wi) * satDot(sp.nrm, wi);
        float weight = luminance(g / p);
        if (isNanOrInf(weight) || p <= 0.f) weight = 0.f;
        float rnd = rng.next();
        R.w += weight; R.M++;                                                        // Reservoir::update, restir.h:38
        if (rnd * R.w < weight) { R.wi = wi; R.dist = dist; R.lightId = lid; }
    }
    return R;
}
// restir.cu:180-192, 211-230: temporal merge, history / publication stores, and the final shade when spatial reuse is off
template <bool SPATIAL>
RS_D void temporalAndStore(const DevScene& s, const FrameDev& f, const RstrParams& prm, int iter, int first, size_t li, int index,
                           f3 nrm, f3 wo, int matId, int type, float metallic, float roughness, Resv R, Rng rng) {
    if (!first && (prm.reuse & 1)) {                                                 // restir.cu:180-185
        Resv T = findTemporal(f, li, index);
        if (!resvInvalid(T)) {
            float rnd = rng.next();
            if (R.M > 0) resvClamp(T, (prm.temporalCap - 1) * R.M);                  // preClampedMerge, restir.h:96
            resvMerge(R, T, rnd);
        }
    }
    Resv hist = R;                                                                   // tempReservoir, restir.cu:188
    resvCheck(hist);
    storeResv(f.resvOut + li, hist);                                                 // restir.cu:211-212
    if (SPATIAL) {
        resvCheck(R);                                                                // restir.cu:191-192
        storeResv(f.resvTemp + li, R);
        float4* q = (float4*)(f.hit + li);
        q[0] = make_float4(nrm.x, nrm.y, nrm.z, __int_as_float(matId));
        q[1] = make_float4(wo.x, wo.y, wo.z, __uint_as_float(rng.x));
        if (f.hitMR) f.hitMR[li] = make_float2(metallic, roughness);
    } else {
        writeRadiance(f, li, shadeReservoir(s, R, type, metallic, roughness, nrm, wo), iter);
    }
}

// ------------------------------------------------------------------------------------------------ unbiased reuse
// RstrParams::unbiased -- an ADDITIONAL mode; the reference has none (its reuse, restir.cu:172-199 / restir.h:61-70, is the biased
// form above).  After Bitterli et al. 2020, Alg. 6, in area measure: a reservoir carries the light POINT y (in the wi slot), its
// unbiased contribution weight W = wSum / (M * phat_owner(y)) (in the dist slot), wSum, M and the light id.  A pixel that reuses
// a sample re-evaluates the target at ITS shading point, phat_q(y) = lum(Le f_q cos_q) * cos_l / d^2, and merges with weight
// phat_q(y) * W * M; spatial merges are normalised by 1 / Z with Z the sum of M over the merged pixels whose own target is
// non-zero for the chosen y; visibility is part of the integrand: ONE shadow ray per pixel, for the final sample, at the
// receiving pixel.  Candidate generation, neighbour selection, similarity tests and the RNG draw pattern are the reference's.
struct UnbPoint { f3 pos, nrm, wo; int type; float metallic, roughness; };
// integrand without visibility and without albedo, (Le f cos_q) * cos_l / d^2, for light point y of triangle light `lightId`
RS_D f3 unbIntegrand(const DevScene& s, const UnbPoint& q, f3 y, int lightId) {
    if (lightId < 0) return mk3(0.f);
    const float4* lp = s.lights + 4 * (size_t)lightId;
    const F8 lB = ldg256(lp + 2);
    const f3 nl = mk3(lB.lo.y, lB.lo.z, lB.lo.w), Le = mk3(lB.hi.x, lB.hi.y, lB.hi.z);
    const f3 pts = y - q.pos;
    const float cl = dot(nl, pts);
    if (cl > -1e-6f) return mk3(0.f);                                  // single-sided emitters (scene.h:415)
    const float d2 = dot(pts, pts);
    const float dist = sqrtf(d2);
    const f3 wi = pts * (1.f / dist);
    const f3 g = Le * materialBSDF(q.type, q.metallic, q.roughness, mk3(1.f), diffuseTerm(mk3(1.f)), q.nrm, q.wo, wi) * satDot(q.nrm, wi);
    return g * (fabsf(dot(nl, wi)) / d2);
}
RS_D float unbTarget(const DevScene& s, const UnbPoint& q, f3 y, int lightId) {
    const float t = luminance(unbIntegrand(s, q, y, lightId));
    return (isNanOrInf(t) || t < 0.f) ? 0.f : t;
}
// reservoir in unbiased form: wi = y, dist = W
RS_D void unbFinalize(const DevScene& s, const UnbPoint& q, Resv& R) {
    const float ph = R.lightId >= 0 ? unbTarget(s, q, R.wi, R.lightId) : 0.f;
    R.dist = (ph > 0.f && R.M > 0) ? R.w / ((float)R.M * ph) : 0.f;
    if (isNanOrInf(R.dist)) R.dist = 0.f;
}
RS_D void unbMerge(const DevScene& s, const UnbPoint& q, Resv& R, const Resv& N, float rnd) {
    const float m = N.lightId >= 0 ? unbTarget(s, q, N.wi, N.lightId) * N.dist * (float)N.M : 0.f;
    R.w += m; R.M += N.M;
    if (rnd * R.w < m) { R.wi = N.wi; R.lightId = N.lightId; }
}

// unbiased counterpart of temporalAndStore: R arrives from candidateLoop in the reference's form {wi, dist}
template <bool SPATIAL>
RS_D void temporalAndStoreUnbiased(const DevScene& s, const FrameDev& f, const RstrParams& prm, int iter, int first, size_t li, int index,
                                   const ShadePoint& sp, Resv R, Rng rng, Stack& stack) {
    UnbPoint q;
    q.pos = sp.pos; q.nrm = sp.nrm; q.wo = sp.wo; q.type = sp.type; q.metallic = sp.metallic; q.roughness = sp.roughness;
    if (R.lightId >= 0) R.wi = sp.pos + R.wi * R.dist;                               // the light point
    if (!first && (prm.reuse & 1)) {
        Resv T = findTemporal(f, li, index);                                         // same tests as restir.cu:20-45
        if (!resvInvalid(T)) {
            float rnd = rng.next();
            if (R.M > 0) resvClamp(T, (prm.temporalCap - 1) * R.M);                  // M cap as preClampedMerge (wSum is not used of T)
            unbMerge(s, q, R, T, rnd);
        }
    }
    resvCheck(R);
    unbFinalize(s, q, R);
    storeResv(f.resvOut + li, R);
    if (SPATIAL) {
        storeResv(f.resvTemp + li, R);
        float4* h = (float4*)(f.hit + li);
        h[0] = make_float4(sp.nrm.x, sp.nrm.y, sp.nrm.z, __int_as_float(sp.matId));
        h[1] = make_float4(sp.wo.x, sp.wo.y, sp.wo.z, __uint_as_float(rng.x));
        f.hitPos[li] = make_float4(sp.pos.x, sp.pos.y, sp.pos.z, 0.f);
        if (f.hitMR) f.hitMR[li] = make_float2(sp.metallic, sp.roughness);
    } else {
        f3 direct = mk3(0.f);
        if (R.lightId >= 0 && R.dist > 0.f && traceOccluded<false>(s, sp.pos, R.wi, stack) == 0) direct = unbIntegrand(s, q, R.wi, R.lightId) * R.dist;
        if (hasNanOrInf(direct)) direct = mk3(0.f);
        writeRadiance(f, li, direct, iter);
    }
}

// all stages in one thread; false = undecided, nothing written
template <bool EXACT, bool SPATIAL>
RS_D bool restirAAfterHit(const DevScene& s, const FrameDev& f, const RstrParams& prm, int iter, int first,
                          int x, int y, Stack& stack, Rng rng, f3 d, const Hit& h) {
    size_t li = planeIndex(f, x, y);
    ShadePoint sp;
    const int status = shadePointOf(s, h, d, sp);
    if (status != 2) { unshadedPixel<SPATIAL>(s, f, li, status, d, iter); return true; }
    Resv R = candidateLoop(s, prm, sp, rng);
    if (prm.unbiased) { temporalAndStoreUnbiased<SPATIAL>(s, f, prm, iter, first, li, y * f.W + x, sp, R, rng, stack); return true; }
    // restir.cu:172-176.  With weight == 0 the test cannot change anything, so the ray is skipped.
    if (R.w != 0.f) {
        int occ = traceOccluded<EXACT>(s, sp.pos, sp.pos + R.wi * R.dist, stack);
        if (occ < 0) return false;
        if (occ) R.w = 0.f;
    }
    temporalAndStore<SPATIAL>(s, f, prm, iter, first, li, y * f.W + x, sp.nrm, sp.wo, sp.matId, sp.type, sp.metallic, sp.roughness, R, rng);
    return true;
}

// phase A on the traced tree: the warp's 32 jittered rays walk as one packet
template <bool SPATIAL>
__global__ void __launch_bounds__(RS_BLOCK, RS_MINB_RESTIR) k_restir_a(const __grid_constant__ DevScene s, const __grid_constant__ FrameDev f,
                                                       const __grid_constant__ CamDev cam, const __grid_constant__ RstrParams prm,
                                                       int looper, int iter, int first) {
    RS_DECLARE_REFSTACK(stack);
    RS_DECLARE_PACKET(pk, 1);
    __shared__ unsigned int t0;
    if (f.rowCost && threadIdx.x == 0) t0 = (unsigned int)clock64();
    int x, y;
    const bool active = pixelOf(f, x, y);
    Rng rng;
    rng.x = 1;
    f3 o = mk3(0.f), d = mk3(0.f, 0.f, 1.f);
    if (active) jitteredRay(f, cam, looper, x, y, rng, o, d);
    const f3 oc = cameraOrigin(cam);
    PRay a = prayBegin(oc, d, active, pk_ta);
    packetWalk<false>(s, oc, a, a, pk_ta, pk_ta, pk_wst);
    if (active) {
        Hit h;
        if (!prayResolve(s, oc, d, a, pk_ta, h) || !restirAAfterHit<false, SPATIAL>(s, f, prm, iter, first, x, y, stack, rng, d, h)) enqueuePixel(f, x, y);
    }
    if (f.rowCost) accountBlock(f, &t0);
}
// reference-order walk for every pixel (RS_TRAVERSAL_EXACT)
template <bool SPATIAL>
__global__ void __launch_bounds__(RS_BLOCK, RS_MINB_RESTIR) k_restir_a_exact(const __grid_constant__ DevScene s, const __grid_constant__ FrameDev f,
                                                             const __grid_constant__ CamDev cam, const __grid_constant__ RstrParams prm,
                                                             int looper, int iter, int first) {
    RS_DECLARE_STACK(stack);
    __shared__ unsigned int t0;
    if (f.rowCost && threadIdx.x == 0) t0 = (unsigned int)clock64();
    int x, y;
    if (pixelOf(f, x, y)) restirAPixelExact<SPATIAL>(s, f, cam, prm, looper, iter, first, x, y, stack);
    if (f.rowCost) accountBlock(f, &t0);
}
// G-buffer + phase A of one pixel in one kernel: the two primary rays of every pixel of the warp's tile (64 rays) share
// ONE packet walk of the tree
template <bool SPATIAL>
__global__ void __launch_bounds__(RS_BLOCK, RS_MINB_FUSED) k_gbuffer_restir_a(const __grid_constant__ DevScene s, const __grid_constant__ FrameDev f,
                                                               const __grid_constant__ CamDev cam, const __grid_constant__ CamDev lastCam,
                                                               const __grid_constant__ RstrParams prm, int looper, int iter, int first) {
    RS_DECLARE_REFSTACK(stack);
    RS_DECLARE_PACKET(pk, 2);
    __shared__ unsigned int t0;
    if (f.rowCost && threadIdx.x == 0) t0 = (unsigned int)clock64();
    int x, y;
    const bool active = pixelOf(f, x, y);
    Rng rng;
    rng.x = 1;
    f3 oC = mk3(0.f), dC = mk3(0.f, 0.f, 1.f), oJ = oC, dJ = dC;
    if (active) {
        cameraRay(cam, x, y, .5f, .5f, oC, dC);                                      // gbuffer.cu:11-23
        jitteredRay(f, cam, looper, x, y, rng, oJ, dJ);                              // restir.cu:129
    }
    Hit hC, hJ;
    bool ok;
    {
        const f3 oc = cameraOrigin(cam);
        PRay a = prayBegin(oc, dC, active, pk_ta), b = prayBegin(oc, dJ, active, pk_tb);
        packetWalk<true>(s, oc, a, b, pk_ta, pk_tb, pk_wst);
        ok = prayResolve(s, oc, dC, a, pk_ta, hC);
        ok = prayResolve(s, oc, dJ, b, pk_tb, hJ) && ok;
    }
    if (active) {
        if (ok) {
            gbufferFinish(s, f, lastCam, x, y, oC, dC, hC);
            ok = restirAAfterHit<false, SPATIAL>(s, f, prm, iter, first, x, y, stack, rng, dJ, hJ);
        }
        if (!ok) enqueuePixel(f, x, y);
    }
    if (f.rowCost) accountBlock(f, &t0);
}
// queued pixels: both stages again, with the reference-order walk
template <bool SPATIAL>
__global__ void __launch_bounds__(RS_BLOCK) k_gbuffer_restir_a_fix(const __grid_constant__ DevScene s, const __grid_constant__ FrameDev f,
                                                                   const __grid_constant__ CamDev cam, const __grid_constant__ CamDev lastCam,
                                                                   const __grid_constant__ RstrParams prm, int looper, int iter, int first) {
    RS_DECLARE_STACK(stack);
    const unsigned n = *f.queueCount;
    // entry i -> lane i / numWarps of warp i % numWarps: the first numWarps entries each get a warp of their own
    // (reference-order walks diverge completely, sharing a warp would serialise them)
    const unsigned numWarps = gridDim.x * (RS_BLOCK / 32), gw = blockIdx.x * (RS_BLOCK / 32) + (threadIdx.x >> 5);
    for (unsigned i = (threadIdx.x & 31) * numWarps + gw; i < n; i += 32 * numWarps) {
        int idx = f.queue[i];
        gbufferPixelExact(s, f, cam, lastCam, idx % f.W, idx / f.W, stack);
        restirAPixelExact<SPATIAL>(s, f, cam, prm, looper, iter, first, idx % f.W, idx / f.W, stack);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { atomicAdd(s.fallbackRays + 0, n); atomicAdd(s.fallbackRays + 1, n); }
}

template <bool SPATIAL>
__global__ void __launch_bounds__(RS_BLOCK) k_restir_a_fix(const __grid_constant__ DevScene s, const __grid_constant__ FrameDev f,
                                                           const __grid_constant__ CamDev cam, const __grid_constant__ RstrParams prm,
                                                           int looper, int iter, int first) {
    RS_DECLARE_STACK(stack);
    const unsigned n = *f.queueCount;
    const unsigned numWarps = gridDim.x * (RS_BLOCK / 32), gw = blockIdx.x * (RS_BLOCK / 32) + (threadIdx.x >> 5);
    for (unsigned i = (threadIdx.x & 31) * numWarps + gw; i < n; i += 32 * numWarps) {
        int idx = f.queue[i];
        restirAPixelExact<SPATIAL>(s, f, cam, prm, looper, iter, first, idx % f.W, idx / f.W, stack);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(s.fallbackRays + 1, n);
}

// ------------------------------------------------------------------------------------------------ staged phase A
// The same work as k_gbuffer_restir_a, cut where the shape of the parallelism changes (default pipeline):
//   k_primary      one packet walk per 8x4 tile for both primary rays of every pixel; G-buffer; pixels whose jittered ray
//                  missed or hit an emitter are finished here; the others are appended to a compact queue
//   k_candidates   one thread per QUEUED pixel: 32 light candidates (restir.cu:155-169) -- full warps, no sky lanes
//   k_shadow       shadow rays (restir.cu:172-176) by persistent warps that refill finished lanes from the queue: shadow rays of
//                  neighbouring pixels go to unrelated lights, so lockstep walks of a fixed set of 32 keep 6 lanes busy
//                  (scripts/travsim.cpp); here a lane takes the next ray as soon as enough lanes are idle
//   k_temporal     temporal merge + stores (restir.cu:180-192, 211-230) per queued pixel
// Between the stages a pixel's state lives in planes that are free at that point: the {prim, barycentrics} of the
// jittered hit in HitRec (rewritten by k_candidates), the reservoir before the shadow test in resvTemp (resvTemp2 when spatial
// reuse is off: the reference leaves reservoirTemp alone then, restir.cu:190), the shading
// point in this frame's history plane resvOut (written for good by k_temporal).  Same arithmetic, same RNG stream per
// pixel: every buffer is bit-identical to the fused kernel's (test_staged_pipeline_equals_fused_kernel).
#ifndef RS_MINB_PRIMARY
#define RS_MINB_PRIMARY 8
#endif
#ifndef RS_MINB_CAND
#define RS_MINB_CAND 8
#endif
#ifndef RS_MINB_SHADOW
#define RS_MINB_SHADOW 8
#endif
#ifndef RS_REFILL_MIN
#define RS_REFILL_MIN 8       /* k_shadow refills when at least this many lanes are idle */
#endif
#ifndef RS_LEAF_MIN
#define RS_LEAF_MIN 8         /* k_shadow tests triangles when at least this many lanes wait at a leaf */
#endif

template <bool SPATIAL>
__global__ void __launch_bounds__(RS_BLOCK, RS_MINB_PRIMARY) k_primary(const __grid_constant__ DevScene s, const __grid_constant__ FrameDev f,
                                                           const __grid_constant__ CamDev cam, const __grid_constant__ CamDev lastCam,
                                                           int looper, int iter) {
    RS_DECLARE_PACKET(pk, 2);
    __shared__ unsigned int t0;
    if (f.rowCost && threadIdx.x == 0) t0 = (unsigned int)clock64();
    int x, y;
    const bool active = pixelOf(f, x, y);
    Rng rng;
    rng.x = 1;
    f3 oC = mk3(0.f), dC = mk3(0.f, 0.f, 1.f), oJ = oC, dJ = dC;
    if (active) {
        cameraRay(cam, x, y, .5f, .5f, oC, dC);                                      // gbuffer.cu:11-23
        jitteredRay(f, cam, looper, x, y, rng, oJ, dJ);                              // restir.cu:129
    }
    Hit hC, hJ;
    bool ok;
    {
        const f3 oc = cameraOrigin(cam);
        PRay a = prayBegin(oc, dC, active, pk_ta), b = prayBegin(oc, dJ, active, pk_tb);
        packetWalk<true>(s, oc, a, b, pk_ta, pk_tb, pk_wst);
        ok = prayResolve(s, oc, dC, a, pk_ta, hC);
        ok = prayResolve(s, oc, dJ, b, pk_tb, hJ) && ok;
    }
    bool shaded = false;
    if (active) {
        if (!ok) enqueuePixel(f, x, y);                                              // both stages again in the fix-up kernel
        else {
            gbufferFinish(s, f, lastCam, x, y, oC, dC, hC);
            const size_t li = planeIndex(f, x, y);
            int status = 0;
            if (hJ.prim >= 0) {
                const int matId = __float_as_int(__ldg(s.triGeom + 3 * (size_t)__ldg(s.primToFast + hJ.prim) + 2).y);
                status = __ldg(&s.materials[matId].type) == 4 ? 1 : 2;              // restir.cu:143
            }
            if (status != 2) unshadedPixel<SPATIAL>(s, f, li, status, dJ, iter);
            else {
                shaded = true;
                *(float4*)(f.hit + li) = make_float4(hJ.bx, hJ.by, __int_as_float(hJ.prim), 0.f);
            }
        }
    }
    // warp-aggregated append: the queue keeps the pixels of a tile together
    const unsigned m = __ballot_sync(0xffffffffu, shaded);
    if (m) {
        const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
        unsigned base = 0;
        if (lane == leader) base = atomicAdd(f.shadeCount, (unsigned)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (shaded) f.shadeQueue[base + __popc(m & ((1u << lane) - 1u))] = y * f.W + x;
    }
    if (f.rowCost) accountBlock(f, &t0);
}

// profile for the strip cuts: a thread's share of the cycles its block held its slot, booked on the pixel's row group
RS_D void accountPixel(const FrameDev& f, int y, long long t0) {
    atomicAdd(f.rowCost + (y >> 3), (unsigned long long)((clock64() - t0) / RS_BLOCK));
}

__global__ void __launch_bounds__(RS_BLOCK, RS_MINB_CAND) k_candidates(const __grid_constant__ DevScene s, const __grid_constant__ FrameDev f,
                                                           const __grid_constant__ CamDev cam, const __grid_constant__ RstrParams prm, int looper) {
    const unsigned i = blockIdx.x * RS_BLOCK + threadIdx.x;
    if (i >= f.shadeCount[0]) return;
    const long long t0 = f.rowCost ? clock64() : 0;
    const int index = f.shadeQueue[i];
    const int x = index % f.W, y = index / f.W;
    const size_t li = planeIndex(f, x, y);
    Rng rng;
    f3 o, d;
    jitteredRay(f, cam, looper, x, y, rng, o, d);
    const float4 q0 = *(const float4*)(f.hit + li);
    Hit h;
    h.bx = q0.x; h.by = q0.y; h.prim = __float_as_int(q0.z); h.t = 0.f;
    ShadePoint sp;
    shadePointOf(s, h, d, sp);
    const Resv R = candidateLoop(s, prm, sp, rng);
    storeResv(f.resvStage + li, R);                                                  // before the shadow test
    float4* sc = (float4*)(f.resvOut + li);                                          // free until k_temporal writes the history
    sc[0] = make_float4(sp.pos.x, sp.pos.y, sp.pos.z, sp.metallic);
    sc[1] = make_float4(sp.roughness, 0.f, 0.f, 0.f);
    float4* q = (float4*)(f.hit + li);
    q[0] = make_float4(sp.nrm.x, sp.nrm.y, sp.nrm.z, __int_as_float(sp.matId));
    q[1] = make_float4(sp.wo.x, sp.wo.y, sp.wo.z, __uint_as_float(rng.x));
    if (f.rowCost) accountPixel(f, y, t0);
}

// any-hit test of one leaf of the traced tree with the reference's per-triangle criterion (scene.h:286-316 reduced as above)
RS_D bool leafOccluded(const DevScene& s, const RayT& r, float dist, int ref) {
    const int first = ref & 0x07ffffff, count = ((ref >> 27) & 7) + 1;
    for (int i = 0; i < count; i++) {
        const Tri t = loadTriFast(s, first + i);
        float bx, by, d, tb;
        if (triHit(r, t.v0, t.v1, t.v2, bx, by, d) && d < dist && leafBox(r, t, tb) && tb < dist) return true;
    }
    return false;
}
#ifndef RS_SHADOW_ORDER
#define RS_SHADOW_ORDER 1      /* any-hit rays enter the nearer child first: -4 % of the 1080p / 4K frame on the 1M-triangle scene (profiles/r02_c15_ab.txt) */
#endif
#ifndef RS_DRAIN_MAX_PIXELS
#define RS_DRAIN_MAX_PIXELS 1500000   /* k_shadow<true> (the warp finishes its last rays together) for launches over at most this many pixels */
#endif
#define RS_COOP_CAP (RS_SMEM_STACK * 32)

// the shadow ray of queued pixel li (restir.cu:172-176 -> traceOccluded, scene.h:286-295): from the shaded point towards the
// reservoir's sample; false when the reservoir has no weight (nothing to test) or the segment is empty
RS_D bool shadowRayOf(const FrameDev& f, size_t li, f3& ro, f3& rd, float& dist) {
    const float4* rp = (const float4*)(f.resvStage + li);
    const float4 ra = rp[0];                                             // {wi, dist}
    const float w = rp[1].x;
    if (w == 0.f) return false;                                          // weight 0 cannot change
    const float4 pp = *(const float4*)(f.resvOut + li);
    const f3 pos = mk3(pp.x, pp.y, pp.z);
    const f3 to = pos + mk3(ra.x, ra.y, ra.z) * ra.w;
    f3 dir = to - pos;
    dist = length(dir);
    if (!(dist > 0.f)) return false;
    dir = dir / dist;
    ro = pos + dir * 1e-5f; rd = dir;
    dist -= 1e-4f * 2.f;
    return true;
}

// ---- k_shadow once the queue is empty.  What is left in a warp then is a handful of rays of very different lengths (a shadow ray
// takes 72 node steps on average and up to ~500), and walked one ray per lane the warp would last as long as its longest
// ray with most lanes idle -- measured, the last 350 us of the kernel on every scene and strip.  Any-hit rays need no order,
// so the warp finishes them TOGETHER: every pending node of every remaining ray becomes an item {owner lane, node} of one
// warp-wide list (kept in the lanes' stack rows in shared memory), each round the lanes take 32 items, fetch the owner's
// ray by shuffle, test both children, append the children that are hit and test hit leaves at once.  The drain lasts as long
// as the deepest chain of dependent node fetches, not the longest ray.  A function of its own (not inlined) so that the
// per-lane loop keeps its registers.  Called by the whole warp; li is the lane's pixel, cur its current node (RS_DONE: no ray),
// sRef / lRef its stack of sp entries.
__device__ __noinline__ void shadowDrain(const DevScene& s, const FrameDev& f, int cur, int sp, size_t li, int* sRef, const int* lRef) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    bool act = cur != RS_DONE;
    f3 ro = mk3(0.f), rd = mk3(0.f, 0.f, 1.f);                           // the lane's ray again, from its pixel (the caller keeps its registers)
    float dist = 0.f;
    if (act) shadowRayOf(f, li, ro, rd, dist);
    RayF rf;
    {
        RayT r0;
        r0.o = ro; r0.d = rd;
        rf = makeRayF(r0);
    }
    int* list = sRef - lane;                                            // item p lives at list[(p >> 5) * RS_BLOCK + (p & 31)]
    const unsigned lt = (1u << lane) - 1u;
    unsigned cnt = 0, doneMask = 0;
    const unsigned maxsp = __reduce_max_sync(FULL, act ? (unsigned)sp : 0u);
    {
        const RayT r = makeRayT(ro, rd);
        for (unsigned k = 0; k <= maxsp; k++) {                         // round k: every lane's k-th stacked entry; last round: its current node
            int ref = RS_DONE;
            if (act) {
                if (k < (unsigned)sp) ref = k < RS_SMEM_STACK ? sRef[k * RS_BLOCK] : lRef[k - RS_SMEM_STACK];
                else if (k == maxsp) ref = cur;
            }
            __syncwarp();                                               // row k has been read before items land in it
            bool occ = false;
            if (ref < 0) occ = leafOccluded(s, r, dist, ref);
            const bool push = ref >= 0 && ref != RS_DONE;
            const unsigned pm = __ballot_sync(FULL, push);
            if (push) { const unsigned p = cnt + __popc(pm & lt); list[(p >> 5) * RS_BLOCK + (p & 31)] = (int)(((unsigned)lane << 27) | (unsigned)ref); }
            cnt += __popc(pm);
            if (occ) { f.resvStage[li].weight = 0.f; act = false; }
            doneMask |= __ballot_sync(FULL, occ);
        }
    }
    __syncwarp();
    while (cnt) {
        const unsigned take = cnt > RS_COOP_CAP - 64 ? 1u : (cnt < 32u ? cnt : 32u);
        unsigned item = 0;
        bool valid = (unsigned)lane < take;
        if (valid) { const unsigned p = cnt - 1u - (unsigned)lane; item = (unsigned)list[(p >> 5) * RS_BLOCK + (p & 31)]; }
        cnt -= take;
        __syncwarp();
        const int owner = valid ? (int)(item >> 27) : lane;
        valid = valid && !((doneMask >> owner) & 1u);
        RayF of;
        of.inv.x = __shfl_sync(FULL, rf.inv.x, owner); of.inv.y = __shfl_sync(FULL, rf.inv.y, owner); of.inv.z = __shfl_sync(FULL, rf.inv.z, owner);
        of.oi.x = __shfl_sync(FULL, rf.oi.x, owner); of.oi.y = __shfl_sync(FULL, rf.oi.y, owner); of.oi.z = __shfl_sync(FULL, rf.oi.z, owner);
        const float od = __shfl_sync(FULL, dist, owner);
        int c0 = RS_DONE, c1 = RS_DONE;                                 // children that are hit
        if (valid) {
            const float4* np = s.fastNodes + 4 * (size_t)(item & 0x07ffffffu);
            const F8 nA = ldg256(np), nB = ldg256(np + 2);
            const float4 a = nA.lo, b = nA.hi, c = nB.lo;
            float tL, tR;
            if (slabHit(of, a.x, a.y, a.z, a.w, b.x, b.y, od, tL)) c0 = __float_as_int(nB.hi.x);
            if (slabHit(of, b.z, b.w, c.x, c.y, c.z, c.w, od, tR)) c1 = __float_as_int(nB.hi.y);
        }
        for (int side = 0; side < 2; side++) {
            const int ch = side ? c1 : c0;
            const bool push = ch >= 0 && ch != RS_DONE;
            const unsigned pm = __ballot_sync(FULL, push);
            if (push) { const unsigned p = cnt + __popc(pm & lt); list[(p >> 5) * RS_BLOCK + (p & 31)] = (int)(((unsigned)owner << 27) | (unsigned)ch); }
            cnt += __popc(pm);
        }
        bool occ = false;
        if (__any_sync(FULL, c0 < 0 || c1 < 0)) {
            f3 oo, dd;
            oo.x = __shfl_sync(FULL, ro.x, owner); oo.y = __shfl_sync(FULL, ro.y, owner); oo.z = __shfl_sync(FULL, ro.z, owner);
            dd.x = __shfl_sync(FULL, rd.x, owner); dd.y = __shfl_sync(FULL, rd.y, owner); dd.z = __shfl_sync(FULL, rd.z, owner);
            if (c0 < 0 || c1 < 0) {
                const RayT rr = makeRayT(oo, dd);
                if (c0 < 0) occ = leafOccluded(s, rr, od, c0);
                if (!occ && c1 < 0) occ = leafOccluded(s, rr, od, c1);
            }
        }
        const unsigned occBits = __reduce_or_sync(FULL, occ ? (1u << owner) : 0u);
        doneMask |= occBits;
        if ((occBits >> lane) & 1u) f.resvStage[li].weight = 0.f;        // this lane's own ray
        __syncwarp();
    }
}

#ifdef RS_SHADOW_STATS
// development build only (RSTR_DEFINES=-DRS_SHADOW_STATS): steps per shadow ray as a log2 histogram [0..31], [32] rays, [33] steps,
// [34] longest ray, [40] / [41] / [42]: globaltimer (ns) of the first warp that found the queue empty / the last warp's exit / start
__device__ unsigned long long g_shadowStats[64];
RS_D unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
RS_D void shadowRayDone(int steps) {
#if RS_SHADOW_STATS >= 2      /* time stamps only: no per-ray atomics that would stretch the very tail they are meant to measure */
    return;
#endif
    int b = 0;
    while ((2 << b) <= steps && b < 31) b++;
    atomicAdd(g_shadowStats + b, 1ull); atomicAdd(g_shadowStats + 32, 1ull); atomicAdd(g_shadowStats + 33, (unsigned long long)steps);
    atomicMax(g_shadowStats + 34, (unsigned long long)steps);
}
#endif
// DevScene::testOcclusion (scene.h:286-316) for every queued pixel whose reservoir has weight: occluded -> weight = 0
// DRAIN: the tail is finished by shadowDrain.  It shortens the kernel by ~80 us whatever the queue length, but the build that can
// leave the loop early runs the per-lane phase ~12 % slower (measured: profiles/r02_c14_tail.txt), so it is the choice for
// short queues only (strips, small frames): launchPhaseAStaged picks by the launch's pixel count.
template <bool DRAIN>
__global__ void __launch_bounds__(RS_BLOCK, RS_MINB_SHADOW) k_shadow(const __grid_constant__ DevScene s, const __grid_constant__ FrameDev f) {
    RS_DECLARE_REFSTACK(stack);
    const unsigned FULL = 0xffffffffu;
    const unsigned n = f.shadeCount[0];
    const int lane = threadIdx.x & 31;
    RayT r;
    RayF rf;
    float dist = 0.f;
    int cur = RS_DONE, sp = 0, rowY = 0;
    size_t li = 0;
    long long tRay = 0;
    bool exhausted = false;                       // warp-uniform: the queue is empty
    bool drain = false;                           // warp-uniform: leave the loop for shadowDrain
    const unsigned sAddr = stackAddr(stack);
    r.o = r.d = r.inv = mk3(0.f); r.flags = 0; r.dim = r.lesser = 0;
    rf.inv = rf.oi = mk3(0.f);
#ifdef RS_SHADOW_STATS
    int steps = 0;
    if (lane == 0) atomicMin(g_shadowStats + 42, gtimer());
#endif
    for (;;) {
        const unsigned idle = __ballot_sync(FULL, cur == RS_DONE);
#ifdef RS_SHADOW_STATS
        if (idle == FULL && exhausted && lane == 0) atomicMax(g_shadowStats + 41, gtimer());
#endif
        if (idle == FULL && exhausted) break;
        if (!exhausted && (__popc(idle) >= RS_REFILL_MIN || idle == FULL)) {
            unsigned base = 0;
            const int leader = __ffs(idle) - 1;
            if (lane == leader) base = atomicAdd(f.shadeCount + 1, (unsigned)__popc(idle));
            base = __shfl_sync(FULL, base, leader);
            if (base >= n) exhausted = true;
#ifdef RS_SHADOW_STATS
            if (exhausted && lane == 0) atomicMin(g_shadowStats + 40, gtimer());
#endif
            const unsigned i = base + __popc(idle & ((1u << lane) - 1u));
            if (cur == RS_DONE && i < n) {
                const int index = f.shadeQueue[i];
                rowY = index / f.W;
                li = planeIndex(f, index % f.W, rowY);
                f3 ro, rd;
                if (shadowRayOf(f, li, ro, rd, dist)) {
                    r = makeRayT(ro, rd);
                    rf = makeRayF(r);
                    float tr;
                    if (slabHit(rf, s.fastRootMin[0], s.fastRootMin[1], s.fastRootMin[2], s.fastRootMax[0], s.fastRootMax[1], s.fastRootMax[2], dist, tr)) {
                        cur = s.fastRoot; sp = 0;
                        if (f.rowCost) tRay = clock64();
#ifdef RS_SHADOW_STATS
                        steps = 0;
#endif
                    }
                }
            }
        }
        // the queue is empty: the warp finishes what it still holds together (shadowDrain, after the loop), as soon as that fits the list
        if (DRAIN && exhausted) {
            const unsigned need = __reduce_add_sync(FULL, cur != RS_DONE ? (unsigned)sp + 1u : 0u);
            if (need <= RS_COOP_CAP - 96) { drain = true; break; }
        }
        // Lanes at an internal node step; lanes that have reached a leaf WAIT until RS_LEAF_MIN of them are there (or nobody is
        // left to step): run lane by lane, the triangle tests -- half of this kernel's instructions -- executed with 1.5 of 32 lanes.
        const unsigned atLeaf = __ballot_sync(FULL, cur < 0);
        const unsigned atNode = __ballot_sync(FULL, cur >= 0 && cur != RS_DONE);
        if (atNode && __popc(atLeaf) < RS_LEAF_MIN) {
            if (cur >= 0 && cur != RS_DONE) {
                const float4* np = s.fastNodes + 4 * (size_t)cur;
                const F8 nA = ldg256(np), nB = ldg256(np + 2);
                const float4 a = nA.lo, b = nA.hi, c = nB.lo;
                const int2 l = make_int2(__float_as_int(nB.hi.x), __float_as_int(nB.hi.y));
                float tL, tR;
                const bool hL = slabHit(rf, a.x, a.y, a.z, a.w, b.x, b.y, dist, tL);
                const bool hR = slabHit(rf, b.z, b.w, c.x, c.y, c.z, c.w, dist, tR);
#if RS_SHADOW_ORDER
                if (hL && hR) { const bool ln = tL <= tR; stackPush(sAddr, stack, sp, ln ? l.y : l.x); sp++; cur = ln ? l.x : l.y; }     // nearer child first (A/B)
#else
                if (hL && hR) { stackPush(sAddr, stack, sp, l.y); sp++; cur = l.x; }
#endif
                else if (hL) cur = l.x;
                else if (hR) cur = l.y;
                else cur = sp == 0 ? RS_DONE : stackRef(sAddr, stack, --sp);
                if (f.rowCost && cur == RS_DONE) atomicAdd(f.rowCost + (rowY >> 3), (unsigned long long)((clock64() - tRay) / 24));
#ifdef RS_SHADOW_STATS
                steps++;
                if (cur == RS_DONE) shadowRayDone(steps);
#endif
            }
        } else if (cur < 0) {
            const bool occluded = leafOccluded(s, r, dist, cur);
            if (occluded) { f.resvStage[li].weight = 0.f; cur = RS_DONE; }
            else cur = sp == 0 ? RS_DONE : stackRef(sAddr, stack, --sp);
            if (f.rowCost && cur == RS_DONE) atomicAdd(f.rowCost + (rowY >> 3), (unsigned long long)((clock64() - tRay) / 24));
#ifdef RS_SHADOW_STATS
            steps++;
            if (cur == RS_DONE) shadowRayDone(steps);
#endif
        }
    }
    if (DRAIN && drain) {
        shadowDrain(s, f, cur, sp, li, stack.sRef, stack.lRef);
#ifdef RS_SHADOW_STATS
        if (lane == 0) atomicMax(g_shadowStats + 41, gtimer());
#endif
    }
}
#ifdef RS_SHADOW_STATS
extern "C" int rstr_debug_shadow_stats(unsigned long long* out64, int reset) {
    cudaDeviceSynchronize();
    if (out64) cudaMemcpyFromSymbol(out64, g_shadowStats, sizeof g_shadowStats);
    if (reset) {
        unsigned long long z[64] = {};
        z[40] = z[42] = ~0ull;
        cudaMemcpyToSymbol(g_shadowStats, z, sizeof z);
    }
    return 0;
}
#endif

template <bool SPATIAL>
__global__ void __launch_bounds__(RS_BLOCK) k_temporal(const __grid_constant__ DevScene s, const __grid_constant__ FrameDev f,
                                                       const __grid_constant__ RstrParams prm, int iter, int first) {
    const unsigned i = blockIdx.x * RS_BLOCK + threadIdx.x;
    if (i >= f.shadeCount[0]) return;
    const long long t0 = f.rowCost ? clock64() : 0;
    const int index = f.shadeQueue[i];
    const int x = index % f.W, y = index / f.W;
    const size_t li = planeIndex(f, x, y);
    const float4* q = (const float4*)(f.hit + li);
    const float4 h0 = q[0], h1 = q[1];
    const float4* sc = (const float4*)(f.resvOut + li);
    const float metallic = sc[0].w, roughness = sc[1].x;
    const int matId = __float_as_int(h0.w);
    Rng rng;
    rng.x = __float_as_uint(h1.w);
    const Resv R = loadResv(f.resvStage + li);
    temporalAndStore<SPATIAL>(s, f, prm, iter, first, li, index, mk3(h0.x, h0.y, h0.z), mk3(h1.x, h1.y, h1.z), matId,
                              __ldg(&s.materials[matId].type), metallic, roughness, R, rng);
    if (f.rowCost) accountPixel(f, y, t0);
}

// unbiased mode: k_temporal's counterpart (k_shadow is not launched: visibility is tested once, for the final sample)
template <bool SPATIAL>
__global__ void __launch_bounds__(RS_BLOCK) k_temporal_unb(const __grid_constant__ DevScene s, const __grid_constant__ FrameDev f,
                                                           const __grid_constant__ RstrParams prm, int iter, int first) {
    RS_DECLARE_REFSTACK(stack);
    const unsigned i = blockIdx.x * RS_BLOCK + threadIdx.x;
    if (i >= f.shadeCount[0]) return;
    const int index = f.shadeQueue[i];
    const int x = index % f.W, y = index / f.W;
    const size_t li = planeIndex(f, x, y);
    const float4* q = (const float4*)(f.hit + li);
    const float4 h0 = q[0], h1 = q[1];
    const float4* sc = (const float4*)(f.resvOut + li);
    const float4 s0 = sc[0];
    ShadePoint sp;
    sp.pos = mk3(s0.x, s0.y, s0.z); sp.metallic = s0.w; sp.roughness = sc[1].x;
    sp.nrm = mk3(h0.x, h0.y, h0.z); sp.wo = mk3(h1.x, h1.y, h1.z);
    sp.matId = __float_as_int(h0.w);
    sp.type = __ldg(&s.materials[sp.matId].type);
    Rng rng;
    rng.x = __float_as_uint(h1.w);
    const Resv R = loadResv(f.resvStage + li);
    temporalAndStoreUnbiased<SPATIAL>(s, f, prm, iter, first, li, index, sp, R, rng, stack);
}

#ifndef RS_SPATIAL_GATHER_TOGETHER
#define RS_SPATIAL_GATHER_TOGETHER 1
#endif
// restir.cu:47-85 (+ mathUtil.h:128-132 toConcentricDisk)
RS_D Resv findSpatial(const FrameDev& f, const ResvD* src, int x, int y, int ownMat, float4 g, float rx, float ry, float radius) {
    float rr = sqrtf(rx);
    float theta = ry * RS_PI * 2.0f;
    float px_ = cosf(theta) * rr * radius, py_ = sinf(theta) * rr * radius;
    int px = __float2int_rz((float)x + .5f + px_);
    int py = __float2int_rz((float)y + .5f + py_);
    if (px < 0 || px >= f.W || py < 0 || py >= f.H || (px == x && py == y)) return emptyResv();
    if (!rowResident(f, py)) { atomicAdd(f.haloMiss, 1u); return emptyResv(); }
    size_t pli = planeIndex(f, px, py);
#if RS_SPATIAL_GATHER_TOGETHER
    // The neighbour's material id, geometry record and reservoir are requested TOGETHER (three independent loads in flight)
    // instead of one after the other behind the two rejection tests: the next neighbour cannot be drawn before this one
    // is decided (a rejected try draws no sample1D, restir.cu:93-99, so the RNG stream -- and the next position -- depends
    // on it), which makes the chain of dependent L2 round trips per try what bounds this kernel, not its bytes.
    const int pm = f.matId[0][pli];
    const float4 pg = f.geom[0][pli];
    const Resv N = loadResv(src + pli);
    if (pm != ownMat) return emptyResv();
    bool diff = dot(mk3(g.x, g.y, g.z), mk3(pg.x, pg.y, pg.z)) < .9f;
    if (fabsf(g.w - pg.w) > g.w * .1f) diff = true;
    return diff ? emptyResv() : N;
#else
    if (f.matId[0][pli] != ownMat) return emptyResv();
    float4 pg = f.geom[0][pli];
    bool diff = dot(mk3(g.x, g.y, g.z), mk3(pg.x, pg.y, pg.z)) < .9f;
    if (fabsf(g.w - pg.w) > g.w * .1f) diff = true;
    if (diff) return emptyResv();
    return loadResv(src + pli);
#endif
}

// One spatial pass.  pass 1 = restir.cu:196-199; passes 2.. = the commented-out block restir.cu:201-209
// (preClampedMerge<4>).  `src` holds every pixel's reservoir of the previous stage, `dst` receives this pass's result
// when another pass follows; the last pass shades.  The reference publishes through ONE buffer (reservoirTemp) between
// grid barriers; with two buffers the same content is kept by copying un-shaded pixels through and by leaving the
// last published reservoir in f.resvTemp.
__global__ void __launch_bounds__(128) k_restir_b(const __grid_constant__ DevScene s, const __grid_constant__ FrameDev f,
                                                  const __grid_constant__ RstrParams prm, int iter,
                                                  const ResvD* __restrict__ src, ResvD* __restrict__ dst, int pass, int last) {
    int x, y;
    if (!pixelOf(f, x, y)) return;
    size_t li = planeIndex(f, x, y);
    float4* q = (float4*)(f.hit + li);
    float4 h0 = q[0], h1 = q[1];
    int matId = __float_as_int(h0.w);
    if (matId < 0) {                                // phase A already wrote this pixel's radiance
        if (!last) { const float4* a = (const float4*)(src + li); float4* b = (float4*)(dst + li); b[0] = a[0]; b[1] = a[1]; }
        else if (src != f.resvTemp) { const float4* a = (const float4*)(src + li); float4* b = (float4*)(f.resvTemp + li); b[0] = a[0]; b[1] = a[1]; }
        return;
    }
    Rng rng; rng.x = __float_as_uint(h1.w);
    Resv R = loadResv(src + li);
    const Resv published = R;
    Resv S = emptyResv();                                                            // restir.cu:87-100
    const int ownMat = f.matId[0][li];
    const float4 ownGeom = f.geom[0][li];
    for (int i = 0; i < prm.numSpatial; i++) {
        float rx = rng.next(), ry = rng.next();
        Resv N = findSpatial(f, src, x, y, ownMat, ownGeom, rx, ry, prm.spatialRadius);
        if (!resvInvalid(N)) resvMerge(S, N, rng.next());
    }
    if (pass == 1) {
        if (!resvInvalid(S) && !resvInvalid(R)) resvMerge(R, S, rng.next());         // restir.cu:197-199
    } else if (!resvInvalid(S)) {                                                    // restir.cu:206-208
        float rnd = rng.next();
        if (R.M > 0) resvClamp(S, (4 - 1) * R.M);
        resvMerge(R, S, rnd);
    }
    if (!last) {
        storeResv(dst + li, R);
        q[1] = make_float4(h1.x, h1.y, h1.z, __uint_as_float(rng.x));
        return;
    }
    if (src != f.resvTemp) storeResv(f.resvTemp + li, published);
    f3 nrm = mk3(h0.x, h0.y, h0.z), wo = mk3(h1.x, h1.y, h1.z);
    const RstrMaterial* m = s.materials + matId;
    float metallic, roughness;
    if (f.hitMR) { float2 mr = f.hitMR[li]; metallic = mr.x; roughness = mr.y; }
    else { metallic = __ldg(&m->metallic); roughness = __ldg(&m->roughness); }
    writeRadiance(f, li, shadeReservoir(s, R, __ldg(&m->type), metallic, roughness, nrm, wo), iter);
}

// restir.cu:47-85 without the reservoir fetch: plane index of the neighbour drawn by (rx, ry), or -1 when it is rejected
RS_D long long spatialNeighbour(const FrameDev& f, int x, int y, size_t li, float rx, float ry, float radius) {
    float rr = sqrtf(rx);
    float theta = ry * RS_PI * 2.0f;
    float px_ = cosf(theta) * rr * radius, py_ = sinf(theta) * rr * radius;
    int px = __float2int_rz((float)x + .5f + px_);
    int py = __float2int_rz((float)y + .5f + py_);
    if (px < 0 || px >= f.W || py < 0 || py >= f.H || (px == x && py == y)) return -1;
    if (!rowResident(f, py)) { atomicAdd(f.haloMiss, 1u); return -1; }
    size_t pli = planeIndex(f, px, py);
    if (f.matId[0][pli] != f.matId[0][li]) return -1;
    float4 g = f.geom[0][li], pg = f.geom[0][pli];
    bool diff = dot(mk3(g.x, g.y, g.z), mk3(pg.x, pg.y, pg.z)) < .9f;
    if (fabsf(g.w - pg.w) > g.w * .1f) diff = true;
    return diff ? -1 : (long long)pli;
}
RS_D UnbPoint unbPointOf(const DevScene& s, const FrameDev& f, size_t li) {
    const float4* q = (const float4*)(f.hit + li);
    const float4 h0 = q[0], hp = f.hitPos[li];
    // wo as three words: the fourth word of that float4 is the pixel's RNG state, which its owner rewrites between spatial passes while
    // neighbours read its wo (k_restir_b_unb, !last) -- a 16-byte load here would overlap that store
    const float* w3 = (const float*)(q + 1);
    UnbPoint p;
    p.pos = mk3(hp.x, hp.y, hp.z); p.nrm = mk3(h0.x, h0.y, h0.z); p.wo = mk3(w3[0], w3[1], w3[2]);
    const int matId = __float_as_int(h0.w);
    const RstrMaterial* m = s.materials + (matId < 0 ? 0 : matId);
    p.type = matId < 0 ? 2 : __ldg(&m->type);                          // a pixel phase A did not shade: target 0 everywhere
    if (f.hitMR) { float2 mr = f.hitMR[li]; p.metallic = mr.x; p.roughness = mr.y; }
    else { p.metallic = __ldg(&m->metallic); p.roughness = __ldg(&m->roughness); }
    return p;
}

// One spatial pass of the unbiased mode (see "unbiased reuse" above): the reference's neighbour draws and tests, every merged
// sample re-evaluated here, 1 / Z normalisation, and on the last pass ONE shadow ray for the chosen sample.
__global__ void __launch_bounds__(RS_BLOCK) k_restir_b_unb(const __grid_constant__ DevScene s, const __grid_constant__ FrameDev f,
                                                           const __grid_constant__ RstrParams prm, int iter,
                                                           const ResvD* __restrict__ src, ResvD* __restrict__ dst, int pass, int last) {
    RS_DECLARE_REFSTACK(stack);
    int x, y;
    if (!pixelOf(f, x, y)) return;
    size_t li = planeIndex(f, x, y);
    float4* hq = (float4*)(f.hit + li);
    const float4 h1 = hq[1];
    const int matId = __float_as_int(hq[0].w);
    if (matId < 0) {                                // phase A already wrote this pixel's radiance
        if (!last) { const float4* a = (const float4*)(src + li); float4* b = (float4*)(dst + li); b[0] = a[0]; b[1] = a[1]; }
        else if (src != f.resvTemp) { const float4* a = (const float4*)(src + li); float4* b = (float4*)(f.resvTemp + li); b[0] = a[0]; b[1] = a[1]; }
        return;
    }
    const UnbPoint q = unbPointOf(s, f, li);
    Rng rng; rng.x = __float_as_uint(h1.w);
    const Rng rng0 = rng;                           // the neighbour draws are replayed for Z
    const Resv own = loadResv(src + li);
    const int capM = pass == 1 ? 0x7fffffff : 3 * own.M;     // later passes clamp a neighbour's M like preClampedMerge<4> (restir.cu:206)
    Resv S = own;                                   // own sample first: weight phat_q(y_own) * W_own * M_own
    S.w = own.lightId >= 0 ? unbTarget(s, q, own.wi, own.lightId) * own.dist * (float)own.M : 0.f;
    for (int i = 0; i < prm.numSpatial; i++) {
        float rx = rng.next(), ry = rng.next();
        long long pli = spatialNeighbour(f, x, y, li, rx, ry, prm.spatialRadius);
        float rnd = rng.next();                     // the reference draws for every try (a rejected one merges an empty reservoir)
        if (pli < 0) continue;
        Resv N = loadResv(src + pli);
        if (resvInvalid(N)) continue;
        if (N.M > capM) N.M = capM;
        unbMerge(s, q, S, N, rnd);
    }
    // Z: the M of every merged pixel whose OWN target is non-zero for the chosen light point
    float Z = 0.f;
    const float phq = S.lightId >= 0 ? unbTarget(s, q, S.wi, S.lightId) : 0.f;
    if (phq > 0.f) {
        Z = (float)own.M;
        Rng r2 = rng0;
        for (int i = 0; i < prm.numSpatial; i++) {
            float rx = r2.next(), ry = r2.next();
            r2.next();
            long long pli = spatialNeighbour(f, x, y, li, rx, ry, prm.spatialRadius);
            if (pli < 0) continue;
            Resv N = loadResv(src + pli);
            if (resvInvalid(N)) continue;
            if (N.M > capM) N.M = capM;
            const UnbPoint pn = unbPointOf(s, f, (size_t)pli);
            if (unbTarget(s, pn, S.wi, S.lightId) > 0.f) Z += (float)N.M;
        }
    }
    const float W = (phq > 0.f && Z > 0.f) ? S.w / (Z * phq) : 0.f;
    S.dist = isNanOrInf(W) ? 0.f : W;
    S.w = S.dist * phq * (float)S.M;                // consistent {wSum, W, M} for a later pass
    if (!last) {
        storeResv(dst + li, S);
        ((float*)(hq + 1))[3] = __uint_as_float(rng.x);      // only the RNG word: neighbours are reading wo in the first three (unbPointOf)
        return;
    }
    if (src != f.resvTemp) storeResv(f.resvTemp + li, own);
    f3 direct = mk3(0.f);
    if (S.lightId >= 0 && S.dist > 0.f && traceOccluded<false>(s, q.pos, S.wi, stack) == 0) direct = unbIntegrand(s, q, S.wi, S.lightId) * S.dist;
    if (hasNanOrInf(direct)) direct = mk3(0.f);
    writeRadiance(f, li, direct, iter);
}

// pathtrace.cu:279-328 with scene.h:427-459 (occlusion test BEFORE the facing test)
template <bool EXACT>
RS_D bool ptdirectAfterHit(const DevScene& s, const FrameDev& f, int iter, int x, int y, Stack& stack, Rng rng, f3 d, const Hit& h) {
    size_t li = planeIndex(f, x, y);
    f3 direct = mk3(0.f);
    if (h.prim < 0) {
        if (s.envTex >= 0) direct = envLookup(s, d);                                 // pathtrace.cu:295-300
    } else {
        Tri t = loadTri(s, h.prim);
        const float4* np = s.triNorm + 3 * (size_t)h.prim;
        float4 na4 = __ldg(np), nb4 = __ldg(np + 1), nc4 = __ldg(np + 2);
        f3 na = mk3(na4.x, na4.y, na4.z), nb = mk3(na4.w, nb4.x, nb4.y), nc = mk3(nb4.z, nb4.w, nc4.x);
        float bz = 1.f - h.bx - h.by;
        f3 pos = t.v1 * h.bx + t.v2 * h.by + t.v0 * bz;
        f3 nrm = normalize(nb * h.bx + nc * h.by + na * bz);
        Surf m = texturedMaterial(s, t.matId, h.prim, h.bx, h.by, nrm);              // pathtrace.cu:302
        const int type = m.type;
        const f3 baseColor = m.baseColor;
        if (type == 4) direct = baseColor;
        else if (type != 2 && s.numLights > 0) {
            f3 wo = -d;
            if (dot(nrm, wo) < 0.f) nrm = -nrm;
            float c0 = rng.next(), c1 = rng.next(), c2 = rng.next(), c3 = rng.next();
            int len = s.numLights;
            int pass = min(__float2int_rz((float)len * c0), len - 1);
            float2 e = __ldg(s.alias + pass);
            int lightId = (c1 < e.x) ? pass : __float_as_int(e.y);
            if (s.envTex >= 0 && lightId == len - 1) {                               // scene.h:433-435 -> :377-392
                f3 Li, wi;
                envSample(s, c2, c3, Li, wi);
                int occ = traceOccluded<EXACT>(s, pos, pos + wi * 1e6f, stack);
                if (occ < 0) return false;
                float pdf = envPdf(s, Li);
                if (!occ && pdf > 0.f)
                    direct = Li * materialBSDF(type, m.metallic, m.roughness, baseColor, diffuseTerm(baseColor), nrm, wo, wi) * satDot(nrm, wi) / pdf;
            } else {
                const float4* lp = s.lights + 4 * (size_t)lightId;
                float4 a = __ldg(lp), b = __ldg(lp + 1), c = __ldg(lp + 2), d4 = __ldg(lp + 3);
                f3 v0 = mk3(a.x, a.y, a.z), v1 = mk3(a.w, b.x, b.y), v2 = mk3(b.z, b.w, c.x);
                f3 n = mk3(c.y, c.z, c.w);
                float sr = sqrtf(c3);
                float u = 1.f - sr, v = c2 * sr;
                f3 sampled = v1 * u + v2 * v + v0 * (1.f - u - v);
                int occ = traceOccluded<EXACT>(s, pos, sampled, stack);
                if (occ < 0) return false;
                if (!occ) {
                    f3 pts = sampled - pos;
                    if (!(dot(n, pts) > -1e-6f)) {
                        f3 Li = mk3(d4.x, d4.y, d4.z);
                        float len2 = dot(pts, pts);
                        f3 wi = pts * (1.f / sqrtf(len2));
                        float pdf = d4.w * len2 / fabsf(dot(n, wi));
                        if (pdf > 0.f)
                            direct = Li * materialBSDF(type, m.metallic, m.roughness, baseColor, diffuseTerm(baseColor), nrm, wo, wi) * satDot(nrm, wi) / pdf;
                    }
                }
            }
        }
    }
    float* out = f.radiance + 3 * li;
    f3 prev = mk3(out[0], out[1], out[2]);
    f3 v = (prev * (float)iter + direct) / (float)(iter + 1);
    out[0] = v.x; out[1] = v.y; out[2] = v.z;
    return true;
}

RS_D void ptdirectPixelExact(const DevScene& s, const FrameDev& f, const CamDev& cam, int looper, int iter, int x, int y, Stack& stack) {
    Rng rng;
    f3 o, d;
    jitteredRay(f, cam, looper, x, y, rng, o, d);
    RayT ray = makeRayT(o, d);
    Hit h;
    traceClosestExact(s, ray, h, stack);
    ptdirectAfterHit<true>(s, f, iter, x, y, stack, rng, d, h);
}

__global__ void __launch_bounds__(RS_BLOCK) k_ptdirect(const __grid_constant__ DevScene s, const __grid_constant__ FrameDev f,
                                                       const __grid_constant__ CamDev cam, int looper, int iter) {
    RS_DECLARE_REFSTACK(stack);
    RS_DECLARE_PACKET(pk, 1);
    int x, y;
    const bool active = pixelOf(f, x, y);
    Rng rng;
    rng.x = 1;
    f3 o = mk3(0.f), d = mk3(0.f, 0.f, 1.f);
    if (active) jitteredRay(f, cam, looper, x, y, rng, o, d);
    const f3 oc = cameraOrigin(cam);
    PRay a = prayBegin(oc, d, active, pk_ta);
    packetWalk<false>(s, oc, a, a, pk_ta, pk_ta, pk_wst);
    if (active) {
        Hit h;
        if (!prayResolve(s, oc, d, a, pk_ta, h) || !ptdirectAfterHit<false>(s, f, iter, x, y, stack, rng, d, h)) enqueuePixel(f, x, y);
    }
}
__global__ void __launch_bounds__(RS_BLOCK) k_ptdirect_exact(const __grid_constant__ DevScene s, const __grid_constant__ FrameDev f,
                                                             const __grid_constant__ CamDev cam, int looper, int iter) {
    RS_DECLARE_STACK(stack);
    int x, y;
    if (pixelOf(f, x, y)) ptdirectPixelExact(s, f, cam, looper, iter, x, y, stack);
}
__global__ void __launch_bounds__(RS_BLOCK) k_ptdirect_fix(const __grid_constant__ DevScene s, const __grid_constant__ FrameDev f,
                                                           const __grid_constant__ CamDev cam, int looper, int iter) {
    RS_DECLARE_STACK(stack);
    const unsigned n = *f.queueCount;
    const unsigned numWarps = gridDim.x * (RS_BLOCK / 32), gw = blockIdx.x * (RS_BLOCK / 32) + (threadIdx.x >> 5);
    for (unsigned i = (threadIdx.x & 31) * numWarps + gw; i < n; i += 32 * numWarps) {
        int idx = f.queue[i];
        ptdirectPixelExact(s, f, cam, looper, iter, idx % f.W, idx / f.W, stack);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(s.fallbackRays + 2, n);
}

// pathtrace.cu:30-56 (Math::filmic / ACES / correctGamma: mathUtil.h:103-117)
RS_D float calcFilmic(float c) { return (c * (c * 0.22f + 0.03f) + 0.002f) / (c * (c * 0.22f + 0.3f) + 0.06f) - 1.f / 30.f; }
__global__ void __launch_bounds__(256) k_tonemap(const float* __restrict__ radiance, uchar4* __restrict__ ldr, size_t n, int toneMapping, float scale) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    f3 c = mk3(radiance[3 * i], radiance[3 * i + 1], radiance[3 * i + 2]) * scale;
    if (toneMapping == 1) {
        float w = calcFilmic(11.2f);
        c = mk3(calcFilmic(c.x * 1.6f), calcFilmic(c.y * 1.6f), calcFilmic(c.z * 1.6f)) / w;
    } else if (toneMapping == 2) {
        c = (c * (c * 2.51f + mk3(0.03f))) / (c * (c * 2.43f + mk3(0.59f)) + mk3(0.14f));
    }
    c = mk3(powf(c.x, 1.f / 2.2f), powf(c.y, 1.f / 2.2f), powf(c.z, 1.f / 2.2f));
    int r = min(max(__float2int_rz(c.x * 255.f), 0), 255);
    int g = min(max(__float2int_rz(c.y * 255.f), 0), 255);
    int b = min(max(__float2int_rz(c.z * 255.f), 0), 255);
    ldr[i] = make_uchar4((unsigned char)r, (unsigned char)g, (unsigned char)b, 0);
}

// ---- device-internal planes -> reference host layouts ----
__global__ void k_export_geom(const float4* __restrict__ geom, const float4* __restrict__ am, float* albedo, float* normal, float* depth, int* motion, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 g = geom[i], a = am[i];
    if (albedo) { albedo[3 * i] = a.x; albedo[3 * i + 1] = a.y; albedo[3 * i + 2] = a.z; }
    if (normal) { normal[3 * i] = g.x; normal[3 * i + 1] = g.y; normal[3 * i + 2] = g.z; }
    if (depth) depth[i] = g.w;
    if (motion) motion[i] = __float_as_int(a.w);
}
__global__ void k_export_resv(const DevScene s, const ResvD* __restrict__ src, float* out36, int* lightIdx, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Resv r = loadResv(src + i);
    if (lightIdx) lightIdx[i] = r.lightId;
    if (out36) {
        f3 Li = lightLe(s, r.lightId);
        float* o = out36 + 9 * i;
        o[0] = Li.x; o[1] = Li.y; o[2] = Li.z; o[3] = r.wi.x; o[4] = r.wi.y; o[5] = r.wi.z; o[6] = r.dist;
        o[7] = __int_as_float(r.M); o[8] = r.w;
    }
}

// ------------------------------------------------------------------------------------------------ launchers
#ifndef RS_HOST_EMU
static inline dim3 pixelGrid(const FrameDev& f) { return dim3((f.W + 15) / 16, (f.rowHi - f.rowLo + 7) / 8); }

#define RS_FIX_BLOCKS 296   /* 2 per SM; the fix-up kernels stride over the queue */

int launchGBuffer(const DevScene& s, const FrameDev& f, const CamDev& cam, const CamDev& lastCam, cudaStream_t st) {
    if (s.traversal == RS_TRAVERSAL_EXACT) { k_gbuffer_exact<<<pixelGrid(f), RS_BLOCK, 0, st>>>(s, f, cam, lastCam); return 1; }
    cudaMemsetAsync(f.queueCount, 0, sizeof(unsigned int), st);
    k_gbuffer<<<pixelGrid(f), RS_BLOCK, 0, st>>>(s, f, cam, lastCam);
    k_gbuffer_fix<<<RS_FIX_BLOCKS, RS_BLOCK, 0, st>>>(s, f, cam, lastCam);
    return 2;
}
int launchRestirA(const DevScene& s, const FrameDev& f, const CamDev& cam, const RstrParams& p, int looper, int iter, int first, cudaStream_t st) {
    const bool sp = (p.reuse & 2) != 0;
    if (s.traversal == RS_TRAVERSAL_EXACT) {
        if (sp) k_restir_a_exact<true><<<pixelGrid(f), RS_BLOCK, 0, st>>>(s, f, cam, p, looper, iter, first);
        else k_restir_a_exact<false><<<pixelGrid(f), RS_BLOCK, 0, st>>>(s, f, cam, p, looper, iter, first);
        return 1;
    }
    cudaMemsetAsync(f.queueCount, 0, sizeof(unsigned int), st);
    if (sp) {
        k_restir_a<true><<<pixelGrid(f), RS_BLOCK, 0, st>>>(s, f, cam, p, looper, iter, first);
        k_restir_a_fix<true><<<RS_FIX_BLOCKS, RS_BLOCK, 0, st>>>(s, f, cam, p, looper, iter, first);
    } else {
        k_restir_a<false><<<pixelGrid(f), RS_BLOCK, 0, st>>>(s, f, cam, p, looper, iter, first);
        k_restir_a_fix<false><<<RS_FIX_BLOCKS, RS_BLOCK, 0, st>>>(s, f, cam, p, looper, iter, first);
    }
    return 2;
}
// G-buffer + phase A in one launch (traced tree only); returns 0 when the traversal mode has no fused kernel
int launchGBufferRestirA(const DevScene& s, const FrameDev& f, const CamDev& cam, const CamDev& lastCam, const RstrParams& p, int looper, int iter, int first, cudaStream_t st) {
    if (s.traversal == RS_TRAVERSAL_EXACT) return 0;
    cudaMemsetAsync(f.queueCount, 0, sizeof(unsigned int), st);
    if (p.reuse & 2) {
        k_gbuffer_restir_a<true><<<pixelGrid(f), RS_BLOCK, 0, st>>>(s, f, cam, lastCam, p, looper, iter, first);
        k_gbuffer_restir_a_fix<true><<<RS_FIX_BLOCKS, RS_BLOCK, 0, st>>>(s, f, cam, lastCam, p, looper, iter, first);
    } else {
        k_gbuffer_restir_a<false><<<pixelGrid(f), RS_BLOCK, 0, st>>>(s, f, cam, lastCam, p, looper, iter, first);
        k_gbuffer_restir_a_fix<false><<<RS_FIX_BLOCKS, RS_BLOCK, 0, st>>>(s, f, cam, lastCam, p, looper, iter, first);
    }
    return 2;
}
// staged form of the above (k_primary -> k_candidates -> k_shadow -> k_temporal -> fix-up).  The rows are cut into up to
// RS_MAX_BANDS horizontal bands, each with its own queue: band b's k_primary runs on the frame's stream, its three queue
// kernels on one of two side streams, so that they overlap the NEXT band's k_primary -- an issue-bound packet walk next
// to gather-bound kernels, and every kernel's tail hidden under another kernel (measured: profiles/README.md).
int launchPhaseAStaged(const DevScene& s, const FrameDev& f, const CamDev& cam, const CamDev& lastCam, const RstrParams& p, int looper, int iter, int first, int numSMs,
                       cudaStream_t st, const StagedStreams& ss) {
    if (s.traversal == RS_TRAVERSAL_EXACT) return 0;
    const int rows = f.rowHi - f.rowLo;
    int bands = ss.bands < 1 ? 1 : (ss.bands > RS_MAX_BANDS ? RS_MAX_BANDS : ss.bands);
    int bandRows = ((rows + bands - 1) / bands + 7) & ~7;            // whole 8-row tiles
    if (bandRows * (bands - 1) >= rows) bands = (rows + bandRows - 1) / bandRows;
    cudaMemsetAsync(f.queueCount, 0, (4 + 2 * RS_MAX_BANDS) * sizeof(unsigned int), st);
    const bool sp = (p.reuse & 2) != 0;
    int launches = 0;
    for (int b = 0; b < bands; b++) {
        FrameDev fb = f;
        fb.rowLo = f.rowLo + b * bandRows;
        fb.rowHi = fb.rowLo + bandRows < f.rowHi ? fb.rowLo + bandRows : f.rowHi;
        fb.shadeCount = f.queueCount + 4 + 2 * b;
        fb.shadeQueue = f.shadeQueue + (size_t)(fb.rowLo - f.rowLo) * f.W;
        const dim3 grid = pixelGrid(fb);
        const unsigned linear = grid.x * grid.y;           // one thread per pixel of the band: upper bound of its queue length
        if (sp) k_primary<true><<<grid, RS_BLOCK, 0, st>>>(s, fb, cam, lastCam, looper, iter);
        else k_primary<false><<<grid, RS_BLOCK, 0, st>>>(s, fb, cam, lastCam, looper, iter);
        cudaStream_t q = st;
        if (bands > 1) {
            q = ss.side[b & 1];
            cudaEventRecord(ss.evPrimary[b], st);
            cudaStreamWaitEvent(q, ss.evPrimary[b], 0);
        }
        k_candidates<<<linear, RS_BLOCK, 0, q>>>(s, fb, cam, p, looper);
        // the bands' shadow kernels share the machine with each other and with the next band's k_primary
        if (!p.unbiased) {
            const unsigned sg = (unsigned)(numSMs * (bands > 1 ? RS_MINB_SHADOW / 2 : RS_MINB_SHADOW));
            static const int forced = getenv("RSTR_SHADOW_DRAIN") ? atoi(getenv("RSTR_SHADOW_DRAIN")) : -1;      // A/B switch
            const bool drainTail = forced >= 0 ? forced != 0 : (size_t)(fb.rowHi - fb.rowLo) * (size_t)f.W <= (size_t)RS_DRAIN_MAX_PIXELS;
            if (drainTail) k_shadow<true><<<sg, RS_BLOCK, 0, q>>>(s, fb);
            else k_shadow<false><<<sg, RS_BLOCK, 0, q>>>(s, fb);
        }
        if (p.unbiased) {
            if (sp) k_temporal_unb<true><<<linear, RS_BLOCK, 0, q>>>(s, fb, p, iter, first);
            else k_temporal_unb<false><<<linear, RS_BLOCK, 0, q>>>(s, fb, p, iter, first);
        } else if (sp) k_temporal<true><<<linear, RS_BLOCK, 0, q>>>(s, fb, p, iter, first);
        else k_temporal<false><<<linear, RS_BLOCK, 0, q>>>(s, fb, p, iter, first);
        if (bands > 1) cudaEventRecord(ss.evDone[b], q);
        launches += 4;
    }
    if (bands > 1)
        for (int b = 0; b < bands; b++) cudaStreamWaitEvent(st, ss.evDone[b], 0);
    if (sp) k_gbuffer_restir_a_fix<true><<<RS_FIX_BLOCKS, RS_BLOCK, 0, st>>>(s, f, cam, lastCam, p, looper, iter, first);
    else k_gbuffer_restir_a_fix<false><<<RS_FIX_BLOCKS, RS_BLOCK, 0, st>>>(s, f, cam, lastCam, p, looper, iter, first);
    return launches + 1;
}
void launchRestirB(const DevScene& s, const FrameDev& f, const RstrParams& p, int iter, const ResvD* src, ResvD* dst, int pass, int last, cudaStream_t st) {
    if (p.unbiased) { k_restir_b_unb<<<pixelGrid(f), RS_BLOCK, 0, st>>>(s, f, p, iter, src, dst, pass, last); return; }
    k_restir_b<<<pixelGrid(f), 128, 0, st>>>(s, f, p, iter, src, dst, pass, last);
}
int launchPTDirect(const DevScene& s, const FrameDev& f, const CamDev& cam, int looper, int iter, cudaStream_t st) {
    if (s.traversal == RS_TRAVERSAL_EXACT) { k_ptdirect_exact<<<pixelGrid(f), RS_BLOCK, 0, st>>>(s, f, cam, looper, iter); return 1; }
    cudaMemsetAsync(f.queueCount, 0, sizeof(unsigned int), st);
    k_ptdirect<<<pixelGrid(f), RS_BLOCK, 0, st>>>(s, f, cam, looper, iter);
    k_ptdirect_fix<<<RS_FIX_BLOCKS, RS_BLOCK, 0, st>>>(s, f, cam, looper, iter);
    return 2;
}
void launchTonemap(const float* radiance, uchar4* ldr, size_t n, int toneMapping, float scale, cudaStream_t st) {
    k_tonemap<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(radiance, ldr, n, toneMapping, scale);
}
void launchExportGeom(const float4* geom, const float4* am, float* albedo, float* normal, float* depth, int* motion, size_t n, cudaStream_t st) {
    k_export_geom<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(geom, am, albedo, normal, depth, motion, n);
}
void launchExportResv(const DevScene& s, const ResvD* src, float* out36, int* lightIdx, size_t n, cudaStream_t st) {
    k_export_resv<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s, src, out36, lightIdx, n);
}

#endif  // RS_HOST_EMU

// ------------------------------------------------------------------------------------------------ ReSTIR GI
#include "gi_kernels.inl"

}  // namespace rs
