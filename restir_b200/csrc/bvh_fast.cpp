// bvh_fast.cpp -- the acceleration structure the B200 kernels actually walk.
//
// The reference's tree (bvh.cpp:10-131) has to be REPRODUCED bit for bit (it is part of the host-facing surface and
// fixes the tie-break order of coincident hits), but it is a poor structure to trace: its "SAH" compares single
// neighbouring buckets (bvh.cpp:96-106), it stores one triangle per leaf and is 54-63 levels deep on the benchmark
// scenes (~160 box tests per primary ray).  Hit results do not depend on the tree, only on
//   (a) the triangle test arithmetic (intersections.h:17-53, kept exactly), and
//   (b) which of several hits at (nearly) the same distance is kept: that follows the reference's visiting order and
//       box-distance pruning (scene.h:260,267) -> such rays are detected during the walk and re-traced in reference order.
// So the kernels trace a separately built binned-SAH BVH2 (all three axes, 32 bins, <= 4 triangles per leaf, boxes
// padded by a few ulps so that the FMA slab test stays conservative w.r.t. the exact triangle test) and fall back to
// the reference-order walk only for the rays whose reference box test is not a conservative slab test (|d_a| > 1-1e-6,
// bvh.h:91-123).
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <limits>

#include "scene_host.h"

namespace rs {

namespace {

struct FBox {
    float lo[3], hi[3];
    void reset() { lo[0] = lo[1] = lo[2] = FLT_MAX; hi[0] = hi[1] = hi[2] = -FLT_MAX; }
    void grow(const FBox& b) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    void grow(const float* p) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    float halfArea() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return dx * dy + dy * dz + dz * dx;
    }
};

struct Ref { FBox b; float c[3]; int prim; };

struct FastBuilder {
    static const int NBINS = 32;
    int MAX_LEAF = 4;             // RSTR_MAX_LEAF (1..8) / RSTR_LEAF_CI: development knobs for tree experiments (scripts/travsim.cpp)
    float CI = 1.0f;
    static const int TASK_MIN = 4096;
    std::vector<Ref> refs;
    std::vector<FastNode>& nodes;
    std::vector<int>& order;
    std::atomic<int> maxDepth{0};

    FastBuilder(std::vector<FastNode>& n, std::vector<int>& o) : nodes(n), order(o) {
        if (const char* e = getenv("RSTR_MAX_LEAF")) MAX_LEAF = std::min(8, std::max(1, atoi(e)));
        if (const char* e = getenv("RSTR_LEAF_CI")) CI = (float)atof(e);
    }

    static int leafRef(int first, int count) { return (int)(0x80000000u | ((unsigned)(count - 1) << 27) | (unsigned)first); }

    // builds the subtree over refs[start,end) and returns the child reference; the subtree's box goes to `box`
    int build(int start, int end, int depth, FBox& box) {
        const int n = end - start;
        FBox cb;
        box.reset(); cb.reset();
        for (int i = start; i < end; i++) { box.grow(refs[i].b); cb.grow(refs[i].c); }
        int d = maxDepth.load(std::memory_order_relaxed);
        while (depth > d && !maxDepth.compare_exchange_weak(d, depth)) {}
        auto makeLeaf = [&]() {
            for (int i = start; i < end; i++) order[i] = refs[i].prim;
            return leafRef(start, n);
        };
        if (n == 1) return makeLeaf();
        // binned SAH over the three axes
        float bestCost = FLT_MAX;
        int bestAxis = -1, bestBin = -1;
        const float parentArea = box.halfArea();
        for (int axis = 0; axis < 3; axis++) {
            float lo = cb.lo[axis], ext = cb.hi[axis] - cb.lo[axis];
            if (!(ext > 0.f)) continue;
            FBox bb[NBINS];
            int cnt[NBINS];
            for (int b = 0; b < NBINS; b++) { bb[b].reset(); cnt[b] = 0; }
            const float scale = NBINS / ext;
            for (int i = start; i < end; i++) {
                int b = std::min(NBINS - 1, std::max(0, (int)((refs[i].c[axis] - lo) * scale)));
                bb[b].grow(refs[i].b); cnt[b]++;
            }
            float rightArea[NBINS];
            int rightCnt[NBINS];
            FBox acc; acc.reset();
            int c = 0;
            for (int b = NBINS - 1; b > 0; b--) { acc.grow(bb[b]); c += cnt[b]; rightArea[b] = acc.halfArea(); rightCnt[b] = c; }
            acc.reset(); c = 0;
            for (int b = 0; b < NBINS - 1; b++) {
                acc.grow(bb[b]); c += cnt[b];
                if (c == 0 || rightCnt[b + 1] == 0) continue;
                float cost = acc.halfArea() * c + rightArea[b + 1] * rightCnt[b + 1];
                if (cost < bestCost) { bestCost = cost; bestAxis = axis; bestBin = b; }
            }
        }
        // leaf when small and not worth splitting (traversal step ~ 1 triangle test)
        if (n <= MAX_LEAF) {
            float leafCost = (float)n * parentArea;
            float splitCost = bestAxis < 0 ? FLT_MAX : CI * parentArea + bestCost;
            if (leafCost <= splitCost) return makeLeaf();
        }
        int mid;
        if (bestAxis < 0) {                       // coincident centroids: split the range in the middle
            mid = start + n / 2;
        } else {
            const float lo = cb.lo[bestAxis], scale = NBINS / (cb.hi[bestAxis] - cb.lo[bestAxis]);
            auto it = std::partition(refs.begin() + start, refs.begin() + end, [&](const Ref& r) {
                int b = std::min(NBINS - 1, std::max(0, (int)((r.c[bestAxis] - lo) * scale)));
                return b <= bestBin;
            });
            mid = (int)(it - refs.begin());
            if (mid == start || mid == end) mid = start + n / 2;
        }
        const int me = mid - 1;                   // every internal node has a distinct split position: deterministic, subtree-local index
        FBox lb, rb;
        int l, r;
        if (n > TASK_MIN) {
#pragma omp task default(shared)
            l = build(start, mid, depth + 1, lb);
            r = build(mid, end, depth + 1, rb);
#pragma omp taskwait
        } else {
            l = build(start, mid, depth + 1, lb);
            r = build(mid, end, depth + 1, rb);
        }
        FastNode& nd = nodes[me];
        // pad: keeps the FMA slab test conservative w.r.t. the exact Moller-Trumbore arithmetic (flat boxes!)
        auto put = [&](const FBox& b, float* mn, float* mx) {
            for (int a = 0; a < 3; a++) {
                float pad = 4e-6f * std::max(fabsf(b.lo[a]), fabsf(b.hi[a])) + 1e-7f;
                mn[a] = b.lo[a] - pad; mx[a] = b.hi[a] + pad;
            }
        };
        put(lb, nd.lmin, nd.lmax);
        put(rb, nd.rmin, nd.rmax);
        nd.left = l; nd.right = r; nd.pad[0] = nd.pad[1] = 0;
        return me;
    }
};

// BVH2 -> BVH4: every node adopts its grandchildren, largest surface area first, until it has four children (or only
// leaves are left).  Halves the number of dependent node fetches per ray, which is what the traversal kernels wait on.
void collapse4(HostScene& hs) {
    hs.fastNodes4.clear();
    hs.fastDepth4 = 0;
    hs.fastRoot4 = hs.fastRoot;
    if (hs.fastRoot < 0) return;                       // the whole scene is one leaf
    struct Entry { int ref; float lo[3], hi[3]; };
    struct Work { int node2, out, depth; };
    std::vector<Work> stack;
    hs.fastNodes4.reserve(hs.fastNodes.size() / 2 + 1);
    hs.fastNodes4.push_back(FastNode4{});
    stack.push_back(Work{hs.fastRoot, 0, 1});
    auto children = [&](int n2, Entry& a, Entry& b) {
        const FastNode& nd = hs.fastNodes[n2];
        a.ref = nd.left; memcpy(a.lo, nd.lmin, 12); memcpy(a.hi, nd.lmax, 12);
        b.ref = nd.right; memcpy(b.lo, nd.rmin, 12); memcpy(b.hi, nd.rmax, 12);
    };
    auto area = [](const Entry& e) {
        float dx = e.hi[0] - e.lo[0], dy = e.hi[1] - e.lo[1], dz = e.hi[2] - e.lo[2];
        return dx * dy + dy * dz + dz * dx;
    };
    while (!stack.empty()) {
        Work w = stack.back(); stack.pop_back();
        hs.fastDepth4 = std::max(hs.fastDepth4, w.depth);
        Entry e[4];
        int n = 2;
        children(w.node2, e[0], e[1]);
        while (n < 4) {
            int pick = -1;
            for (int i = 0; i < n; i++)
                if (e[i].ref >= 0 && (pick < 0 || area(e[i]) > area(e[pick]))) pick = i;
            if (pick < 0) break;
            int n2 = e[pick].ref;
            children(n2, e[pick], e[n]);
            n++;
        }
        FastNode4 nd;
        const float inf = std::numeric_limits<float>::infinity();
        for (int i = 0; i < 4; i++) {
            if (i < n) {
                nd.lox[i] = e[i].lo[0]; nd.loy[i] = e[i].lo[1]; nd.loz[i] = e[i].lo[2];
                nd.hix[i] = e[i].hi[0]; nd.hiy[i] = e[i].hi[1]; nd.hiz[i] = e[i].hi[2];
                if (e[i].ref >= 0) {
                    nd.child[i] = (int)hs.fastNodes4.size();
                    hs.fastNodes4.push_back(FastNode4{});
                    stack.push_back(Work{e[i].ref, nd.child[i], w.depth + 1});
                } else nd.child[i] = e[i].ref;
            } else {
                nd.lox[i] = nd.loy[i] = nd.loz[i] = nd.hix[i] = nd.hiy[i] = nd.hiz[i] = inf;        // unused slot: child == 0x7fffffff
                nd.child[i] = 0x7fffffff;
            }
            nd.pad[i] = 0;
        }
        hs.fastNodes4[w.out] = nd;
    }
    hs.fastRoot4 = 0;
}

}  // namespace

void buildFastBVH(HostScene& hs) {
    auto t0 = std::chrono::steady_clock::now();
    const int T = hs.T;
    hs.fastNodes.assign(T > 1 ? T - 1 : 1, FastNode{});
    hs.fastOrder.assign(T, 0);
    FastBuilder b(hs.fastNodes, hs.fastOrder);
    b.refs.resize(T);
    for (int i = 0; i < T; i++) {
        Ref& r = b.refs[i];
        r.b.reset();
        r.b.grow(&hs.vertices[3 * i].x); r.b.grow(&hs.vertices[3 * i + 1].x); r.b.grow(&hs.vertices[3 * i + 2].x);
        for (int a = 0; a < 3; a++) r.c[a] = 0.5f * (r.b.lo[a] + r.b.hi[a]);
        r.prim = i;
    }
    FBox root;
    int rootRef;
#pragma omp parallel
#pragma omp single
    rootRef = b.build(0, T, 1, root);
    hs.fastRoot = rootRef;
    hs.fastDepth = b.maxDepth.load();
    if (RS_BVH4) collapse4(hs);
    for (int a = 0; a < 3; a++) {
        float pad = 4e-6f * std::max(fabsf(root.lo[a]), fabsf(root.hi[a])) + 1e-7f;
        hs.fastRootMin[a] = root.lo[a] - pad; hs.fastRootMax[a] = root.hi[a] + pad;
    }
    // triangles in leaf order; each record keeps its original primitive id
    hs.fastTris.resize(T);
    hs.primToFast.assign(T, 0);
    for (int i = 0; i < T; i++) {
        int p = hs.fastOrder[i];
        TriGeom& g = hs.fastTris[i];
        memcpy(g.v0, &hs.vertices[3 * p], 36);
        g.matId = hs.materialIds[p];
        g.pad[0] = p; g.pad[1] = 0;
        hs.primToFast[p] = i;
    }
    // visiting rank of every triangle's leaf in each of the reference's 6 orderings (bvh.cpp:156-193)
    hs.rank.assign(6 * (size_t)T, 0);
#pragma omp parallel for schedule(dynamic, 1)
    for (int o = 0; o < 6; o++) {
        std::vector<MTNode> mt;
        exportMTBVH(hs, o, mt);
        int* rk = hs.rank.data() + (size_t)o * T;
        for (int i = 0; i < (int)mt.size(); i++)
            if (mt[i].prim >= 0) rk[mt[i].prim] = i;
    }
    hs.fastBuildSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace rs
