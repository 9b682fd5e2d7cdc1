// bvh_gpu.cu -- device-side build of the TRACED tree ("GPU BVH build / refit replacing the 1.5 s CPU
// builder for dynamic scenes, keeping the bit-exact CPU path as oracle").
//
// What a ray reports does not depend on the tree it walks (the reference's hit is decided per triangle,
// the reference-identical tree of bvh.cpp stays on the host for rstr_scene_read and for the rank table), so the tree the
// kernels trace may be rebuilt on the device at any time.  Two builders over the same front end (63-bit Morton codes of the
// triangle centroids, 21 bits per axis, one radix sort):
//   mode 0  PLOC (parallel locally-ordered clustering, Meister & Bittner 2018): the Morton-ordered clusters are merged bottom-up,
//           every round each cluster picks the neighbour within RS_PLOC_RADIUS positions whose union with it has the smallest
//           surface area and mutual pairs merge; the SAH cost of every subtree is known when it is formed, so the cheapest cut
//           into leaves of <= 4 triangles (the rule of bvh_fast.cpp, evaluated bottom-up) falls out of the same pass;
//   mode 1  the binary radix tree over the sorted codes (Karras 2012), same leaf rule: the fastest build (0.5 ms for 1M
//           triangles), a markedly slower tree to trace.
// Both end in the kernels' 64-byte FastNode records (both children's padded boxes + links) and the triangle records re-ordered
// into leaf order.  cub's radix sort / scan are library code on a non-hot path (scene (re)build).
#ifndef RS_HOST_EMU      /* tests/emu compiles the kernels below with g++ and drives them itself (std::stable_sort / a serial scan for cub's) */
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#endif

#include "capi_internal.h"

using namespace rs;

namespace {

#ifndef RS_PLOC_RADIUS
#define RS_PLOC_RADIUS 16      /* search window of the clustering, in sorted positions on each side */
#endif
#define RS_GPU_MAX_LEAF 4

struct GBox { float lo[3], hi[3]; };

__device__ __forceinline__ unsigned long long expandBits(unsigned int v) {      // 21 bits -> every third bit
    unsigned long long x = v & 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

// (RS_PEEK, device_types.h: the first read of the CAS loops below is a plain load racing with other blocks' CAS)
__device__ __forceinline__ void atomicMinF(float* a, float v) {
    int* ai = (int*)a;
    int old = RS_PEEK(ai);
    while (v < __int_as_float(old)) { int assumed = old; old = atomicCAS(ai, assumed, __float_as_int(v)); if (old == assumed) break; }
}
__device__ __forceinline__ void atomicMaxF(float* a, float v) {
    int* ai = (int*)a;
    int old = RS_PEEK(ai);
    while (v > __int_as_float(old)) { int assumed = old; old = atomicCAS(ai, assumed, __float_as_int(v)); if (old == assumed) break; }
}

// triangle records: 3 x float4 {v0.xyz v1.x}{v1.yz v2.xy}{v2.z, matId, prim, -}
__device__ __forceinline__ void triVerts(const float4* tg, int i, float v[9]) {
    float4 a = tg[3 * (size_t)i], b = tg[3 * (size_t)i + 1], c = tg[3 * (size_t)i + 2];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; v[8] = c.x;
}
__device__ __forceinline__ GBox triBox(const float4* tg, int i) {
    float v[9];
    triVerts(tg, i, v);
    GBox b;
    for (int a = 0; a < 3; a++) { b.lo[a] = fminf(fminf(v[a], v[3 + a]), v[6 + a]); b.hi[a] = fmaxf(fmaxf(v[a], v[3 + a]), v[6 + a]); }
    return b;
}
__device__ __forceinline__ GBox unite(const GBox& l, const GBox& r) {
    GBox b;
    for (int a = 0; a < 3; a++) { b.lo[a] = fminf(l.lo[a], r.lo[a]); b.hi[a] = fmaxf(l.hi[a], r.hi[a]); }
    return b;
}
__device__ __forceinline__ float halfArea(const GBox& b) {
    const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    return dx * dy + dy * dz + dz * dx;
}

__global__ void k_bvh_scene_bounds(const float4* tg, int T, float* bounds /* lo[3], hi[3] of the centroids */) {
    __shared__ float slo[3][256], shi[3][256];
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T; i += gridDim.x * blockDim.x) {
        float v[9];
        triVerts(tg, i, v);
        for (int a = 0; a < 3; a++) {
            float c = (v[a] + v[3 + a] + v[6 + a]) * (1.f / 3.f);
            lo[a] = fminf(lo[a], c); hi[a] = fmaxf(hi[a], c);
        }
    }
    for (int a = 0; a < 3; a++) { slo[a][threadIdx.x] = lo[a]; shi[a][threadIdx.x] = hi[a]; }
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s)
            for (int a = 0; a < 3; a++) {
                slo[a][threadIdx.x] = fminf(slo[a][threadIdx.x], slo[a][threadIdx.x + s]);
                shi[a][threadIdx.x] = fmaxf(shi[a][threadIdx.x], shi[a][threadIdx.x + s]);
            }
        __syncthreads();
    }
    if (threadIdx.x == 0)
        for (int a = 0; a < 3; a++) { atomicMinF(bounds + a, slo[a][0]); atomicMaxF(bounds + 3 + a, shi[a][0]); }
}

__global__ void k_bvh_morton(const float4* tg, int T, const float* bounds, unsigned long long* keys, unsigned int* slots) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    float v[9];
    triVerts(tg, i, v);
    unsigned long long code = 0;
    for (int a = 0; a < 3; a++) {
        float c = (v[a] + v[3 + a] + v[6 + a]) * (1.f / 3.f);
        float ext = bounds[3 + a] - bounds[a];
        float u = ext > 0.f ? (c - bounds[a]) / ext : 0.f;
        unsigned int q = (unsigned int)fminf(fmaxf(u * 2097152.f, 0.f), 2097151.f);
        code |= expandBits(q) << (2 - a);
    }
    keys[i] = code;
    slots[i] = (unsigned int)i;
}

// Node ids of both builders: 0 .. T-1 are the triangles in Morton order, T .. 2T-2 the internal nodes.
// SAH cost of a subtree in its cheapest form: ONE leaf when it holds <= 4 triangles and n * area is no more than a node
// step plus the children's costs (bvh_fast.cpp's rule, here decided bottom-up where both alternatives are known).
// Stored with the sign bit set when the leaf form won (a comparison of recomputed floats would depend on contraction).
__device__ __forceinline__ float subtreeCost(int n, float area, float costL, float costR) {
    const float split = area + fabsf(costL) + fabsf(costR), leaf = (float)n * area;
    return (n <= RS_GPU_MAX_LEAF && leaf <= split) ? -leaf : split;
}
__device__ __forceinline__ bool leafForm(float cost) { return __float_as_int(cost) < 0; }

// ---------------------------------------------------------------------------------------------- mode 1: binary radix tree
// Karras 2012: internal node i of the binary radix tree over the sorted keys (equal codes: the sorted position tells them apart)
__device__ __forceinline__ int delta(const unsigned long long* keys, int T, int i, int j) {
    if (j < 0 || j >= T) return -1;
    const unsigned long long x = keys[i] ^ keys[j];
    return x ? __clzll(x) : 64 + __clz(i ^ j);
}
__global__ void k_bvh_radix_tree(const unsigned long long* keys, int T, int* left, int* right, int* parent) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T - 1) return;
    int d = (delta(keys, T, i, i + 1) - delta(keys, T, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, T, i, i - d);
    int lmax = 2;
    while (delta(keys, T, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, T, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, T, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (delta(keys, T, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    int gamma = i + s * d + min(d, 0);
    int first = min(i, j), last = max(i, j);
    const int lc = (first == gamma) ? gamma : T + gamma;              // a triangle (Morton position) or internal node T + index
    const int rc = (last == gamma + 1) ? gamma + 1 : T + gamma + 1;
    left[T + i] = lc; right[T + i] = rc;
    parent[lc] = T + i; parent[rc] = T + i;
    if (i == 0) parent[T] = -1;
}
// boxes, triangle counts and SAH costs bottom-up: one thread per triangle; the second child to arrive at a node finishes it
__global__ void k_bvh_bottom_up(const float4* tg, const unsigned int* slots, int T, const int* left, const int* right, const int* parent,
                                GBox* box, int* count, float* cost, unsigned int* arrived) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    GBox b = triBox(tg, (int)slots[i]);
    box[i] = b; count[i] = 1; cost[i] = -halfArea(b);
    __threadfence();
    int node = parent[i];
    while (node >= 0) {
        if (atomicAdd(arrived + (node - T), 1u) == 0u) return;        // the sibling subtree is not finished yet
        __threadfence();
        const int l = left[node], r = right[node];
        b = unite(box[l], box[r]);
        const int n = count[l] + count[r];
        box[node] = b; count[node] = n;
        cost[node] = subtreeCost(n, halfArea(b), cost[l], cost[r]);
        __threadfence();
        node = parent[node];
    }
}

// ---------------------------------------------------------------------------------------------- mode 0: PLOC
__global__ void k_ploc_init(const float4* tg, const unsigned int* slots, int T, GBox* box, int* count, float* cost, int* parent, int* cid) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    const GBox b = triBox(tg, (int)slots[i]);
    box[i] = b; count[i] = 1; cost[i] = -halfArea(b); parent[i] = -1; cid[i] = i;
}
// every cluster's best partner within the window: smallest surface area of the union, ties to the lower position
__global__ void __launch_bounds__(256) k_ploc_nearest(const int* cid, int n, const GBox* box, int* nearest) {
    __shared__ GBox tile[256 + 2 * RS_PLOC_RADIUS];
    const int base = blockIdx.x * 256 - RS_PLOC_RADIUS;
    for (int t = threadIdx.x; t < 256 + 2 * RS_PLOC_RADIUS; t += 256) {
        const int j = base + t;
        if (j >= 0 && j < n) tile[t] = box[cid[j]];
    }
    __syncthreads();
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const GBox me = tile[threadIdx.x + RS_PLOC_RADIUS];
    float best = FLT_MAX;
    int bestJ = -1;
    for (int k = -RS_PLOC_RADIUS; k <= RS_PLOC_RADIUS; k++) {
        const int j = i + k;
        if (k == 0 || j < 0 || j >= n) continue;
        const float a = halfArea(unite(me, tile[threadIdx.x + RS_PLOC_RADIUS + k]));
        if (a < best) { best = a; bestJ = j; }
    }
    nearest[i] = bestJ;
}
// mutual pairs merge: the lower position creates the node and keeps the slot, the upper one leaves the list
__global__ void k_ploc_merge(const int* cid, int n, const int* nearest, int T, GBox* box, int* count, float* cost, int* left, int* right, int* parent,
                             int* counter, int* cidTmp, int* keep) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int j = nearest[i];
    int me = cid[i], flag = 1;
    if (j >= 0 && nearest[j] == i) {
        if (i < j) {
            const int l = me, r = cid[j];
            const int v = T + atomicAdd(counter, 1);
            const GBox b = unite(box[l], box[r]);
            const int c = count[l] + count[r];
            box[v] = b; count[v] = c; left[v] = l; right[v] = r; parent[v] = -1; parent[l] = v; parent[r] = v;
            cost[v] = subtreeCost(c, halfArea(b), cost[l], cost[r]);
            me = v;
        } else flag = 0;
    }
    cidTmp[i] = me; keep[i] = flag;
}
__global__ void k_ploc_compact(const int* cidTmp, const int* keep, const int* pos, int n, int* cidOut, int* nOut) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (keep[i]) cidOut[pos[i]] = cidTmp[i];
    if (i == n - 1) *nOut = pos[i] + keep[i];
}

// ---------------------------------------------------------------------------------------------- common back end
__device__ __forceinline__ int leafRefOf(int first, int count) { return (int)(0x80000000u | ((unsigned)(count - 1) << 27) | (unsigned)first); }
__device__ __forceinline__ void padBox(const GBox& b, float* mn, float* mx) {          // as bvh_fast.cpp: keeps the FMA slab test conservative
    for (int a = 0; a < 3; a++) {
        float pad = 4e-6f * fmaxf(fabsf(b.lo[a]), fabsf(b.hi[a])) + 1e-7f;
        mn[a] = b.lo[a] - pad; mx[a] = b.hi[a] + pad;
    }
}
__device__ __forceinline__ bool asLeaf(int v, int T, const float* cost) { return v < T || leafForm(cost[v]); }
// position of a node's first triangle in the depth-first (leaf) order: the triangles of all left siblings on the way to the root
__global__ void k_bvh_first(int T, int numNodes, const int* left, const int* right, const int* parent, const int* count, int* firstPos, int* maxDepth) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= numNodes) return;
    int off = 0, depth = 0;
    for (int c = v, p = parent[v]; p >= 0; c = p, p = parent[p]) {
        if (right[p] == c) off += count[left[p]];
        depth++;
    }
    firstPos[v] = off;
    if (v >= T) atomicMax(maxDepth, depth + 1);
}
__global__ void k_bvh_emit(int T, const int* left, const int* right, const GBox* box, const int* count, const float* cost, const int* firstPos, FastNode* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T - 1) return;
    const int v = T + i;
    FastNode nd;
    memset(&nd, 0, sizeof nd);
    if (!asLeaf(v, T, cost)) {                             // (nodes inside a leaf are never referenced)
        const int c[2] = {left[v], right[v]};
        int ref[2];
        for (int k = 0; k < 2; k++) ref[k] = asLeaf(c[k], T, cost) ? leafRefOf(firstPos[c[k]], count[c[k]]) : c[k] - T;
        padBox(box[c[0]], nd.lmin, nd.lmax);
        padBox(box[c[1]], nd.rmin, nd.rmax);
        nd.left = ref[0]; nd.right = ref[1];
    }
    out[i] = nd;
}
__global__ void k_bvh_reorder(const float4* tgOld, const unsigned int* slots, const int* firstPos, int T, float4* tgNew, int* primToFast) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    const int src = (int)slots[i], dst = firstPos[i];
    const float4 a = tgOld[3 * (size_t)src], b = tgOld[3 * (size_t)src + 1], c = tgOld[3 * (size_t)src + 2];
    tgNew[3 * (size_t)dst] = a; tgNew[3 * (size_t)dst + 1] = b; tgNew[3 * (size_t)dst + 2] = c;
    primToFast[__float_as_int(c.z)] = dst;
}

#ifndef RS_HOST_EMU
struct DevBuf {                       // frees what it allocated (and still owns) when it goes out of scope
    std::vector<void*> all;
    cudaError_t err = cudaSuccess;
    template <typename T_> T_* get(size_t n) {
        void* p = nullptr;
        if (err == cudaSuccess) err = cudaMalloc(&p, n * sizeof(T_));
        if (p) all.push_back(p);
        return (T_*)p;
    }
    void release(void* p) { for (void*& q : all) if (q == p) q = nullptr; }
    ~DevBuf() { for (void* p : all) cudaFree(p); }
};
#endif

}  // namespace

#ifndef RS_HOST_EMU
extern "C" {

// Rebuild the traced tree of an uploaded scene on the device.  *milliseconds (may be NULL) receives the device time of the build.
int rstr_scene_build_traced_gpu(RstrScene* sc, int mode, float* milliseconds) {
    if (!sc) return rsFail(RSTR_ERR_ARG, "rstr_scene_build_traced_gpu: null scene");
    if (mode != 0 && mode != 1) return rsFail(RSTR_ERR_ARG, "rstr_scene_build_traced_gpu: mode must be 0 (PLOC) or 1 (radix tree)");
    { int rc = rsEnsureUploaded(sc); if (rc != RSTR_OK) return rc; }
    const int T = sc->hs.T;
    if (T < 8) return rsFail(RSTR_ERR_ARG, "rstr_scene_build_traced_gpu: needs at least 8 triangles");
    CU(cudaDeviceSynchronize());                      // no frame may be tracing the old tree
    const size_t N2 = 2 * (size_t)T - 1;
    DevBuf m;
    unsigned long long *keys = m.get<unsigned long long>(T), *keysSorted = m.get<unsigned long long>(T);
    unsigned int *slots = m.get<unsigned int>(T), *slotsSorted = m.get<unsigned int>(T);
    int *left = m.get<int>(N2), *right = m.get<int>(N2), *parent = m.get<int>(N2), *count = m.get<int>(N2), *firstPos = m.get<int>(N2);
    float* cost = m.get<float>(N2);
    GBox* box = m.get<GBox>(N2);
    unsigned int* arrived = m.get<unsigned int>(T);
    int *cidA = m.get<int>(T), *cidB = m.get<int>(T), *cidTmp = m.get<int>(T), *keep = m.get<int>(T), *pos = m.get<int>(T), *nearest = m.get<int>(T);
    int* scalars = m.get<int>(4);                     // [0] node counter, [1] cluster count, [2] max depth
    float* bounds = m.get<float>(6);
    FastNode* nodes = m.get<FastNode>(T - 1);
    float4* tgNew = m.get<float4>(3 * (size_t)T);
    int* primToFast = m.get<int>(T);
    size_t sortBytes = 0, scanBytes = 0;
    if (m.err == cudaSuccess) m.err = cub::DeviceRadixSort::SortPairs(nullptr, sortBytes, keys, keysSorted, slots, slotsSorted, T, 0, 63);
    if (m.err == cudaSuccess) m.err = cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, keep, pos, T);
    const size_t tmpBytes = sortBytes > scanBytes ? sortBytes : scanBytes;
    void* tmp = m.get<char>(tmpBytes);
    if (m.err != cudaSuccess) return rsFail(RSTR_ERR_CUDA, std::string("rstr_scene_build_traced_gpu: ") + cudaGetErrorString(m.err));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    const float4* tg = (const float4*)sc->dTriGeom;
    const float init[6] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
    cudaMemcpy(bounds, init, sizeof init, cudaMemcpyHostToDevice);
    cudaEventRecord(e0, 0);
    const int B = 256, G = (T + B - 1) / B;
    cudaMemsetAsync(scalars, 0, 4 * sizeof(int), 0);
    k_bvh_scene_bounds<<<std::min(G, 1184), B>>>(tg, T, bounds);
    k_bvh_morton<<<G, B>>>(tg, T, bounds, keys, slots);
    size_t tb = tmpBytes;
    cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keysSorted, slots, slotsSorted, T, 0, 63);
    int root = T;
    cudaError_t e = cudaSuccess;
    if (mode == 1) {
        cudaMemsetAsync(arrived, 0, sizeof(unsigned int) * T, 0);
        k_bvh_radix_tree<<<G, B>>>(keysSorted, T, left, right, parent);
        k_bvh_bottom_up<<<G, B>>>(tg, slotsSorted, T, left, right, parent, box, count, cost, arrived);
    } else {
        k_ploc_init<<<G, B>>>(tg, slotsSorted, T, box, count, cost, parent, cidA);
        int n = T;
        int *cur = cidA, *nxt = cidB;
        while (n > 1) {
            const int g = (n + B - 1) / B;
            k_ploc_nearest<<<g, 256>>>(cur, n, box, nearest);
            k_ploc_merge<<<g, B>>>(cur, n, nearest, T, box, count, cost, left, right, parent, scalars, cidTmp, keep);
            tb = tmpBytes;
            cub::DeviceScan::ExclusiveSum(tmp, tb, keep, pos, n);
            k_ploc_compact<<<g, B>>>(cidTmp, keep, pos, n, nxt, scalars + 1);
            int n2 = n;
            e = cudaMemcpy(&n2, scalars + 1, sizeof n2, cudaMemcpyDeviceToHost);
            if (e == cudaSuccess && n2 >= n) e = cudaErrorUnknown;     // cannot happen: the pair with the globally smallest union is always mutual
            if (e != cudaSuccess) break;
            n = n2;
            std::swap(cur, nxt);
        }
        if (e == cudaSuccess) e = cudaMemcpy(&root, cur, sizeof root, cudaMemcpyDeviceToHost);
    }
    if (e == cudaSuccess) {
        const int G2 = (int)((N2 + B - 1) / B);
        k_bvh_first<<<G2, B>>>(T, (int)N2, left, right, parent, count, firstPos, scalars + 2);
        k_bvh_emit<<<G, B>>>(T, left, right, box, count, cost, firstPos, nodes);
        k_bvh_reorder<<<G, B>>>(tg, slotsSorted, firstPos, T, tgNew, primToFast);
    }
    cudaEventRecord(e1, 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();
    int depth = 0;
    GBox rootBox;
    if (e == cudaSuccess) e = cudaMemcpy(&depth, scalars + 2, sizeof depth, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(&rootBox, box + root, sizeof rootBox, cudaMemcpyDeviceToHost);
    float ms = 0.f;
    if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (e != cudaSuccess) return rsFail(RSTR_ERR_CUDA, std::string("rstr_scene_build_traced_gpu: ") + cudaGetErrorString(e));
    if (depth + 1 > RS_PACKET_STACK) return rsFail(RSTR_ERR_LIMIT, "rstr_scene_build_traced_gpu: tree deeper than the traversal stacks; the previous tree stays in place");
    // swap the new tree in
    cudaFree(sc->dFastNodes); cudaFree(sc->dTriGeom); cudaFree(sc->dPrimToFast);
    m.release(nodes); m.release(tgNew); m.release(primToFast);
    sc->dFastNodes = nodes; sc->dTriGeom = tgNew; sc->dPrimToFast = primToFast;
    DevScene& d = sc->dev;
    d.fastNodes = (const float4*)nodes; d.triGeom = (const float4*)tgNew; d.primToFast = primToFast;
    d.numFastNodes = T - 1; d.fastRoot = root - T;
    for (int a = 0; a < 3; a++) {
        float pad = 4e-6f * fmaxf(fabsf(rootBox.lo[a]), fabsf(rootBox.hi[a])) + 1e-7f;
        d.fastRootMin[a] = rootBox.lo[a] - pad; d.fastRootMax[a] = rootBox.hi[a] + pad;
    }
    sc->gpuTreeDepth = depth; sc->gpuTreeMs = ms;
    if (milliseconds) *milliseconds = ms;
    return RSTR_OK;
}

}  // extern "C"
#endif  // RS_HOST_EMU
