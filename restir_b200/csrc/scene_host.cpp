// scene_host.cpp -- host scene build for the B200 ReSTIR DI pipeline.
//
// Produces (a) the reference-format arrays (boundingBoxes, node info, light list, alias table) with
// bit-identical content and ordering to the reference's host code, and (b) the packed device layouts
// derived from them.  The BVH is built by recursive task-parallel partitioning: a node's split depends
// only on the content of its own primitive range, so building sibling subtrees concurrently yields the
// same tree as the reference's explicit-stack loop (bvh.cpp:37-127).
#include "scene_host.h"

#include <limits.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>

namespace rs {

namespace {

inline Box emptyBox() { Box b; b.pMin = mk3(FLT_MAX); b.pMax = mk3(-FLT_MAX); return b; }    // bvh.h:159-160
inline Box grow(const Box& b, f3 p) { Box r; r.pMin = gmin(b.pMin, p); r.pMax = gmax(b.pMax, p); return r; }            // bvh.h:26
inline Box grow(const Box& b, const Box& o) { Box r; r.pMin = gmin(b.pMin, o.pMin); r.pMax = gmax(b.pMax, o.pMax); return r; }  // bvh.h:30
inline f3 center(const Box& b) { return (b.pMin + b.pMax) * .5f; }                              // bvh.h:47
inline float surfaceArea(const Box& b) {                                                        // bvh.h:51
    f3 s = b.pMax - b.pMin;
    return 2.f * (s.x * s.y + s.y * s.z + s.z * s.x);
}
inline int longestAxis(const Box& b) {                                                          // bvh.h:59
    f3 s = b.pMax - b.pMin;
    if (s.x < s.y) return s.y > s.z ? 1 : 2;
    return s.x > s.z ? 0 : 2;
}
// the reference's builder runs on an x86 host: int(float) is cvttss2si (NaN / overflow -> INT_MIN)
inline int f2i_host(float f) {
    if (!(f > -2147483904.f && f < 2147483648.f)) return INT_MIN;
    return (int)f;
}

struct Prim {
    int id;
    Box b;
    f3 c;
};

struct Builder {
    HostScene& hs;
    std::vector<Prim> prim;
    std::atomic<int> maxDepth{0};
    static const int NB = 16;            // bvh.cpp:34
    static const int TASK_MIN = 8192;    // ranges larger than this spawn an OpenMP task for the left subtree

    explicit Builder(HostScene& s) : hs(s) {}

    // one node: bounds, binned split (bvh.cpp:44-126), then both subtrees; finally the packed record
    void node(int offset, int start, int end, int depth) {
        const int n = end - start + 1;
        Box nodeBound = emptyBox(), centerBound = emptyBox();
        for (int i = start; i <= end; i++) {
            nodeBound = grow(nodeBound, prim[i].b);
            centerBound = grow(centerBound, prim[i].c);
        }
        hs.boxes[offset] = nodeBound;
        int d = maxDepth.load(std::memory_order_relaxed);
        while (depth > d && !maxDepth.compare_exchange_weak(d, depth)) {}
        if (n == 1) {
            hs.nodeInfo[offset] = -prim[start].id - 1;
            return;
        }
        hs.nodeInfo[offset] = 2 * n - 1;

        const int axis = longestAxis(centerBound);
        const float dimMin = comp(centerBound.pMin, axis), dimMax = comp(centerBound.pMax, axis);
        auto bucketOf = [&](const Prim& p) {
            int b = f2i_host((comp(p.c, axis) - dimMin) / (dimMax - dimMin) * NB);
            return b < 0 ? 0 : (b > NB - 1 ? NB - 1 : b);
        };
        Box bucketBounds[NB];
        int bucketCounts[NB];
        for (int i = 0; i < NB; i++) { bucketBounds[i] = emptyBox(); bucketCounts[i] = 0; }
        std::vector<unsigned char> bids(n);
        for (int i = start; i <= end; i++) {
            int bid = bucketOf(prim[i]);
            bids[i - start] = (unsigned char)bid;
            bucketBounds[bid] = grow(bucketBounds[bid], prim[i].b);
            bucketCounts[bid]++;
        }
        // bvh.cpp:89-100: lBounds[i] = own default box grown by bucket i-1 only (not a prefix union); idem rBounds
        Box lB[NB], rB[NB];
        int countPrefix[NB];
        lB[0] = bucketBounds[0];
        rB[NB - 1] = bucketBounds[NB - 1];
        countPrefix[0] = bucketCounts[0];
        for (int i = 1, j = NB - 2; i < NB; i++, j--) {
            lB[i] = grow(emptyBox(), bucketBounds[i - 1]);
            rB[j] = grow(emptyBox(), bucketBounds[j + 1]);
            countPrefix[i] = countPrefix[i - 1] + bucketCounts[i];
        }
        float minSAH = FLT_MAX;
        int divBucket = 0;
        for (int i = 0; i < NB - 1; i++) {
            float SAH = mixf(surfaceArea(lB[i]), surfaceArea(rB[i + 1]), (float)countPrefix[i] / n);
            if (SAH < minSAH) { minSAH = SAH; divBucket = i; }
        }
        // stable front fill / reversed back fill (bvh.cpp:113-121)
        std::vector<Prim> temp(prim.begin() + start, prim.begin() + start + n);
        int divPrim = start, divEnd = end;
        for (int i = 0; i < n; i++) {
            if (bids[i] <= divBucket) prim[divPrim++] = temp[i];
            else prim[divEnd--] = temp[i];
        }
        divPrim = std::min(std::max(divPrim - 1, start), end - 1);                    // bvh.cpp:122
        const int lSize = 2 * (divPrim - start + 1) - 1;
        const int lo = offset + 1, ro = offset + 1 + lSize;
        if (n > TASK_MIN) {
#pragma omp task default(shared)
            node(lo, start, divPrim, depth + 1);
            node(ro, divPrim + 1, end, depth + 1);
#pragma omp taskwait
        } else {
            node(lo, start, divPrim, depth + 1);
            node(ro, divPrim + 1, end, depth + 1);
        }
        // packed record: internal nodes keep their pre-order rank among internal nodes = offset - start
        PackedNode& pn = hs.packed[offset - start];
        const Box& lb = hs.boxes[lo];
        const Box& rb = hs.boxes[ro];
        pn.lmin[0] = lb.pMin.x; pn.lmin[1] = lb.pMin.y; pn.lmin[2] = lb.pMin.z;
        pn.lmax[0] = lb.pMax.x; pn.lmax[1] = lb.pMax.y; pn.lmax[2] = lb.pMax.z;
        pn.rmin[0] = rb.pMin.x; pn.rmin[1] = rb.pMin.y; pn.rmin[2] = rb.pMin.z;
        pn.rmax[0] = rb.pMax.x; pn.rmax[1] = rb.pMax.y; pn.rmax[2] = rb.pMax.z;
        pn.left = HostScene::isLeaf(hs.nodeInfo[lo]) ? ~HostScene::leafPrim(hs.nodeInfo[lo]) : lo - start;
        pn.right = HostScene::isLeaf(hs.nodeInfo[ro]) ? ~HostScene::leafPrim(hs.nodeInfo[ro]) : ro - (divPrim + 1);
        f3 cl = center(lb), cr = center(rb);
        pn.orderMask = (cl.x < cr.x ? 1 : 0) | (cl.y < cr.y ? 2 : 0) | (cl.z < cr.z ? 4 : 0);
        pn.pad = 0;
    }
};

}  // namespace

// sampler.h:79-121.  Two LIFO work lists; entries > 1 donate to entries <= 1 until one list empties.
void buildAliasTable(const std::vector<float>& valuesIn, std::vector<AliasEntry>& table, float& sumAll) {
    std::vector<float> v(valuesIn);
    const int n = (int)v.size();
    sumAll = 0.f;
    for (float x : v) sumAll += x;
    const float scale = (float)n / sumAll;
    for (float& x : v) x *= scale;
    table.assign(n, AliasEntry{0.f, 0});
    std::vector<AliasEntry> over, under;
    over.reserve(n); under.reserve(2 * n);
    for (int i = 0; i < n; i++) (v[i] > 1.f ? over : under).push_back(AliasEntry{v[i], i});
    while (!over.empty() && !under.empty()) {
        AliasEntry g = over.back(); over.pop_back();
        AliasEntry l = under.back(); under.pop_back();
        table[l.failId] = AliasEntry{l.prob, g.failId};
        g.prob -= (1.f - l.prob);
        (g.prob > 1.f ? over : under).push_back(g);
    }
    for (int i = (int)over.size() - 1; i >= 0; i--) table[over[i].failId] = over[i];
    for (int i = (int)under.size() - 1; i >= 0; i--) table[under[i].failId] = under[i];
}

bool buildHostScene(HostScene& hs, std::string& err) {
    auto t0 = std::chrono::steady_clock::now();
    const int T = hs.T;
    if (T <= 0) { err = "scene has no triangles (scene.cpp:192-195)"; return false; }
    if ((int)hs.vertices.size() != 3 * T || (int)hs.normals.size() != 3 * T || (int)hs.materialIds.size() != T) {
        err = "scene arrays have inconsistent sizes"; return false;
    }
    for (int p = 0; p < T; p++) {
        int m = hs.materialIds[p];
        if (m < 0 || m >= (int)hs.materials.size()) { err = "material id out of range"; return false; }
    }
    const int numTex = (int)hs.textures.size();
    hs.anyMaps = hs.anyMRMaps = false;
    for (const RstrMaterial& m : hs.materials) {
        const int ids[4] = {m.baseColorMapId, m.metallicMapId, m.roughnessMapId, m.normalMapId};
        for (int k = 0; k < 4; k++) {
            if (ids[k] >= numTex || ids[k] < (k == 0 ? -2 : -1)) { err = "material references a texture that does not exist"; return false; }
            if (ids[k] != -1) hs.anyMaps = true;
        }
        if (m.metallicMapId > -1 || m.roughnessMapId > -1) hs.anyMRMaps = true;
    }
    if (hs.envMapTexId >= numTex) { err = "environment map references a texture that does not exist"; return false; }
    for (const HostTexture& t : hs.textures)
        if (t.w <= 0 || t.h <= 0 || t.rgb.size() != (size_t)t.w * t.h) { err = "texture has inconsistent size"; return false; }
    // ---- light list (scene.cpp:163-186) + alias table (scene.cpp:154) ----
    hs.lightPrimIds.clear(); hs.lightUnitRadiance.clear(); hs.lightPower.clear();
    for (int p = 0; p < T; p++) {
        const RstrMaterial& m = hs.materials[hs.materialIds[p]];
        if (m.type != 4) continue;
        f3 radianceUnitArea = mk3(m.baseColor[0], m.baseColor[1], m.baseColor[2]);
        float powerUnitArea = luminance(radianceUnitArea) * 2.f * RS_GLM_PI;
        float area = triangleArea(hs.vertices[3 * p], hs.vertices[3 * p + 1], hs.vertices[3 * p + 2]);
        hs.lightPrimIds.push_back(p);
        hs.lightUnitRadiance.push_back(radianceUnitArea);
        hs.lightPower.push_back(powerUnitArea * area);
    }
    if (!hs.lightPowerFromFile.empty()) {
        if (hs.lightPowerFromFile.size() != hs.lightPower.size()) { err = "internal: light power list size"; return false; }
        hs.lightPower = hs.lightPowerFromFile;
    }
    // ---- environment map sampler (scene.cpp:136-152): one more light, appended last ----
    hs.envAlias.clear(); hs.envSumAll = 0.f; hs.envDir.clear();
    if (hs.envMapTexId >= 0) {
        const HostTexture& env = hs.textures[hs.envMapTexId];
        std::vector<float> pdf((size_t)env.w * env.h);
        hs.envDir.resize(pdf.size() * 4);
        const float PiTwo = 6.2831853071795864769252867665590057683943f;
        for (int i = 0; i < env.h; i++) {
            for (int j = 0; j < env.w; j++) {
                int idx = i * env.w + j;
                pdf[idx] = luminance(env.rgb[idx]) * sinf((.5f + i) / env.h * RS_PI);           // scene.cpp:144
                // scene.h:371 wi = toSphere(((.5+x)/W, (.5+y)/H)), mathUtil.h:134-137: depends on the texel only
                float vx = (.5f + j) / env.w * PiTwo, vy = (.5f + i) / env.h * RS_PI;
                hs.envDir[4 * idx + 0] = cosf(vx) * sinf(vy);
                hs.envDir[4 * idx + 1] = cosf(vy);
                hs.envDir[4 * idx + 2] = sinf(vx) * sinf(vy);
                hs.envDir[4 * idx + 3] = 0.f;
            }
        }
        buildAliasTable(pdf, hs.envAlias, hs.envSumAll);
        hs.lightPower.push_back(hs.envSumAll);                                                   // scene.cpp:151
    }
    hs.alias.clear(); hs.sumAll = 0.f; hs.sumLightPowerInv = 0.f;
    if (!hs.lightPower.empty()) {
        buildAliasTable(hs.lightPower, hs.alias, hs.sumAll);
        hs.sumLightPowerInv = 1.f / hs.sumAll;                                       // scene.cpp:493
    }
    // ---- BVH ----
    hs.bvhSize = 2 * T - 1;
    hs.boxes.assign(hs.bvhSize, emptyBox());
    hs.nodeInfo.assign(hs.bvhSize, 0);
    hs.packed.assign(T > 1 ? T - 1 : 1, PackedNode{});
    Builder b(hs);
    b.prim.resize(T);
    for (int i = 0; i < T; i++) {                                                    // bvh.cpp:23-27
        f3 va = hs.vertices[3 * i], vb = hs.vertices[3 * i + 1], vc = hs.vertices[3 * i + 2];
        b.prim[i].id = i;
        b.prim[i].b.pMin = gmin(gmin(va, vb), vc);
        b.prim[i].b.pMax = gmax(gmax(va, vb), vc);
        b.prim[i].c = center(b.prim[i].b);
    }
#pragma omp parallel
#pragma omp single
    b.node(0, 0, T - 1, 1);
    hs.bvhDepth = b.maxDepth.load();
    hs.rootBox = hs.boxes[0];
    hs.rootRef = T > 1 ? 0 : ~HostScene::leafPrim(hs.nodeInfo[0]);
    // ---- packed normals / lights; traced tree + leaf-ordered triangles ----
    hs.triNorm.resize(T);
    for (int p = 0; p < T; p++) {
        TriNorm& nn = hs.triNorm[p];
        memcpy(nn.n0, &hs.normals[3 * p], 36);
        nn.pad[0] = nn.pad[1] = nn.pad[2] = 0.f;
    }
    // ---- textures as float4 texels; per-triangle texture coordinates ----
    hs.texData.clear(); hs.texInfo.clear(); hs.triUV.clear();
    for (const HostTexture& t : hs.textures) {
        int first = (int)(hs.texData.size() / 4);
        hs.texInfo.insert(hs.texInfo.end(), {t.w, t.h, first, 0});
        for (const f3& c : t.rgb) hs.texData.insert(hs.texData.end(), {c.x, c.y, c.z, 0.f});
    }
    if (hs.anyMaps) {
        hs.triUV.resize(T);
        for (int p = 0; p < T; p++) {
            memcpy(hs.triUV[p].t0, &hs.texcoords[6 * (size_t)p], 24);
            hs.triUV[p].pad[0] = hs.triUV[p].pad[1] = 0.f;
        }
    }
    buildFastBVH(hs);
    const int L = (int)hs.lightPrimIds.size();
    hs.lights.resize(L);
    for (int l = 0; l < L; l++) {
        int p = hs.lightPrimIds[l];
        f3 v0 = hs.vertices[3 * p], v1 = hs.vertices[3 * p + 1], v2 = hs.vertices[3 * p + 2];
        LightRec& r = hs.lights[l];
        memcpy(r.v0, &v0, 12); memcpy(r.v1, &v1, 12); memcpy(r.v2, &v2, 12);
        f3 n = triangleNormal(v0, v1, v2);                                           // scene.h:411
        memcpy(r.n, &n, 12);
        f3 Le = hs.lightUnitRadiance[l];
        memcpy(r.Le, &Le, 12);
        float area = triangleArea(v0, v1, v2);                                       // scene.h:419
        float power = luminance(Le) / (area * 2.f * RS_GLM_PI);                      // scene.h:423
        r.pdfArea = power * hs.sumLightPowerInv;                                     // scene.h:424 (first factor)
    }
    hs.buildSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return true;
}

// bvh.cpp:156-193: pre-order emission with the near child first for ordering i
void exportMTBVH(const HostScene& hs, int ordering, std::vector<MTNode>& out) {
    const int N = hs.bvhSize;
    out.assign(N, MTNode{0, 0, 0});
    std::vector<int> stack;
    stack.reserve(256);
    stack.push_back(0);
    int idNew = 0;
    const int dim = ordering / 2;
    const bool lesser = ordering & 1;
    while (!stack.empty()) {
        int orig = stack.back(); stack.pop_back();
        int info = hs.nodeInfo[orig];
        bool leaf = HostScene::isLeaf(info);
        int size = leaf ? 1 : info;
        out[idNew] = MTNode{leaf ? HostScene::leafPrim(info) : -1, orig, idNew + size};
        idNew++;
        if (leaf) continue;
        int linfo = hs.nodeInfo[orig + 1];
        int lsize = HostScene::isLeaf(linfo) ? 1 : linfo;
        int left = orig + 1, right = orig + 1 + lsize;
        bool lt = comp(center(hs.boxes[left]), dim) < comp(center(hs.boxes[right]), dim);
        if (lt ^ lesser) std::swap(left, right);
        stack.push_back(right);
        stack.push_back(left);
    }
}

// sceneStructs.h:88-102; glm::inverse(mat3) = type_mat3x3.inl:37-58
void cameraUpdate(RstrCamera& c) {
    float yaw = radians(c.rotation[0]), pitch = radians(c.rotation[1]);
    f3 view;
    view.x = cosf(yaw) * cosf(pitch);
    view.z = sinf(yaw) * cosf(pitch);
    view.y = sinf(pitch);
    view = normalize(view);
    f3 right = normalize(cross(view, mk3(0.f, 1.f, 0.f)));
    f3 up = normalize(cross(right, view));
    memcpy(c.view, &view, 12); memcpy(c.right, &right, 12); memcpy(c.up, &up, 12);
    const float m[3][3] = {{right.x, right.y, right.z}, {up.x, up.y, up.z}, {view.x, view.y, view.z}};   // m[col][row]
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
    for (int col = 0; col < 3; col++)
        for (int row = 0; row < 3; row++) c.rotationMatInv[col * 3 + row] = inv[col][row];
}

}  // namespace rs
