// scene_host.h -- host-side scene build: light list, alias table, BVH (bit-identical to the
// reference's builder output), and the cache-line-packed device layouts derived from it.
#pragma once

#include <string>
#include <vector>

#include "../../include/restir_b200.h"
#include "vecmath.h"

#ifndef RS_BVH4
#define RS_BVH4 0               /* 0: the traced tree is the binned-SAH BVH2 (FastNode); 1: its 4-wide collapse (FastNode4) -- measured
                                   6 % SLOWER on B200 (profiles/README.md): kept as a build option for A/B runs only */
#endif

namespace rs {

struct Box {            // reference AABB, 24 B (bvh.h:159-160)
    f3 pMin, pMax;
};

// ---- device-facing packed layouts ----
// One 64-byte record per INTERNAL node: both children's boxes + links, fetched as 4 x LDG.128.
struct alignas(16) PackedNode {
    float lmin[3], lmax[3];   // left child box
    float rmin[3], rmax[3];   // right child box
    int left, right;          // >= 0: packed index of an internal child; < 0: ~primId of a leaf child
    int orderMask;            // bit d: center(left)[d] < center(right)[d]  (bvh.cpp:184-188)
    int pad;
};
static_assert(sizeof(PackedNode) == 64, "PackedNode");

// One 64-byte record per internal node of the traced (binned-SAH) tree: both children's padded boxes + links.
struct alignas(16) FastNode {
    float lmin[3], lmax[3];
    float rmin[3], rmax[3];
    int left, right;          // >= 0: node index; < 0: leaf = 0x80000000 | (count-1) << 27 | firstTriangle
    int pad[2];
};
static_assert(sizeof(FastNode) == 64, "FastNode");

// One 128-byte record (= one L1 line) per internal node of the 4-wide tree the kernels walk: the four children's padded
// boxes, structure-of-arrays so that one LDG.128 brings one plane of all four, + links.  Collapsed from the BVH2.
struct alignas(16) FastNode4 {
    float lox[4], hix[4], loy[4], hiy[4], loz[4], hiz[4];
    int child[4];             // >= 0: node index; < 0: leaf (same encoding as FastNode); unused slot: 0x7fffffff
    int pad[4];
};
static_assert(sizeof(FastNode4) == 128, "FastNode4");

struct alignas(16) TriGeom {  // 48 B: raw vertices (Moller-Trumbore recomputes the edges exactly as intersections.h:20-21)
    float v0[3], v1[3], v2[3];
    int matId;
    int pad[2];               // pad[0] = original primitive id (records are stored in the traced tree's leaf order)
};
static_assert(sizeof(TriGeom) == 48, "TriGeom");

struct alignas(16) TriNorm {  // 48 B: vertex normals, read once per ray at the hit
    float n0[3], n1[3], n2[3];
    float pad[3];
};
static_assert(sizeof(TriNorm) == 48, "TriNorm");

struct alignas(16) LightRec { // 64 B: everything sampleDirectLightNoVisibility needs after the alias lookup
    float v0[3], v1[3], v2[3];
    float n[3];               // Math::triangleNormal(v0,v1,v2)                    (scene.h:411)
    float Le[3];              // lightUnitRadiance                                 (scene.h:420)
    float pdfArea;            // luminance(Le) / (area*2*pi) * sumLightPowerInv    (scene.h:423-424)
};
static_assert(sizeof(LightRec) == 64, "LightRec");

struct AliasEntry {           // sampler.h:63-67
    float prob;
    int failId;
};

struct MTNode {               // bvh.h:163-171
    int prim, box, miss;
};

struct HostTexture {          // image.h:7-39: width x height linear-RGB texels
    int w = 0, h = 0;
    std::vector<f3> rgb;
};

struct alignas(16) TriUV {    // 32 B: texture coordinates of the three vertices (scene.h:144-150), read only for mapped materials
    float t0[2], t1[2], t2[2];
    float pad[2];
};
static_assert(sizeof(TriUV) == 32, "TriUV");

struct HostScene {
    int T = 0;
    std::vector<f3> vertices, normals;
    std::vector<float> texcoords;          // 2 per vertex
    std::vector<int> materialIds;
    std::vector<RstrMaterial> materials;

    // reference-format build outputs
    int bvhSize = 0, bvhDepth = 0;
    std::vector<Box> boxes;                // pre-order, Scene::boundingBoxes
    std::vector<int> nodeInfo;             // pre-order: > 0 subtree size (internal), <= 0: -(primId) - 1 ... see isLeaf()
    std::vector<int> lightPrimIds;
    std::vector<f3> lightUnitRadiance;
    std::vector<float> lightPower;
    // Scene files only: power of every emissive triangle as the reference computes it.  scene.cpp:176-180 indexes the
    // scene-wide flattened vertex array with the INSTANCE-LOCAL vertex index, so for every instance but the first the
    // area comes from an unrelated triangle of the first instances.  The alias table and sumLightPowerInv inherit that;
    // the per-candidate area in scene.h:419 does not.  Reproduced verbatim for drop-in parity.
    std::vector<float> lightPowerFromFile;
    std::vector<AliasEntry> alias;         // lightSampler: emissive triangles, then the environment map (if any) as the last entry
    float sumAll = 0.f, sumLightPowerInv = 0.f;
    // textures + environment map (scene.cpp:136-152, 362-375)
    std::vector<HostTexture> textures;
    int envMapTexId = -1;
    std::vector<AliasEntry> envAlias;      // envMapSampler over the W*H texels
    float envSumAll = 0.f;
    bool anyMaps = false;                  // some material references a texture / the procedural pattern
    bool anyMRMaps = false;                // ... a metallic or roughness map (phase B then needs the per-pixel values)
    std::vector<float> texData;            // all textures as float4 texels (rgb + pad), concatenated
    std::vector<int> texInfo;              // per texture {w, h, first texel, 0}
    std::vector<float> envDir;             // float4 per environment texel: Math::toSphere of its centre (scene.h:371), evaluated on the host
    std::vector<TriUV> triUV;              // original primitive order (empty unless anyMaps)

    // packed device layouts
    Box rootBox;
    int rootRef = 0;
    std::vector<PackedNode> packed;
    std::vector<TriNorm> triNorm;          // original primitive order

    // traced tree (bvh_fast.cpp)
    std::vector<FastNode> fastNodes;       // binned-SAH BVH2 (host only: input of the 4-wide collapse)
    std::vector<FastNode4> fastNodes4;     // the tree that is traced
    int fastRoot4 = 0, fastDepth4 = 0;
    std::vector<int> fastOrder;            // leaf order -> original primitive id
    std::vector<int> primToFast;           // original primitive id -> position in fastTris
    std::vector<TriGeom> fastTris;         // triangles in leaf order
    std::vector<int> rank;                 // [6][T]: visiting rank of each primitive's leaf in the reference's 6 orderings
    int fastRoot = 0, fastDepth = 0;
    float fastRootMin[3] = {0, 0, 0}, fastRootMax[3] = {0, 0, 0};
    double fastBuildSeconds = 0.0;
    std::vector<LightRec> lights;
    double buildSeconds = 0.0;

    static bool isLeaf(int info) { return info < 0; }
    static int leafPrim(int info) { return -info - 1; }
};

// scene.cpp:159-215 minus the upload.  Returns false (and sets err) on invalid input.
bool buildHostScene(HostScene& hs, std::string& err);
// binned-SAH BVH2 that the kernels trace + tie-break ranks (bvh_fast.cpp)
void buildFastBVH(HostScene& hs);
// bvh.cpp:133-201: the reference's i-th threaded ordering, regenerated from the single tree
void exportMTBVH(const HostScene& hs, int ordering, std::vector<MTNode>& out);
// sampler.h:79-121
void buildAliasTable(const std::vector<float>& values, std::vector<AliasEntry>& table, float& sumAll);

// Scene::Scene(filename): scene text + OBJ -> flattened arrays + camera (scene.cpp:96-131, 222-433, 159-190)
bool loadSceneFile(const std::string& path, HostScene& hs, RstrCamera& cam, std::string& err);
// Image::Image(filename) (image.cpp:16-33): PNG / Radiance .hdr -> linear float RGB; flipY = stbi_set_flip_vertically_on_load
bool loadImageRGB(const std::string& path, bool flipY, HostTexture& out, std::string& err);
// Image::savePNG (image.cpp:41-57): W x H x 3 bytes, top row first
bool writePNG(const std::string& path, int W, int H, const unsigned char* rgb, std::string& err);
// Image::saveJPG (image.cpp:59-75): stbi_write_jpg at the given quality (the reference uses 90); image_jpeg.cpp
bool writeJPG(const std::string& path, int W, int H, const unsigned char* rgb, int quality, std::string& err);
// Camera::update (sceneStructs.h:88-102)
void cameraUpdate(RstrCamera& c);

}  // namespace rs
