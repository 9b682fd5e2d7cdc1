/*
 * ref_harness.cpp -- drives the REFERENCE'S OWN CODE on the CPU.  TEST INFRASTRUCTURE ONLY.
 *
 * Compiled by oracle/Makefile with g++ against the headers and host .cpp files where they lie
 * under /root/reference (never copied into this repo); the output is oracle/_ref/libref_harness.so
 * (git-ignored).  It exports the same C API as restir_oracle.cpp so tests can diff the two.
 *
 * What is the reference's code here: DevScene::{intersect,testOcclusion,sampleDirectLight*},
 * AABB::intersect, intersectTriangle, Material::BSDF, Reservoir<DirectLiSample>, Camera::{sample,
 * getRasterCoord,update}, makeSeededRandomEngine/sample1D (thrust minstd), DiscreteSampler1D,
 * BVHBuilder::build, Scene(file)/Scene::buildDevData/DevScene::create (over oracle/fake_cudart.cpp).
 * Under g++ the __device__ attribute is ignored, so these are ordinary functions.
 *
 * What is restated here (a __global__ body cannot be compiled by g++): renderGBuffer
 * (gbuffer.cu:3-73), ReSTIRDirectKernel + its file-static helpers (restir.cu:20-231), PTDirectKernel
 * (pathtrace.cu:279-328) and the light-list loop of Scene::buildDevData (scene.cpp:159-190) for
 * scenes given as flat arrays.  Two mandatory deviations: RNG draws are sequenced
 * x,y,z,w explicitly (g++ evaluates sample4D's arguments right-to-left, nvcc left-to-right) and
 * spatial reuse is two-phase.  Float->int conversions the reference performs on the GPU use CUDA
 * semantics (f2i_cuda).
 */
#include "scene.h"
#include "restir.h"

#include <climits>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include <stb_image.h>

#include "restir_oracle.h"

using DirectReservoir = Reservoir<DirectLiSample>;
static_assert(sizeof(DirectReservoir) == 36, "reservoir layout");
static_assert(sizeof(Camera) == sizeof(OrcCamera), "camera layout");
static_assert(sizeof(Material) == sizeof(OrcMaterial), "material layout");

static inline int f2i_cuda(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.f) return INT_MAX;
    if (f <= -2147483648.f) return INT_MIN;
    return (int)f;
}

struct OrcScene {
    DevScene dev;                 /* host pointers */
    Scene* fileScene = nullptr;   /* when loaded through the reference's parser */
    std::vector<glm::vec3> vertices, normals;
    std::vector<glm::vec2> texcoords;
    std::vector<int> materialIds;
    std::vector<Material> materials;
    std::vector<AABB> boxes;
    std::vector<std::vector<MTBVHNode>> nodes;
    std::vector<int> lightPrimIds;
    std::vector<glm::vec3> lightUnitRadiance;
    DiscreteSampler1D<float> lightSampler, envMapSampler;
    std::vector<float> lightPower;
    std::vector<std::vector<glm::vec3>> texData;
    std::vector<DevTextureObj> texObjs;
    int T = 0;
};

struct OrcFrame {
    OrcScene* sc;
    GBuffer g;
    std::vector<glm::vec3> albedo, normal[2], radiance;
    std::vector<int> motion, matId[2];
    std::vector<float> depth[2];
    bool haveLast = false;
    std::vector<DirectReservoir> resv, lastResv, temp;
    bool first = true;
    struct Carry { Sampler rng; int status; DirectReservoir r; Intersection is; Material mat; glm::vec3 direct; };
    std::vector<Carry> carry;
};

static glm::vec4 draw4(Sampler& rng) {   /* nvcc device order: x first */
    float a = sample1D(rng), b = sample1D(rng), c = sample1D(rng), d = sample1D(rng);
    return glm::vec4(a, b, c, d);
}

extern "C" {

void orc_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : omp_get_num_procs());
#endif
}
int orc_get_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static void finishScene(OrcScene* sc) {
    DevScene& d = sc->dev;
    d.vertices = sc->vertices.data(); d.normals = sc->normals.data(); d.texcoords = sc->texcoords.data();
    d.boundingBoxes = sc->boxes.data();
    for (int i = 0; i < 6; i++) d.BVHNodes[i] = sc->nodes[i].data();
    d.materialIds = sc->materialIds.data(); d.materials = sc->materials.data();
    d.lightPrimIds = sc->lightPrimIds.data(); d.lightUnitRadiance = sc->lightUnitRadiance.data();
    d.lightSampler.devBinomDistribs = sc->lightSampler.binomDistribs.data();
    d.lightSampler.length = (int)sc->lightSampler.binomDistribs.size();
    d.lightSampler.sumAll = sc->lightSampler.sumAll;
    d.sumLightPowerInv = 1.f / sc->lightSampler.sumAll;               /* scene.cpp:493 */
}

OrcScene* orc_scene_create(int T, const float* vertices, const float* normals, const float* texcoords,
                           const int* materialIds, int numMaterials, const OrcMaterial* materials) {
    OrcScene* sc = new OrcScene;
    sc->T = T;
    sc->vertices.resize(T * 3); sc->normals.resize(T * 3); sc->texcoords.resize(T * 3);
    memcpy(sc->vertices.data(), vertices, sizeof(float) * 9 * T);
    memcpy(sc->normals.data(), normals, sizeof(float) * 9 * T);
    if (texcoords) memcpy(sc->texcoords.data(), texcoords, sizeof(float) * 6 * T);
    sc->materialIds.assign(materialIds, materialIds + T);
    sc->materials.resize(numMaterials);
    memcpy((void*)sc->materials.data(), materials, sizeof(Material) * numMaterials);
    std::vector<float>& lightPower = sc->lightPower;
    for (int p = 0; p < T; p++) {                                      /* scene.cpp:163-186 */
        const Material& material = sc->materials[sc->materialIds[p]];
        if (material.type != Material::Light) continue;
        glm::vec3 radianceUnitArea = material.baseColor;
        float powerUnitArea = Math::luminance(radianceUnitArea) * 2.f * glm::pi<float>();
        float area = Math::triangleArea(sc->vertices[p * 3], sc->vertices[p * 3 + 1], sc->vertices[p * 3 + 2]);
        sc->lightPrimIds.push_back(p);
        sc->lightUnitRadiance.push_back(radianceUnitArea);
        lightPower.push_back(powerUnitArea * area);
    }
    if (!lightPower.empty()) sc->lightSampler = DiscreteSampler1D<float>(lightPower);   /* scene.cpp:154 */
    sc->dev.BVHSize = BVHBuilder::build(sc->vertices, sc->boxes, sc->nodes);            /* scene.cpp:199 */
    finishScene(sc);
    return sc;
}

/* Scene::textures / envMapTexId as Scene::addTexture + createLightSampler (scene.cpp:136-157) would leave them, with
 * DevTextureObj / DevDiscreteSampler1D pointing at host memory */
int orc_scene_set_textures(OrcScene* sc, int numTextures, const int* widths, const int* heights, const float* const* rgb, int envMapTexId) {
    sc->texData.resize(numTextures); sc->texObjs.resize(numTextures);
    for (int t = 0; t < numTextures; t++) {
        size_t n = (size_t)widths[t] * heights[t];
        sc->texData[t].resize(n);
        memcpy((void*)sc->texData[t].data(), rgb[t], n * sizeof(glm::vec3));
        sc->texObjs[t].width = widths[t]; sc->texObjs[t].height = heights[t]; sc->texObjs[t].devData = sc->texData[t].data();
    }
    for (const Material& m : sc->materials) {
        const int ids[4] = {m.baseColorMapId, m.metallicMapId, m.roughnessMapId, m.normalMapId};
        for (int k = 0; k < 4; k++)
            if (ids[k] >= numTextures || ids[k] < (k == 0 ? ProceduralTexId : NullTextureId)) return -1;
    }
    if (envMapTexId >= numTextures) return -1;
    sc->dev.textures = sc->texObjs.data();
    if (envMapTexId >= 0) {
        const DevTextureObj& envMap = sc->texObjs[envMapTexId];
        std::vector<float> pdf(envMap.width * envMap.height);
        for (int i = 0; i < envMap.height; i++) {                                       /* scene.cpp:141-146 */
            for (int j = 0; j < envMap.width; j++) {
                int idx = i * envMap.width + j;
                pdf[idx] = Math::luminance(envMap.devData[idx]) * glm::sin((.5f + i) / envMap.height * Pi);
            }
        }
        sc->envMapSampler = DiscreteSampler1D<float>(pdf);
        sc->lightPower.push_back(sc->envMapSampler.sumAll);                             /* scene.cpp:151 */
        sc->lightSampler = DiscreteSampler1D<float>(sc->lightPower);                    /* scene.cpp:154 */
        sc->dev.envMap = sc->texObjs.data() + envMapTexId;                              /* scene.cpp:496-499 */
        sc->dev.envMapSampler.devBinomDistribs = sc->envMapSampler.binomDistribs.data();
        sc->dev.envMapSampler.length = (int)sc->envMapSampler.binomDistribs.size();
        sc->dev.envMapSampler.sumAll = sc->envMapSampler.sumAll;
        finishScene(sc);
    }
    return 0;
}
const void* orc_scene_env_alias(const OrcScene* s, int* lengthOut, float* sumAllOut) {
    if (lengthOut) *lengthOut = s->dev.envMapSampler.length;
    if (sumAllOut) *sumAllOut = s->dev.envMapSampler.sumAll;
    return s->dev.envMapSampler.devBinomDistribs;
}

/* Image::Image(filename) (image.cpp:16-33) = stbi_loadf(..., 3) under the loader state Scene::Scene sets (scene.cpp:97-98,124) */
int ref_image_load(const char* path, int flipY, int* w, int* h, float* out, size_t capacityBytes) {
    stbi_ldr_to_hdr_gamma(1.f);
    stbi_set_flip_vertically_on_load(flipY);
    int n;
    float* probe = stbi_loadf(path, w, h, &n, 3);      /* Image's ctor would `throw;` (terminate) on failure */
    if (!probe) return -1;
    stbi_image_free(probe);
    Image img(path);
    if (out) {
        if (capacityBytes < img.byteSize()) return -2;
        memcpy(out, img.data(), img.byteSize());
    }
    return 0;
}

/* Image::setPixel + Image::savePNG / saveJPG (image.cpp:35-75): writes <base>.png or <base>.jpg from float RGB */
int ref_image_save(const char* base, int w, int h, const float* rgb, int jpg) {
    Image img(w, h);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) img.setPixel(x, y, glm::vec3(rgb[3 * (y * w + x)], rgb[3 * (y * w + x) + 1], rgb[3 * (y * w + x) + 2]));
    if (jpg) img.saveJPG(base); else img.savePNG(base);
    return 0;
}

/* the reference's own parser + flattening + upload (over fake cudart) */
OrcScene* ref_scene_load_file(const char* path) {
    OrcScene* sc = new OrcScene;
    sc->fileScene = new Scene(path);
    sc->fileScene->buildDevData();
    sc->dev = sc->fileScene->hstScene;
    sc->T = (sc->dev.BVHSize + 1) / 2;
    return sc;
}
void ref_scene_camera(const OrcScene* sc, OrcCamera* out) { memcpy(out, &sc->fileScene->camera, sizeof(OrcCamera)); }
int ref_scene_num_tris(const OrcScene* sc) { return sc->T; }
int ref_scene_texture_info(const OrcScene* sc, int i, int* w, int* h, int* isEnv) {
    if (!sc->fileScene || i < 0 || i >= (int)sc->fileScene->textures.size()) return -1;
    *w = sc->fileScene->textures[i]->width(); *h = sc->fileScene->textures[i]->height(); *isEnv = sc->fileScene->envMapTexId == i;
    return (int)sc->fileScene->textures.size();
}
int ref_scene_num_materials(const OrcScene* sc) { return sc->fileScene ? (int)sc->fileScene->materials.size() : (int)sc->materials.size(); }
const void* ref_scene_array(const OrcScene* sc, int which) {
    switch (which) {
    case 0: return sc->dev.vertices;
    case 1: return sc->dev.normals;
    case 2: return sc->dev.texcoords;
    case 3: return sc->dev.materialIds;
    case 4: return sc->dev.materials;
    }
    if (which >= 32 && sc->fileScene && which - 32 < (int)sc->fileScene->textures.size()) return sc->fileScene->textures[which - 32]->data();
    return nullptr;
}

void orc_scene_destroy(OrcScene* s) { delete s; }
int orc_scene_bvh_size(const OrcScene* s) { return s->dev.BVHSize; }
int orc_scene_bvh_depth(const OrcScene*) { return -1; }
const float* orc_scene_boxes(const OrcScene* s) { return (const float*)s->dev.boundingBoxes; }
const int* orc_scene_mtbvh(const OrcScene* s, int i) { return (const int*)s->dev.BVHNodes[i]; }
int orc_scene_num_lights(const OrcScene* s) { return s->dev.lightSampler.length; }
const int* orc_scene_light_prim_ids(const OrcScene* s) { return s->dev.lightPrimIds; }
const float* orc_scene_light_radiance(const OrcScene* s) { return (const float*)s->dev.lightUnitRadiance; }
const void* orc_scene_alias_table(const OrcScene* s) { return s->dev.lightSampler.devBinomDistribs; }
float orc_scene_sum_light_power(const OrcScene* s) { return s->dev.lightSampler.sumAll; }

void orc_camera_update(OrcCamera* c) { ((Camera*)c)->update(); }

OrcFrame* orc_frame_create(const OrcScene* sc, int w, int h) {
    OrcFrame* f = new OrcFrame;
    f->sc = (OrcScene*)sc;
    size_t P = (size_t)w * h;
    f->albedo.assign(P, glm::vec3(0.f)); f->radiance.assign(P, glm::vec3(0.f));
    f->motion.assign(P, 0);
    for (int i = 0; i < 2; i++) { f->normal[i].assign(P, glm::vec3(0.f)); f->matId[i].assign(P, 0); f->depth[i].assign(P, 0.f); }
    f->g.devAlbedo = f->albedo.data(); f->g.devMotion = f->motion.data();
    for (int i = 0; i < 2; i++) { f->g.devNormal[i] = f->normal[i].data(); f->g.devPrimId[i] = f->matId[i].data(); f->g.devDepth[i] = f->depth[i].data(); }
    f->g.width = w; f->g.height = h; f->g.frameIdx = 0;
    f->resv.assign(P, DirectReservoir()); f->lastResv.assign(P, DirectReservoir()); f->temp.assign(P, DirectReservoir());
    f->carry.resize(P);
    return f;
}
void orc_frame_destroy(OrcFrame* f) { delete f; }
void orc_frame_reset(OrcFrame* f) { f->first = true; }

/* gbuffer.cu:3-73 */
void orc_gbuffer_render(OrcFrame* f, const OrcCamera* ocam) {
    Camera cam = *(const Camera*)ocam;
    DevScene* scene = &f->sc->dev;
    GBuffer gBuffer = f->g;
    if (!f->haveLast) gBuffer.lastCamera = cam;      /* reference: uninitialised on frame 0 (gbuffer.h:56) */
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < cam.resolution.y; y++) {
        for (int x = 0; x < cam.resolution.x; x++) {
            int idx = y * cam.resolution.x + x;
            float aspect = float(cam.resolution.x) / cam.resolution.y;
            float tanFovY = glm::tan(glm::radians(cam.fov.y));
            glm::vec2 pixelSize = 1.f / glm::vec2(cam.resolution);
            glm::vec2 scr = glm::vec2(x, y) * pixelSize;
            glm::vec2 ruv = scr + pixelSize * glm::vec2(.5f);
            glm::vec3 pLens(0.f);
            glm::vec3 pFocusPlane = glm::vec3((1.f - ruv * 2.f) * glm::vec2(aspect, 1.f) * tanFovY, 1.f) * cam.focalDist;
            glm::vec3 dir = pFocusPlane - pLens;
            Ray ray;
            ray.direction = glm::normalize(glm::mat3(cam.right, cam.up, cam.view) * dir);
            ray.origin = cam.position + cam.right * pLens.x + cam.up * pLens.y;
            Intersection intersec;
            scene->intersect(ray, intersec);
            if (intersec.primId != NullPrimitive) {
                int matId = intersec.matId;
                if (scene->materials[intersec.matId].type == Material::Type::Light) matId = NullPrimitive - 1;
                Material material = scene->getTexturedMaterialAndSurface(intersec);
                gBuffer.devAlbedo[idx] = material.baseColor;
                gBuffer.normal()[idx] = ENCODE_NORM(intersec.norm);
                gBuffer.primId()[idx] = matId;
                gBuffer.depth()[idx] = glm::distance(intersec.pos, ray.origin);
                /* getRasterCoord (sceneStructs.h:43-46) with the GPU's float->int conversion */
                glm::vec2 ndc = gBuffer.lastCamera.getRasterUV(intersec.pos);
                glm::vec2 rc = glm::vec2(gBuffer.lastCamera.resolution) * ndc;
                glm::ivec2 lastPos(f2i_cuda(rc.x), f2i_cuda(rc.y));
                if (lastPos.x >= 0 && lastPos.x < gBuffer.width && lastPos.y >= 0 && lastPos.y < gBuffer.height)
                    gBuffer.devMotion[idx] = lastPos.y * cam.resolution.x + lastPos.x;
                else
                    gBuffer.devMotion[idx] = -1;
            } else {
                glm::vec3 albedo(0.f);
                if (scene->envMap != nullptr) {                                          /* gbuffer.cu:59-62 */
                    glm::vec2 uv = Math::toPlane(ray.direction);
                    albedo = scene->envMap->linearSample(uv);
                }
                gBuffer.devAlbedo[idx] = albedo;
                gBuffer.normal()[idx] = GBuffer::NormT(0.f);
                gBuffer.primId()[idx] = NullPrimitive;
                gBuffer.depth()[idx] = 1.f;
                gBuffer.devMotion[idx] = 0;
            }
        }
    }
}

void orc_gbuffer_update(OrcFrame* f, const OrcCamera* cam) {
    f->g.lastCamera = *(const Camera*)cam;    /* gbuffer.cu:75-78 (GBuffer::update lives in a .cu) */
    f->g.frameIdx ^= 1;
    f->haveLast = true;
}
void orc_last_trace_stats(OrcFrame*, uint64_t* n, uint64_t* t, uint64_t* r) { *n = *t = *r = 0; }

/* restir.cu:20-45 */
static DirectReservoir findTemporalNeighbor(DirectReservoir* reservoir, int idx, const GBuffer& gBuffer) {
    int primId = gBuffer.primId()[idx];
    int lastIdx = gBuffer.devMotion[idx];
    bool diff = false;
    if (lastIdx < 0) diff = true;
    else if (primId <= NullPrimitive) diff = true;
    else if (gBuffer.lastPrimId()[lastIdx] != primId) diff = true;
    else {
        glm::vec3 norm = DECODE_NORM(gBuffer.normal()[idx]);
        glm::vec3 lastNorm = DECODE_NORM(gBuffer.lastNormal()[lastIdx]);
        float depth = gBuffer.depth()[idx];
        float pdepth = gBuffer.lastDepth()[lastIdx];
        if (Math::absDot(norm, lastNorm) < .9f || glm::abs(pdepth - depth) > depth * .1f) diff = true;
    }
    return diff ? DirectReservoir() : reservoir[lastIdx];
}

/* restir.cu:47-85 */
static DirectReservoir findSpatialNeighborDisk(DirectReservoir* reservoir, int x, int y, const GBuffer& gBuffer, glm::vec2 r, float Radius) {
    int idx = y * gBuffer.width + x;
    glm::vec2 p = Math::toConcentricDisk(r.x, r.y) * Radius;
    int px = f2i_cuda(x + .5f + p.x);
    int py = f2i_cuda(y + .5f + p.y);
    int pidx = py * gBuffer.width + px;
    bool diff = false;
    if (px < 0 || px >= gBuffer.width || py < 0 || py >= gBuffer.height || (px == x && py == y)) diff = true;
    else if (gBuffer.primId()[pidx] != gBuffer.primId()[idx]) diff = true;
    else {
        glm::vec3 norm = DECODE_NORM(gBuffer.normal()[idx]);
        glm::vec3 pnorm = DECODE_NORM(gBuffer.normal()[pidx]);
        if (glm::dot(norm, pnorm) < .9f) diff = true;
        float depth = gBuffer.depth()[idx];
        float pdepth = gBuffer.depth()[pidx];
        if (glm::abs(depth - pdepth) > depth * .1f) diff = true;
    }
    return diff ? DirectReservoir() : reservoir[pidx];
}

/* the reference's preClampedMerge<20> is a template on the cap; run-time cap for config sweeps */
static void preClampedMergeRT(DirectReservoir& self, int cap, DirectReservoir rhs, float r) {   /* restir.h:96-102 */
    if (self.numSamples > 0) rhs.clamp((cap - 1) * self.numSamples);
    self.merge(rhs, r);
}

void orc_restir_direct(OrcFrame* f, const OrcCamera* ocam, const OrcParams* prm, int looper, int iter) {
    Camera cam = *(const Camera*)ocam;
    DevScene* scene = &f->sc->dev;
    const GBuffer gBuffer = f->g;
    const bool first = f->first;
    const int reuseState = prm->reuse;
    DirectReservoir* reservoirOut = f->resv.data();
    DirectReservoir* reservoirIn = f->lastResv.data();
    DirectReservoir* reservoirTemp = f->temp.data();
    /* phase A: restir.cu:119-192 */
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < cam.resolution.y; y++) {
        for (int x = 0; x < cam.resolution.x; x++) {
            int index = y * cam.resolution.x + x;
            OrcFrame::Carry& c = f->carry[index];
            Sampler rng = makeSeededRandomEngine(looper, index, 0, scene->sampleSequence);
            Ray ray = cam.sample(x, y, draw4(rng));
            Intersection intersec;
            scene->intersect(ray, intersec);
            if (intersec.primId == NullPrimitive) {                                      /* restir.cu:133-138 */
                c.status = 0; c.direct = glm::vec3(0.f);
                if (scene->envMap != nullptr) c.direct = scene->envMap->linearSample(Math::toPlane(ray.direction));
                continue;
            }
            Material material = scene->getTexturedMaterialAndSurface(intersec);
            material.baseColor = glm::vec3(1.f);
            if (material.type == Material::Type::Light) { c.status = 1; continue; }
            intersec.wo = -ray.direction;
            bool deltaBSDF = (material.type == Material::Type::Dielectric);
            if (!deltaBSDF && glm::dot(intersec.norm, intersec.wo) < 0.f) intersec.norm = -intersec.norm;
            DirectReservoir reservoir;
            for (int i = 0; i < prm->numCandidates; i++) {
                glm::vec3 Li;
                glm::vec3 wi;
                float dist = 0.f;
                glm::vec4 r4 = draw4(rng);
                float p = scene->sampleDirectLightNoVisibility(intersec.pos, r4, Li, wi, dist);
                glm::vec3 g = Li * material.BSDF(intersec.norm, intersec.wo, wi) * Math::satDot(intersec.norm, wi);
                float weight = DirectReservoir::toScalar(g / p);
                if (Math::isNanOrInf(weight) || p <= 0.f) weight = 0.f;
                reservoir.update({ Li, wi, dist }, weight, sample1D(rng));
            }
            DirectLiSample sample = reservoir.sample;
            if (scene->testOcclusion(intersec.pos, intersec.pos + sample.wi * sample.dist)) reservoir.weight = 0.f;
            if (!first && (reuseState & ReservoirReuse::Temporal)) {
                DirectReservoir temporal = findTemporalNeighbor(reservoirIn, index, gBuffer);
                if (!temporal.invalid()) preClampedMergeRT(reservoir, prm->temporalCap, temporal, sample1D(rng));
            }
            DirectReservoir tempReservoir = reservoir;
            if (reuseState & ReservoirReuse::Spatial) {
                reservoir.checkValidity();
                reservoirTemp[index] = reservoir;
            }
            tempReservoir.checkValidity();
            reservoirOut[index] = tempReservoir;
            c.status = 2; c.rng = rng; c.r = reservoir; c.is = intersec; c.mat = material;
        }
    }
    /* phase B: restir.cu:196-230; passes 2..n = the commented-out block :201-209 (preClampedMerge<4>) */
    const int passes = (reuseState & ReservoirReuse::Spatial) ? (prm->spatialPasses < 1 ? 1 : prm->spatialPasses) : 0;
    for (int pass = 1; pass <= passes; pass++) {
        if (pass > 1) {
#pragma omp parallel for schedule(static)
            for (int i = 0; i < cam.resolution.x * cam.resolution.y; i++)
                if (f->carry[i].status == 2) reservoirTemp[i] = f->carry[i].r;
        }
#pragma omp parallel for schedule(dynamic, 4)
        for (int y = 0; y < cam.resolution.y; y++) {
            for (int x = 0; x < cam.resolution.x; x++) {
                OrcFrame::Carry& c = f->carry[y * cam.resolution.x + x];
                if (c.status != 2) continue;
                Sampler rng = c.rng;
                DirectReservoir agg;
                for (int i = 0; i < prm->numSpatial; i++) {
                    float rx = sample1D(rng), ry = sample1D(rng);
                    DirectReservoir spatial = findSpatialNeighborDisk(reservoirTemp, x, y, gBuffer, glm::vec2(rx, ry), prm->spatialRadius);
                    if (!spatial.invalid()) agg.merge(spatial, sample1D(rng));
                }
                if (pass == 1) {
                    if (!agg.invalid() && !c.r.invalid()) c.r.merge(agg, sample1D(rng));
                } else {
                    if (!agg.invalid()) c.r.preClampedMerge<4>(agg, sample1D(rng));
                }
                c.rng = rng;
            }
        }
    }
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < cam.resolution.y; y++) {
        for (int x = 0; x < cam.resolution.x; x++) {
            int index = y * cam.resolution.x + x;
            OrcFrame::Carry& c = f->carry[index];
            glm::vec3 direct(0.f);
            if (c.status == 0) direct = c.direct;
            if (c.status == 1) direct = glm::vec3(1.f);
            if (c.status == 2) {
                DirectReservoir reservoir = c.r;
                const Intersection& intersec = c.is;
                const Material& material = c.mat;
                DirectLiSample sample = reservoir.sample;
                if (!reservoir.invalid()) {
                    glm::vec3 LiBSDF = sample.Li * material.BSDF(intersec.norm, intersec.wo, sample.wi);
                    direct = LiBSDF / DirectReservoir::toScalar(LiBSDF) * reservoir.weight / static_cast<float>(reservoir.numSamples);
                }
                if (Math::hasNanOrInf(direct)) direct = glm::vec3(0.f);
            }
            direct *= gBuffer.devAlbedo[index];
            f->radiance[index] = (f->radiance[index] * float(iter) + direct) / float(iter + 1);
        }
    }
    std::swap(f->resv, f->lastResv);
    f->first = false;
}

/* pathtrace.cu:279-328 */
void orc_pathtrace_direct(OrcFrame* f, const OrcCamera* ocam, int looper, int iter) {
    Camera cam = *(const Camera*)ocam;
    DevScene* scene = &f->sc->dev;
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < cam.resolution.y; y++) {
        for (int x = 0; x < cam.resolution.x; x++) {
            glm::vec3 direct(0.f);
            int index = y * cam.resolution.x + x;
            Sampler rng = makeSeededRandomEngine(looper, index, 0, scene->sampleSequence);
            Ray ray = cam.sample(x, y, draw4(rng));
            Intersection intersec;
            scene->intersect(ray, intersec);
            if (intersec.primId == NullPrimitive) {
                if (scene->envMap != nullptr) direct = scene->envMap->linearSample(Math::toPlane(ray.direction));   /* pathtrace.cu:295-300 */
            } else {
                Material material = scene->getTexturedMaterialAndSurface(intersec);
                if (material.type == Material::Type::Light) direct = material.baseColor;
                else {
                    intersec.wo = -ray.direction;
                    bool deltaBSDF = (material.type == Material::Type::Dielectric);
                    if (!deltaBSDF && glm::dot(intersec.norm, intersec.wo) < 0.f) intersec.norm = -intersec.norm;
                    if (!deltaBSDF) {
                        glm::vec3 Li;
                        glm::vec3 wi;
                        glm::vec4 r4 = draw4(rng);
                        float lightPdf = scene->sampleDirectLight(intersec.pos, r4, Li, wi);
                        if (lightPdf > 0.f)
                            direct = Li * material.BSDF(intersec.norm, intersec.wo, wi) * Math::satDot(intersec.norm, wi) / lightPdf;
                    }
                }
            }
            f->radiance[index] = (f->radiance[index] * float(iter) + direct) / float(iter + 1);
        }
    }
}

/* ================================================================ ReSTIR GI: ReSTIRIndirectKernel (restir.cu:242-416) restated over the
 * reference's OWN Material::{sample, pdf, BSDF}, DevScene::{intersect, sampleDirectLight, environmentMapPdf, getPrimitiveArea,
 * getTexturedMaterialAndSurface}, Math::{powerHeuristic, pdfAreaToSolidAngle}, Reservoir<IndirectLiSample> and IndirectLiSample.
 * Same documented deviation as the port: a pixel whose jittered ray leaves the scene or hits an emitter writes indirect = 0
 * (the kernel would shade a history sample with uninitialised primMaterial / primWo there). */
using IndirectReservoir = Reservoir<IndirectLiSample>;
struct OrcGI {
    OrcFrame* f;
    std::vector<IndirectReservoir> resv, lastResv;
    std::vector<glm::vec3> indirect;
    std::vector<float> exportBuf;
    bool first = true;
};
static glm::vec3 draw3(Sampler& rng) { float a = sample1D(rng), b = sample1D(rng), c = sample1D(rng); return glm::vec3(a, b, c); }   /* nvcc order */

OrcGI* orc_gi_create(OrcFrame* f) {
    OrcGI* g = new OrcGI;
    size_t P = f->albedo.size();
    g->f = f; g->resv.assign(P, IndirectReservoir()); g->lastResv.assign(P, IndirectReservoir()); g->indirect.assign(P, glm::vec3(0.f));
    return g;
}
void orc_gi_destroy(OrcGI* g) { delete g; }

void orc_restir_indirect(OrcGI* g, const OrcCamera* ocam, int looper, int iter, int maxDepth, int reuseState) {
    OrcFrame* f = g->f;
    Camera cam = *(const Camera*)ocam;
    DevScene* scene = &f->sc->dev;
    GBuffer gBuffer = f->g;
    const bool first = g->first;
    IndirectReservoir* temporalReservoir = g->resv.data();
    IndirectReservoir* lastTemporalReservoir = g->lastResv.data();
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < cam.resolution.y; y++)
        for (int x = 0; x < cam.resolution.x; x++) {
            IndirectLiSample indirectSample;
            int index = y * cam.resolution.x + x;
            Sampler rng = makeSeededRandomEngine(looper, index, 0, scene->sampleSequence);
            Ray ray = cam.sample(x, y, draw4(rng));
            Intersection intersec;
            scene->intersect(ray, intersec);
            bool shaded = false, primSampleDelta = false;
            float primSamplePdf = 1.f;
            glm::vec3 primWo(0.f);
            Material primMaterial;
            if (intersec.primId != NullPrimitive) {
                Material material = scene->getTexturedMaterialAndSurface(intersec);
                if (material.type != Material::Type::Light) {
                    shaded = true;
                    glm::vec3 throughput(1.f);
                    intersec.wo = -ray.direction;
                    primWo = -ray.direction;
                    primMaterial = material;
                    for (int depth = 1; depth <= maxDepth; depth++) {
                        bool deltaBSDF = (material.type == Material::Type::Dielectric);
                        if (material.type != Material::Type::Dielectric && glm::dot(intersec.norm, intersec.wo) < 0.f) intersec.norm = -intersec.norm;
                        if (!deltaBSDF && depth > 1) {
                            glm::vec3 radiance;
                            glm::vec3 wi;
                            glm::vec4 r4 = draw4(rng);
                            float lightPdf = scene->sampleDirectLight(intersec.pos, r4, radiance, wi);
                            if (lightPdf > 0.f) {
                                float BSDFPdf = material.pdf(intersec.norm, intersec.wo, wi);
                                indirectSample.Lo += throughput * material.BSDF(intersec.norm, intersec.wo, wi) *
                                    radiance * Math::satDot(intersec.norm, wi) / lightPdf * Math::powerHeuristic(lightPdf, BSDFPdf);
                            }
                        }
                        BSDFSample sample;
                        sample.dir = glm::vec3(0.f); sample.bsdf = glm::vec3(0.f); sample.pdf = 0.f; sample.type = 0;
                        glm::vec3 r3 = draw3(rng);
                        material.sample(intersec.norm, intersec.wo, r3, sample);
                        if (sample.type == BSDFSampleType::Invalid) break;
                        else if (sample.pdf < 1e-8f) break;
                        bool deltaSample = (sample.type & BSDFSampleType::Specular);
                        if (depth > 1) throughput *= sample.bsdf / sample.pdf * (deltaSample ? 1.f : Math::absDot(intersec.norm, sample.dir));
                        else { primSamplePdf = sample.pdf; primSampleDelta = deltaSample; indirectSample.xv = intersec.pos; indirectSample.nv = intersec.norm; }
                        ray = makeOffsetedRay(intersec.pos, sample.dir);
                        glm::vec3 curPos = intersec.pos;
                        scene->intersect(ray, intersec);
                        intersec.wo = -ray.direction;
                        if (intersec.primId == NullPrimitive) {
                            if (scene->envMap != nullptr) {
                                glm::vec3 radiance = scene->envMap->linearSample(Math::toPlane(ray.direction)) * throughput;
                                float weight = deltaSample ? 1.f : Math::powerHeuristic(sample.pdf, scene->environmentMapPdf(ray.direction));
                                indirectSample.Lo += radiance * weight;
                            }
                            break;
                        }
                        material = scene->getTexturedMaterialAndSurface(intersec);
                        if (material.type == Material::Type::Light) {
                            if (glm::dot(intersec.norm, ray.direction) < 0.f) break;      /* SCENE_LIGHT_SINGLE_SIDED */
                            glm::vec3 radiance = material.baseColor;
                            float weight = (deltaSample || depth == 1) ? 1.f : Math::powerHeuristic(
                                sample.pdf,
                                Math::pdfAreaToSolidAngle(Math::luminance(radiance) * scene->sumLightPowerInv * scene->getPrimitiveArea(intersec.primId), curPos, intersec.pos, intersec.norm));
                            indirectSample.Lo += radiance * throughput * weight;
                            if (depth == 1) { indirectSample.xs = intersec.pos; indirectSample.ns = intersec.norm; }
                            break;
                        }
                        if (depth == 1) { indirectSample.xs = intersec.pos; indirectSample.ns = intersec.norm; }
                    }
                }
            }
            IndirectReservoir reservoir;
            float sampleWeight = 0.f;
            if (!indirectSample.invalid()) {
                sampleWeight = IndirectReservoir::toScalar(indirectSample.Lo / primSamplePdf);
                if (isnan(sampleWeight) || sampleWeight < 0.f) sampleWeight = 0.f;
            }
            reservoir.update(indirectSample, sampleWeight, sample1D(rng));
            if (!first && (reuseState & ReservoirReuse::Temporal)) {
                int primId = gBuffer.primId()[index];
                int lastIdx = gBuffer.devMotion[index];
                bool diff = false;
                if (lastIdx < 0) diff = true;
                else if (primId <= NullPrimitive) diff = true;
                else if (gBuffer.lastPrimId()[lastIdx] != primId) diff = true;
                else {
                    glm::vec3 norm = DECODE_NORM(gBuffer.normal()[index]);
                    glm::vec3 lastNorm = DECODE_NORM(gBuffer.lastNormal()[lastIdx]);
                    float depth = gBuffer.depth()[index];
                    float pdepth = gBuffer.lastDepth()[lastIdx];
                    if (Math::absDot(norm, lastNorm) < .9f || glm::abs(pdepth - depth) > depth * .1f) diff = true;
                }
                IndirectReservoir tempReservoir = diff ? IndirectReservoir() : lastTemporalReservoir[lastIdx];
                if (!tempReservoir.invalid()) reservoir.merge(tempReservoir, sample1D(rng));
            }
            glm::vec3 indirect(0.f);
            IndirectLiSample sample = reservoir.sample;
            reservoir.clamp<20>();
            if (shaded && !reservoir.invalid()) {
                glm::vec3 primWi = glm::normalize(sample.xs - sample.xv);
                indirect = reservoir.sample.Lo / IndirectReservoir::toScalar(reservoir.sample.Lo) * reservoir.weight / static_cast<float>(reservoir.numSamples);
                indirect *= primMaterial.BSDF(sample.nv, primWo, primWi) * (primSampleDelta ? 1.f : Math::satDot(sample.nv, primWi));
            }
            if (Math::hasNanOrInf(indirect)) indirect = glm::vec3(0.f);
            temporalReservoir[index] = reservoir;
            g->indirect[index] = (g->indirect[index] * float(iter) + indirect) / float(iter + 1);
        }
    std::swap(g->resv, g->lastResv);
    g->first = false;
}
const float* orc_gi_indirect(OrcGI* g) { return &g->indirect[0].x; }
const float* orc_gi_reservoirs(OrcGI* g) {
    const size_t P = g->lastResv.size();
    g->exportBuf.resize(P * 17);
    for (size_t i = 0; i < P; i++) {
        const IndirectReservoir& r = g->lastResv[i];
        const glm::vec3 v[5] = {r.sample.Lo, r.sample.xv, r.sample.nv, r.sample.xs, r.sample.ns};
        for (int k = 0; k < 5; k++) { g->exportBuf[i * 17 + 3 * k] = v[k].x; g->exportBuf[i * 17 + 3 * k + 1] = v[k].y; g->exportBuf[i * 17 + 3 * k + 2] = v[k].z; }
        g->exportBuf[i * 17 + 15] = (float)r.numSamples; g->exportBuf[i * 17 + 16] = r.weight;
    }
    return g->exportBuf.data();
}

const void* orc_frame_buffer(OrcFrame* f, int which) {
    const int cur = f->g.frameIdx;
    switch (which) {
    case ORC_BUF_ALBEDO: return f->albedo.data();
    case ORC_BUF_NORMAL: return f->normal[cur].data();
    case ORC_BUF_MATID: return f->matId[cur].data();
    case ORC_BUF_DEPTH: return f->depth[cur].data();
    case ORC_BUF_MOTION: return f->motion.data();
    case ORC_BUF_RADIANCE: return f->radiance.data();
    case ORC_BUF_RESERVOIR: return f->lastResv.data();
    case ORC_BUF_RESERVOIR_TEMP: return f->temp.data();
    }
    return nullptr;     /* ORC_BUF_LIGHT_INDEX: the reference keeps no light index (restir.h:7-11) */
}

void ref_probe(OrcScene* sc, float* out) {
    DevScene* scene = &sc->dev;
    glm::vec3 Li(0.f), wi(0.f);
    float dist = 0.f;
    glm::vec3 pos(0.1f, 0.5f, 0.2f);
    float p = scene->sampleDirectLightNoVisibility(pos, glm::vec4(0.3f, 0.6f, 0.2f, 0.7f), Li, wi, dist);
    Material m = scene->materials[0];
    m.baseColor = glm::vec3(1.f);
    glm::vec3 n(0.f, 1.f, 0.f), wo(0.f, 1.f, 0.f);
    glm::vec3 g = Li * m.BSDF(n, wo, wi) * Math::satDot(n, wi);
    float w = Math::luminance(g / p);
    out[0] = p; out[1] = Li.x; out[2] = wi.y; out[3] = dist; out[4] = scene->sumLightPowerInv; out[5] = (float)scene->lightSampler.length;
    out[6] = w; out[7] = g.x; out[8] = m.BSDF(n, wo, wi).x; out[9] = Math::satDot(n, wi); out[10] = (float)sizeof(DevScene); out[11] = (float)m.type;
}

float orc_alias_build(int n, const float* values, void* outTable) {
    DiscreteSampler1D<float> smp(std::vector<float>(values, values + n));
    memcpy(outTable, smp.binomDistribs.data(), sizeof(BinomialDistrib<float>) * n);
    return smp.sumAll;
}

void orc_rng_draws(int looper, int index, int n, float* out) {
    Sampler rng = makeSeededRandomEngine(looper, index, 0, nullptr);
    for (int i = 0; i < n; i++) out[i] = sample1D(rng);
}

int orc_intersect(const OrcScene* sc, const float* o, const float* d, float* out8, int* outMatId) {
    Ray ray; ray.origin = glm::vec3(o[0], o[1], o[2]); ray.direction = glm::vec3(d[0], d[1], d[2]);
    Intersection is;
    ((OrcScene*)sc)->dev.intersect(ray, is);
    if (is.primId != NullPrimitive) {
        out8[0] = is.pos.x; out8[1] = is.pos.y; out8[2] = is.pos.z;
        out8[3] = is.norm.x; out8[4] = is.norm.y; out8[5] = is.norm.z;
        out8[6] = is.uv.x; out8[7] = is.uv.y;
        *outMatId = is.matId;
    }
    return is.primId;
}
int orc_occluded(const OrcScene* sc, const float* x, const float* y) {
    return ((OrcScene*)sc)->dev.testOcclusion(glm::vec3(x[0], x[1], x[2]), glm::vec3(y[0], y[1], y[2])) ? 1 : 0;
}

} // extern "C"
