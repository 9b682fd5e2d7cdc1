// capi.cu -- the C ABI (include/restir_b200.h) over the host scene builder and the sm_100a kernels.
#include "capi_internal.h"

using namespace rs;

static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};

int rsFail(int code, const std::string& msg) { g_err = msg; return code; }
void rsCountLaunches(int n) { g_launches += n; }
static int fail(int code, const std::string& msg) { return rsFail(code, msg); }

int rsSmCount();
static int smCount() { return rsSmCount(); }
int rsSmCount() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

CamDev rsToCamDev(const RstrCamera& c) {
    CamDev d;
    memcpy(d.position, c.position, 12); memcpy(d.right, c.right, 12); memcpy(d.up, c.up, 12); memcpy(d.view, c.view, 12);
    memcpy(d.rotInv, c.rotationMatInv, 36);
    d.aspect = (float)c.resolution[0] / c.resolution[1];                 // sceneStructs.h:71
    d.tanFovY = tanf(radians(c.fov[1]));                                 // sceneStructs.h:72 (evaluated once on the host)
    d.focalDist = c.focalDist;
    d.pixelSizeX = 1.f / (float)c.resolution[0];                         // sceneStructs.h:73
    d.pixelSizeY = 1.f / (float)c.resolution[1];
    d.resX = (float)c.resolution[0]; d.resY = (float)c.resolution[1];
    return d;
}

FrameDev rsToFrameDev(const RstrFrame* f, int rowLo, int rowHi) {
    FrameDev d{};
    d.W = f->W; d.H = f->H; d.rowLo = rowLo; d.rowHi = rowHi; d.bufRow0 = f->bufRow0; d.bufRows = f->bufRows;
    d.geom[0] = f->geom[f->cur]; d.geom[1] = f->geom[f->cur ^ 1];
    d.matId[0] = f->matId[f->cur]; d.matId[1] = f->matId[f->cur ^ 1];
    d.albedoMotion = f->albedoMotion; d.radiance = f->radiance;
    d.resvOut = f->resv[f->resvOut]; d.resvIn = f->resv[f->resvOut ^ 1]; d.resvTemp = f->resvTemp;
    d.hit = f->hit; d.hitPos = f->hitPos; d.hitMR = f->hitMR; d.rowCost = f->rowCost; d.haloMiss = f->haloMiss; d.motionRows = f->haloMiss + 1; d.queue = f->queue; d.queueCount = f->queueCount; d.shadeQueue = f->shadeQueue;
    return d;
}

template <typename T>
static cudaError_t upload(void** dst, const std::vector<T>& v, size_t& total) {
    size_t bytes = v.size() * sizeof(T);
    if (bytes == 0) bytes = 64;
    cudaError_t e = cudaMalloc(dst, bytes);
    if (e != cudaSuccess) return e;
    total += bytes;
    if (!v.empty()) e = cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    return e;
}

// host build only; the device copy is made by ensureUploaded() when the first frame is created, so that
// scene construction / inspection (rstr_scene_read, rstr_scene_info) also work on a machine without a GPU
static int finishScene(RstrScene* sc, RstrScene** out) {
    std::string err;
    if (!buildHostScene(sc->hs, err)) { delete sc; return fail(RSTR_ERR_ARG, err); }
    if (sc->hs.bvhDepth > RS_STACK_DEPTH) {
        char b[160];
        snprintf(b, sizeof b, "BVH depth %d exceeds the traversal stack (%d); coincident centroids? (SURVEY App. C4)", sc->hs.bvhDepth, RS_STACK_DEPTH);
        delete sc;
        return fail(RSTR_ERR_LIMIT, b);
    }
    if (sc->hs.fastDepth > RS_PACKET_STACK) {
        char b[160];
        snprintf(b, sizeof b, "traced tree depth %d exceeds the packet traversal stack (%d)", sc->hs.fastDepth, RS_PACKET_STACK);
        delete sc;
        return fail(RSTR_ERR_LIMIT, b);
    }
    *out = sc;
    return RSTR_OK;
}

static void freeSceneDevice(RstrScene* sc) {
    void** all[] = {&sc->dNodes, &sc->dTriGeom, &sc->dTriNorm, &sc->dFastNodes, &sc->dPrimToFast, &sc->dFallback, &sc->dRank, &sc->dMaterials,
                    &sc->dAlias, &sc->dLights, &sc->dTexData, &sc->dTexInfo, &sc->dTriUV, &sc->dEnvAlias, &sc->dEnvDir};
    for (void** p : all) { cudaFree(*p); *p = nullptr; }
    sc->uploaded = false;
    sc->deviceBytes = 0;
}

static int ensureUploaded(RstrScene* sc);
int rsEnsureUploaded(RstrScene* sc) { return ensureUploaded(sc); }
static int ensureUploaded(RstrScene* sc) {
    if (sc->uploaded) return RSTR_OK;
    HostScene& hs = sc->hs;
    size_t total = 0;
    cudaError_t e;
    if ((e = upload(&sc->dNodes, hs.packed, total)) != cudaSuccess || (e = upload(&sc->dTriGeom, hs.fastTris, total)) != cudaSuccess ||
        (e = upload(&sc->dFastNodes, hs.fastNodes, total)) != cudaSuccess || (e = upload(&sc->dRank, hs.rank, total)) != cudaSuccess || (e = upload(&sc->dPrimToFast, hs.primToFast, total)) != cudaSuccess ||

        (e = upload(&sc->dTriNorm, hs.triNorm, total)) != cudaSuccess || (e = upload(&sc->dMaterials, hs.materials, total)) != cudaSuccess ||
        (e = upload(&sc->dAlias, hs.alias, total)) != cudaSuccess || (e = upload(&sc->dLights, hs.lights, total)) != cudaSuccess ||
        (e = upload(&sc->dTexData, hs.texData, total)) != cudaSuccess || (e = upload(&sc->dTexInfo, hs.texInfo, total)) != cudaSuccess ||
        (e = upload(&sc->dTriUV, hs.triUV, total)) != cudaSuccess || (e = upload(&sc->dEnvAlias, hs.envAlias, total)) != cudaSuccess ||
        (e = upload(&sc->dEnvDir, hs.envDir, total)) != cudaSuccess) {
        std::string m = std::string("scene upload failed: ") + cudaGetErrorString(e);
        freeSceneDevice(sc);
        return fail(RSTR_ERR_CUDA, m);
    }
    if (cudaMalloc(&sc->dFallback, 4 * sizeof(unsigned int)) != cudaSuccess || cudaMemset(sc->dFallback, 0, 4 * sizeof(unsigned int)) != cudaSuccess) {
        freeSceneDevice(sc);
        return fail(RSTR_ERR_CUDA, "scene upload failed: counter");
    }
    sc->deviceBytes = total;
    DevScene& d = sc->dev;
    d.fallbackRays = (unsigned int*)sc->dFallback;
    d.nodes = (const float4*)sc->dNodes; d.triGeom = (const float4*)sc->dTriGeom; d.triNorm = (const float4*)sc->dTriNorm;
    d.materials = (const RstrMaterial*)sc->dMaterials; d.alias = (const float2*)sc->dAlias; d.lights = (const float4*)sc->dLights;
    d.fastNodes = (const float4*)sc->dFastNodes; d.primToFast = (const int*)sc->dPrimToFast; d.rank = (const int*)sc->dRank;
    d.numTris = hs.T; d.numFastNodes = (int)hs.fastNodes.size(); d.fastRoot = hs.fastRoot; d.traversal = sc->traversalMode;
    memcpy(d.fastRootMin, hs.fastRootMin, 12); memcpy(d.fastRootMax, hs.fastRootMax, 12);
    d.numLights = (int)hs.alias.size();           // lightSampler.length: emissive triangles (+ the environment map, last)
    d.texData = (const float4*)sc->dTexData; d.texInfo = (const int4*)sc->dTexInfo; d.triUV = (const float4*)sc->dTriUV;
    d.anyMaps = hs.anyMaps ? 1 : 0;
    d.envTex = hs.envMapTexId; d.envLen = (int)hs.envAlias.size();
    d.envAlias = (const float2*)sc->dEnvAlias; d.envDir = (const float4*)sc->dEnvDir;
    d.sumLightPowerInv = hs.sumLightPowerInv;
    d.rootRef = hs.rootRef;
    memcpy(d.rootMin, &hs.rootBox.pMin, 12); memcpy(d.rootMax, &hs.rootBox.pMax, 12);
    sc->uploaded = true;
    return RSTR_OK;
}

extern "C" {

const char* rstr_last_error(void) { return g_err.c_str(); }
#ifndef RSTR_BUILD_ID
#define RSTR_BUILD_ID "unknown"
#endif
const char* rstr_build_id(void) { return RSTR_BUILD_ID; }

int rstr_init(int device) {
    int n = 0;
    CU(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail(RSTR_ERR_ARG, "device index out of range");
    CU(cudaSetDevice(device));
    CU(cudaFree(0));
    return RSTR_OK;
}

void rstr_params_default(RstrParams* p) {
    p->numCandidates = 32; p->temporalCap = 20; p->numSpatial = 5; p->spatialRadius = 5.f; p->reuse = RSTR_REUSE_TEMPORAL;   // common.cpp:14
    p->spatialPasses = 1;
    p->unbiased = 0;
}

int rstr_scene_create(const RstrSceneDesc* desc, RstrScene** out) {
    if (!desc || !out || desc->numTris <= 0 || !desc->vertices || !desc->normals || !desc->materialIds || !desc->materials || desc->numMaterials <= 0)
        return fail(RSTR_ERR_ARG, "rstr_scene_create: bad descriptor");
    RstrScene* sc = new RstrScene;
    HostScene& hs = sc->hs;
    const int T = desc->numTris;
    hs.T = T;
    hs.vertices.resize(3 * (size_t)T); hs.normals.resize(3 * (size_t)T); hs.texcoords.assign(6 * (size_t)T, 0.f);
    memcpy(hs.vertices.data(), desc->vertices, 36 * (size_t)T);
    memcpy(hs.normals.data(), desc->normals, 36 * (size_t)T);
    if (desc->texcoords) memcpy(hs.texcoords.data(), desc->texcoords, 24 * (size_t)T);
    hs.materialIds.assign(desc->materialIds, desc->materialIds + T);
    hs.materials.assign(desc->materials, desc->materials + desc->numMaterials);
    if (desc->numTextures < 0 || (desc->numTextures > 0 && !desc->textures) || desc->envMap < 0 || desc->envMap > desc->numTextures) {
        delete sc;
        return fail(RSTR_ERR_ARG, "rstr_scene_create: bad texture list / environment map index");
    }
    hs.textures.resize(desc->numTextures);
    for (int t = 0; t < desc->numTextures; t++) {
        const RstrTexture& src = desc->textures[t];
        if (src.width <= 0 || src.height <= 0 || !src.rgb) { delete sc; return fail(RSTR_ERR_ARG, "rstr_scene_create: empty texture"); }
        hs.textures[t].w = src.width; hs.textures[t].h = src.height;
        hs.textures[t].rgb.resize((size_t)src.width * src.height);
        memcpy(hs.textures[t].rgb.data(), src.rgb, 12 * hs.textures[t].rgb.size());
    }
    hs.envMapTexId = desc->envMap - 1;
    return finishScene(sc, out);
}

int rstr_scene_load_file(const char* path, RstrScene** out, RstrCamera* cameraOut) {
    if (!path || !out) return fail(RSTR_ERR_ARG, "rstr_scene_load_file: bad argument");
    RstrScene* sc = new RstrScene;
    RstrCamera cam;
    memset(&cam, 0, sizeof cam);
    std::string err;
    try {
        if (!loadSceneFile(path, sc->hs, cam, err)) { delete sc; return fail(RSTR_ERR_IO, err); }
    } catch (const std::exception& e) {
        delete sc;
        return fail(RSTR_ERR_IO, std::string("rstr_scene_load_file: ") + e.what());
    }
    if (cameraOut) *cameraOut = cam;
    return finishScene(sc, out);
}

int rstr_scene_destroy(RstrScene* sc) {
    if (!sc) return RSTR_OK;
    freeSceneDevice(sc);
    delete sc;
    return RSTR_OK;
}

int rstr_scene_set_traversal(RstrScene* sc, int mode) {
    if (!sc || (mode != RS_TRAVERSAL_FAST && mode != RS_TRAVERSAL_EXACT)) return fail(RSTR_ERR_ARG, "rstr_scene_set_traversal: bad argument");
    sc->traversalMode = mode;
    sc->dev.traversal = mode;
    return RSTR_OK;
}

int rstr_scene_fallback_rays(RstrScene* sc, unsigned long long* count, int reset) {
    if (!sc || !count) return fail(RSTR_ERR_ARG, "rstr_scene_fallback_rays: bad argument");
    unsigned int v[4] = {0, 0, 0, 0};
    if (sc->dFallback) {
        CU(cudaDeviceSynchronize());
        CU(cudaMemcpy(v, sc->dFallback, sizeof v, cudaMemcpyDeviceToHost));
        if (reset) CU(cudaMemset(sc->dFallback, 0, sizeof v));
    }
    for (int i = 0; i < 4; i++) count[i] = v[i];
    return RSTR_OK;
}

int rstr_scene_info(const RstrScene* sc, RstrSceneInfo* info) {
    if (!sc || !info) return fail(RSTR_ERR_ARG, "rstr_scene_info: bad argument");
    info->numTris = sc->hs.T; info->numLights = (int)sc->hs.alias.size(); info->bvhSize = sc->hs.bvhSize;
    info->bvhDepth = sc->hs.bvhDepth; info->numMaterials = (int)sc->hs.materials.size(); info->sumLightPower = sc->hs.sumAll;
    info->buildSeconds = sc->hs.buildSeconds; info->deviceBytes = sc->deviceBytes;
    info->tracedBvhDepth = sc->hs.fastDepth; info->tracedBuildSeconds = sc->hs.fastBuildSeconds;
    info->tracedNodes = (int)sc->hs.fastNodes.size(); info->tracedRoot = sc->hs.fastRoot;
    info->numEmissiveTris = (int)sc->hs.lightPrimIds.size(); info->numTextures = (int)sc->hs.textures.size();
    info->envWidth = info->envHeight = 0;
    if (sc->hs.envMapTexId >= 0) { info->envWidth = sc->hs.textures[sc->hs.envMapTexId].w; info->envHeight = sc->hs.textures[sc->hs.envMapTexId].h; }
    return RSTR_OK;
}

int rstr_scene_texture_info(const RstrScene* sc, int index, int* width, int* height, int* isEnvMap) {
    if (!sc || index < 0 || index >= (int)sc->hs.textures.size()) return fail(RSTR_ERR_ARG, "rstr_scene_texture_info: bad argument");
    if (width) *width = sc->hs.textures[index].w;
    if (height) *height = sc->hs.textures[index].h;
    if (isEnvMap) *isEnvMap = sc->hs.envMapTexId == index;
    return RSTR_OK;
}

int rstr_scene_read(const RstrScene* sc, int which, void* host, size_t bytes) {
    if (!sc || !host) return fail(RSTR_ERR_ARG, "rstr_scene_read: bad argument");
    const HostScene& hs = sc->hs;
    const void* src = nullptr;
    size_t need = 0;
    std::vector<MTNode> mt;
    switch (which) {
    case RSTR_SCENE_BOXES: src = hs.boxes.data(); need = hs.boxes.size() * sizeof(Box); break;
    case RSTR_SCENE_LIGHT_PRIM_IDS: src = hs.lightPrimIds.data(); need = hs.lightPrimIds.size() * 4; break;
    case RSTR_SCENE_LIGHT_RADIANCE: src = hs.lightUnitRadiance.data(); need = hs.lightUnitRadiance.size() * 12; break;
    case RSTR_SCENE_ALIAS: src = hs.alias.data(); need = hs.alias.size() * 8; break;
    case RSTR_SCENE_VERTICES: src = hs.vertices.data(); need = hs.vertices.size() * 12; break;
    case RSTR_SCENE_NORMALS: src = hs.normals.data(); need = hs.normals.size() * 12; break;
    case RSTR_SCENE_TEXCOORDS: src = hs.texcoords.data(); need = hs.texcoords.size() * 4; break;
    case RSTR_SCENE_MATERIAL_IDS: src = hs.materialIds.data(); need = hs.materialIds.size() * 4; break;
    case RSTR_SCENE_MATERIALS: src = hs.materials.data(); need = hs.materials.size() * sizeof(RstrMaterial); break;
    case RSTR_SCENE_ENV_ALIAS: src = hs.envAlias.data(); need = hs.envAlias.size() * 8; break;
    case RSTR_SCENE_TRACED_NODES: src = hs.fastNodes.data(); need = hs.fastNodes.size() * sizeof(FastNode); break;
    case RSTR_SCENE_TRACED_TRIS: src = hs.fastTris.data(); need = hs.fastTris.size() * sizeof(TriGeom); break;
    default:
        if (which >= RSTR_SCENE_TEXTURE0 && which < RSTR_SCENE_TEXTURE0 + (int)hs.textures.size()) {
            src = hs.textures[which - RSTR_SCENE_TEXTURE0].rgb.data(); need = hs.textures[which - RSTR_SCENE_TEXTURE0].rgb.size() * 12;
        } else if (which >= RSTR_SCENE_MTBVH0 && which < RSTR_SCENE_MTBVH0 + 6) {
            exportMTBVH(hs, which - RSTR_SCENE_MTBVH0, mt);
            src = mt.data(); need = mt.size() * sizeof(MTNode);
        } else return fail(RSTR_ERR_ARG, "rstr_scene_read: unknown array");
    }
    if (bytes != need) return fail(RSTR_ERR_ARG, "rstr_scene_read: size mismatch");
    if (need) memcpy(host, src, need);
    return RSTR_OK;
}

// runCuda()'s camera animation (main.cpp:149-153, 160) with the clock replaced by t = frame / fps * speed
int rstr_camera_orbit(const RstrCamera* base, int frame, float speed, float radius, float fps, RstrCamera* out) {
    if (!base || !out) return fail(RSTR_ERR_ARG, "rstr_camera_orbit: null");
    *out = *base;
    float t = (float)frame / fps * speed;
    out->position[0] = base->position[0] + cosf(t) * radius;
    out->position[1] = base->position[1] + 0.f * radius;
    out->position[2] = base->position[2] + sinf(t) * radius;
    cameraUpdate(*out);
    return RSTR_OK;
}

int rstr_camera_update(RstrCamera* c) {
    if (!c) return fail(RSTR_ERR_ARG, "rstr_camera_update: null");
    cameraUpdate(*c);
    return RSTR_OK;
}

int rstr_frame_destroy(RstrFrame* f) {
    if (!f) return RSTR_OK;
    if (f->stream) cudaStreamSynchronize(f->stream);
    cudaFree(f->slab);             // geom[2], matId[2], resv[2], resvTemp, resvTemp2 live in the exchange slab
    cudaFree(f->albedoMotion); cudaFree(f->radiance); cudaFree(f->hit); cudaFree(f->hitPos); cudaFree(f->hitMR); cudaFree(f->rowCost); cudaFree(f->ldr);
    cudaFree(f->haloMiss); cudaFree(f->scratch); cudaFree(f->queue); cudaFree(f->shadeQueue); cudaFree(f->queueCount);
    for (int i = 0; i < RSTR_LDR_SLOTS; i++) {
        cudaFree(f->ldrB[i]);
        if (f->evRendered[i]) cudaEventDestroy(f->evRendered[i]);
        if (f->evCopied[i]) cudaEventDestroy(f->evCopied[i]);
    }
    if (f->copyStream) cudaStreamDestroy(f->copyStream);
    for (int i = 0; i < 2; i++) if (f->ss.side[i]) { cudaStreamSynchronize(f->ss.side[i]); cudaStreamDestroy(f->ss.side[i]); }
    for (int i = 0; i < RS_MAX_BANDS; i++) { if (f->ss.evPrimary[i]) cudaEventDestroy(f->ss.evPrimary[i]); if (f->ss.evDone[i]) cudaEventDestroy(f->ss.evDone[i]); }
    for (auto& e : f->ev) if (e) cudaEventDestroy(e);
    if (f->xfer) cudaEventDestroy(f->xfer);
    for (auto& e : f->marks) if (e) cudaEventDestroy(e);
    if (f->stream && f->ownStream) cudaStreamDestroy(f->stream);
    delete f;
    return RSTR_OK;
}

int rstr_frame_create_strip(RstrScene* sc, int W, int H, int row0, int row1, int halo, RstrFrame** out) {
    if (!sc || !out || W <= 0 || H <= 0 || row0 < 0 || row1 > H || row0 >= row1 || halo < 0)
        return fail(RSTR_ERR_ARG, "rstr_frame_create: bad argument");
    int urc = ensureUploaded(sc);
    if (urc) return urc;
    RstrFrame* f = new RstrFrame;
    f->sc = sc; f->W = W; f->H = H; f->row0 = row0; f->row1 = row1; f->halo = halo;
    f->bufRow0 = row0 - halo < 0 ? 0 : row0 - halo;
    int bufRow1 = row1 + halo > H ? H : row1 + halo;
    f->bufRows = bufRow1 - f->bufRow0;
    f->nBuf = (size_t)f->bufRows * W;
    const size_t n = f->nBuf;
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void** p, size_t bytes) {
        if (e != cudaSuccess) return;
        e = cudaMalloc(p, bytes);
        if (e == cudaSuccess) e = cudaMemset(*p, 0, bytes);      // restir.cu:482-489 zero-fills the reservoirs
    };
    {
        // exchange slab: header of flags, then the planes a neighbouring strip reads, each 256-byte aligned
        const size_t planeBytes[RS_XP_COUNT] = {n * sizeof(float4), n * sizeof(float4), n * sizeof(int), n * sizeof(int),
                                                n * sizeof(ResvD), n * sizeof(ResvD), n * sizeof(ResvD), n * sizeof(ResvD)};
        size_t off = RS_SLAB_HEADER;
        for (int i = 0; i < RS_XP_COUNT; i++) { f->slabOff[i] = off; off += (planeBytes[i] + 255) & ~(size_t)255; }
        f->slabBytes = off;
        alloc(&f->slab, f->slabBytes);
        if (e == cudaSuccess) {
            char* b = (char*)f->slab;
            f->geom[0] = (float4*)(b + f->slabOff[RS_XP_GEOM0]); f->geom[1] = (float4*)(b + f->slabOff[RS_XP_GEOM1]);
            f->matId[0] = (int*)(b + f->slabOff[RS_XP_MATID0]); f->matId[1] = (int*)(b + f->slabOff[RS_XP_MATID1]);
            f->resv[0] = (ResvD*)(b + f->slabOff[RS_XP_RESV0]); f->resv[1] = (ResvD*)(b + f->slabOff[RS_XP_RESV1]);
            f->resvTemp = (ResvD*)(b + f->slabOff[RS_XP_TEMP]); f->resvTemp2 = (ResvD*)(b + f->slabOff[RS_XP_TEMP2]);
        }
    }
    alloc((void**)&f->albedoMotion, n * sizeof(float4));
    alloc((void**)&f->radiance, n * 3 * sizeof(float));
    alloc((void**)&f->hit, n * sizeof(HitRec));
    if (sc->hs.anyMRMaps) alloc((void**)&f->hitMR, n * sizeof(float2));
    alloc((void**)&f->ldr, n * sizeof(uchar4));
    alloc((void**)&f->haloMiss, 4 * sizeof(unsigned int));     // [0] halo misses, [1] max |row(motion) - row|
    alloc((void**)&f->queue, n * sizeof(int));
    alloc((void**)&f->shadeQueue, n * sizeof(int));
    alloc((void**)&f->queueCount, (4 + 2 * RS_MAX_BANDS) * sizeof(unsigned int));
    if (e == cudaSuccess) {
        // a zero-filled reference reservoir has no sample: lightId must read as "none"
        std::vector<ResvD> init(n);
        memset(init.data(), 0, n * sizeof(ResvD));
        for (size_t i = 0; i < n; i++) init[i].lightId = -1;
        for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaMemcpy(f->resv[i], init.data(), n * sizeof(ResvD), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(f->resvTemp, init.data(), n * sizeof(ResvD), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking);
    for (auto& ev : f->ev) if (e == cudaSuccess) e = cudaEventCreate(&ev);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&f->xfer, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        std::string m = std::string("rstr_frame_create: ") + cudaGetErrorString(e);
        rstr_frame_destroy(f);
        return fail(RSTR_ERR_CUDA, m);
    }
    *out = f;
    return RSTR_OK;
}

int rstr_frame_create(RstrScene* sc, int W, int H, RstrFrame** out) { return rstr_frame_create_strip(sc, W, H, 0, H, 0, out); }

int rstr_frame_reset(RstrFrame* f) {
    if (!f) return fail(RSTR_ERR_ARG, "null frame");
    f->first = true;
    return RSTR_OK;
}

static int checkCam(const RstrFrame* f, const RstrCamera* cam) {
    if (!f || !cam) return fail(RSTR_ERR_ARG, "null frame/camera");
    if (cam->resolution[0] != f->W || cam->resolution[1] != f->H) return fail(RSTR_ERR_ARG, "camera resolution differs from the frame");
    return RSTR_OK;
}

static inline void stageBegin(RstrFrame* f, int s) { cudaEventRecord(f->ev[2 * s], f->stream); f->ran[s] = true; }
static inline void stageEnd(RstrFrame* f, int s) { cudaEventRecord(f->ev[2 * s + 1], f->stream); }

// launches a G-buffer render that rstr_gbuffer_render deferred (see there)
int rsFlushGBuffer(RstrFrame* f) {
    if (!f->gbufPending) return RSTR_OK;
    f->gbufPending = false;
    FrameDev d = f->renderHalo ? rsToFrameDev(f, f->bufRow0, f->bufRow0 + f->bufRows)    // halo rows rendered locally ...
                               : rsToFrameDev(f, f->row0, f->row1);                     // ... or exchanged by the caller
    stageBegin(f, RSTR_T_GBUFFER);
    g_launches += launchGBuffer(f->sc->dev, d, f->pendC, f->pendLC, f->stream);
    stageEnd(f, RSTR_T_GBUFFER);
    CU(cudaGetLastError());
    return RSTR_OK;
}

int rstr_gbuffer_render(RstrFrame* f, const RstrCamera* cam) {
    int rc = checkCam(f, cam);
    if (rc) return rc;
    if ((rc = rsFlushGBuffer(f))) return rc;
    // the reference reads an uninitialised lastCamera before the first GBuffer::update (gbuffer.h:56); use cam
    f->pendCam = *cam;
    f->pendC = rsToCamDev(*cam); f->pendLC = rsToCamDev(f->haveLast ? f->lastCamera : *cam);
    for (bool& r : f->ran) r = false;
    f->gbufPending = true;
    // The pixel's two primary rays (centre ray here, jittered ray in rstr_restir_direct) share one tree walk when the
    // next call on this frame is phase A with the same camera: the launch is deferred until then.  Anything else that
    // looks at the G-buffer first (reads, halo exchange, gbuffer_update, sync, PTDirect) launches the plain kernel.
    const bool sameRows = !f->renderHalo || f->bufRows == f->row1 - f->row0;
    if (!f->fuse || !sameRows || f->sc->dev.traversal != RS_TRAVERSAL_FAST) return rsFlushGBuffer(f);
    return RSTR_OK;
}

// G-buffer + phase A of a frame: 1 = staged pipeline (k_primary, k_candidates, k_shadow, k_temporal), 0 = one fused kernel,
// -1 (default) = staged unless the scene is tiny
int rstr_frame_set_pipeline(RstrFrame* f, int staged) {
    if (!f) return fail(RSTR_ERR_ARG, "null frame");
    int rc = rsFlushGBuffer(f);
    f->staged = staged < 0 ? -1 : (staged != 0);
    return rc;
}

// number of row bands the staged pipeline overlaps (1 = no overlap, the default)
int rstr_frame_set_bands(RstrFrame* f, int bands) {
    if (!f || bands < 1 || bands > RS_MAX_BANDS) return fail(RSTR_ERR_ARG, "rstr_frame_set_bands: 1 .. 8");
    f->bands = bands;
    return RSTR_OK;
}

int rstr_frame_set_fusion(RstrFrame* f, int enable) {
    if (!f) return fail(RSTR_ERR_ARG, "null frame");
    int rc = rsFlushGBuffer(f);
    f->fuse = enable != 0;
    return rc;
}

int rstr_gbuffer_update(RstrFrame* f, const RstrCamera* cam) {
    if (!f || !cam) return fail(RSTR_ERR_ARG, "null frame/camera");
    int rc = rsFlushGBuffer(f);
    if (rc) return rc;
    f->lastCamera = *cam; f->haveLast = true; f->cur ^= 1;                // gbuffer.cu:75-78
    return RSTR_OK;
}

int rstr_restir_phase_a(RstrFrame* f, const RstrCamera* cam, const RstrParams* prm, int looper, int iter) {
    int rc = checkCam(f, cam);
    if (rc) return rc;
    if (!prm || prm->numCandidates < 0 || prm->numSpatial < 0 || prm->temporalCap < 1) return fail(RSTR_ERR_ARG, "bad RstrParams");
    if (prm->unbiased) {
        if (f->sc->hs.envMapTexId >= 0) return fail(RSTR_ERR_ARG, "RstrParams::unbiased: environment-map lights are not supported (triangle lights only)");
        if (f->bufRows != f->H) return fail(RSTR_ERR_ARG, "RstrParams::unbiased: strip frames are not supported (single-GPU frames only)");
        if (!f->hitPos) {
            CU(cudaMalloc((void**)&f->hitPos, f->nBuf * sizeof(float4)));
            CU(cudaMemsetAsync(f->hitPos, 0, f->nBuf * sizeof(float4), f->stream));
        }
    }
    FrameDev d = rsToFrameDev(f, f->row0, f->row1);
    d.resvStage = f->resvTemp;
    if (f->gbufPending && memcmp(cam, &f->pendCam, sizeof(RstrCamera)) == 0 && f->sc->dev.traversal == RS_TRAVERSAL_FAST) {
        stageBegin(f, RSTR_T_RIS);
        // auto: the staged pipeline pays for its extra launches and plane traffic once rays diverge, i.e. on real scenes; a
        // Cornell box (a few dozen triangles, every lane in step) is 9 % faster in the single fused kernel
        const bool staged = f->staged < 0 ? f->sc->hs.fastNodes.size() > 1024 : f->staged != 0;
        if (staged) {
            // the reservoir between k_candidates and k_temporal: resvTemp gets its final value from k_temporal when spatial reuse is
            // on; with spatial reuse off the reference leaves reservoirTemp alone (restir.cu:190), so the free second buffer is used
            d.resvStage = (prm->reuse & 2) ? f->resvTemp : f->resvTemp2;
            if (!(prm->reuse & 2)) f->temp2Ready = false;
        }
        if (staged && !f->ss.side[0]) {
            for (int i = 0; i < 2; i++) CU(cudaStreamCreateWithFlags(&f->ss.side[i], cudaStreamNonBlocking));
            for (int i = 0; i < RS_MAX_BANDS; i++) {
                CU(cudaEventCreateWithFlags(&f->ss.evPrimary[i], cudaEventDisableTiming));
                CU(cudaEventCreateWithFlags(&f->ss.evDone[i], cudaEventDisableTiming));
            }
        }
        f->ss.bands = f->bands;
        int n = staged ? launchPhaseAStaged(f->sc->dev, d, f->pendC, f->pendLC, *prm, looper, iter, f->first ? 1 : 0, smCount(), f->stream, f->ss)
                       : launchGBufferRestirA(f->sc->dev, d, f->pendC, f->pendLC, *prm, looper, iter, f->first ? 1 : 0, f->stream);
        if (n > 0) {
            f->gbufPending = false;
            g_launches += n;
            stageEnd(f, RSTR_T_RIS);
            CU(cudaGetLastError());
            return RSTR_OK;
        }
    }
    if ((rc = rsFlushGBuffer(f))) return rc;
    stageBegin(f, RSTR_T_RIS);
    g_launches += launchRestirA(f->sc->dev, d, rsToCamDev(*cam), *prm, looper, iter, f->first ? 1 : 0, f->stream);
    stageEnd(f, RSTR_T_RIS);
    CU(cudaGetLastError());
    return RSTR_OK;
}

// first use of the second publication buffer: it starts as a copy of resvTemp (what the reference's single buffer holds)
int rsEnsureTemp2(RstrFrame* f) {
    if (f->temp2Ready) return RSTR_OK;
    CU(cudaMemcpyAsync(f->resvTemp2, f->resvTemp, f->nBuf * sizeof(ResvD), cudaMemcpyDeviceToDevice, f->stream));
    f->temp2Ready = true;
    return RSTR_OK;
}

int rstr_restir_phase_b_pass(RstrFrame* f, const RstrCamera* cam, const RstrParams* prm, int looper, int iter, int pass) {
    int rc = checkCam(f, cam);
    if (rc) return rc;
    (void)looper;
    if (!prm || prm->numCandidates < 0 || prm->numSpatial < 0 || prm->temporalCap < 1) return fail(RSTR_ERR_ARG, "bad RstrParams");
    const int passes = (prm->reuse & 2) ? (prm->spatialPasses < 1 ? 1 : prm->spatialPasses) : 0;
    if ((passes && (pass < 1 || pass > passes)) || (!passes && pass != 1)) return fail(RSTR_ERR_ARG, "rstr_restir_phase_b_pass: pass out of range");
    if (passes) {
        if (passes > 1 && (rc = rsEnsureTemp2(f))) return rc;
        FrameDev d = rsToFrameDev(f, f->row0, f->row1);
        ResvD* buf[2] = {f->resvTemp, f->resvTemp2};
        if (pass == 1) stageBegin(f, RSTR_T_SPATIAL);
        launchRestirB(f->sc->dev, d, *prm, iter, buf[(pass - 1) & 1], buf[pass & 1], pass, pass == passes ? 1 : 0, f->stream);
        g_launches++;
        if (pass == passes) stageEnd(f, RSTR_T_SPATIAL);
        CU(cudaGetLastError());
    }
    if (pass >= passes) {
        f->resvOut ^= 1;                                                   // std::swap, restir.cu:434
        f->first = false;                                                  // restir.cu:436-438
    }
    return RSTR_OK;
}

int rstr_restir_phase_b(RstrFrame* f, const RstrCamera* cam, const RstrParams* prm, int looper, int iter) {
    if (!f || !prm) return fail(RSTR_ERR_ARG, "null frame/params");
    const int passes = (prm->reuse & 2) ? (prm->spatialPasses < 1 ? 1 : prm->spatialPasses) : 0;
    if (passes > 1 && f->bufRows != f->H)
        return fail(RSTR_ERR_ARG, "strip frames with spatialPasses > 1: call rstr_restir_phase_b_pass per pass and exchange the published halo rows in between");
    for (int pass = 1; pass <= (passes ? passes : 1); pass++) {
        int rc = rstr_restir_phase_b_pass(f, cam, prm, looper, iter, pass);
        if (rc) return rc;
    }
    return RSTR_OK;
}

// Cost profile for placing strip cuts: while enabled, the G-buffer and phase-A kernels add the SM cycles each block
// (16x8 pixels) held its SM slot to one accumulator per group of 8 image rows.
int rstr_frame_row_cost(RstrFrame* f, int enable, double* cyclesPerRowGroup, int numGroups) {
    if (!f) return fail(RSTR_ERR_ARG, "null frame");
    const int G = (f->H + 7) / 8;
    if (cyclesPerRowGroup) {
        if (numGroups != G || !f->rowCost) return fail(RSTR_ERR_ARG, "rstr_frame_row_cost: numGroups must be ceil(H / 8) and the profile enabled");
        std::vector<unsigned long long> h(G);
        CU(cudaMemcpyAsync(h.data(), f->rowCost, G * sizeof(unsigned long long), cudaMemcpyDeviceToHost, f->stream));
        CU(cudaStreamSynchronize(f->stream));
        for (int i = 0; i < G; i++) cyclesPerRowGroup[i] = (double)h[i];
    }
    if (enable) {
        if (!f->rowCost) CU(cudaMalloc((void**)&f->rowCost, G * sizeof(unsigned long long)));
        CU(cudaMemsetAsync(f->rowCost, 0, G * sizeof(unsigned long long), f->stream));
    } else if (f->rowCost) {
        CU(cudaStreamSynchronize(f->stream));
        cudaFree(f->rowCost); f->rowCost = nullptr;
    }
    return RSTR_OK;
}

int rstr_frame_set_halo_render(RstrFrame* f, int renderHalo) {
    if (!f) return fail(RSTR_ERR_ARG, "null frame");
    f->renderHalo = renderHalo != 0;
    return RSTR_OK;
}

int rstr_restir_direct(RstrFrame* f, const RstrCamera* cam, const RstrParams* prm, int looper, int iter) {
    int rc = rstr_restir_phase_a(f, cam, prm, looper, iter);
    if (rc) return rc;
    return rstr_restir_phase_b(f, cam, prm, looper, iter);
}

int rstr_pathtrace_direct(RstrFrame* f, const RstrCamera* cam, int looper, int iter) {
    int rc = checkCam(f, cam);
    if (rc) return rc;
    if ((rc = rsFlushGBuffer(f))) return rc;
    FrameDev d = rsToFrameDev(f, f->row0, f->row1);
    stageBegin(f, RSTR_T_PTDIRECT);
    g_launches += launchPTDirect(f->sc->dev, d, rsToCamDev(*cam), looper, iter, f->stream);
    stageEnd(f, RSTR_T_PTDIRECT);
    CU(cudaGetLastError());
    return RSTR_OK;
}

int rstr_tonemap(RstrFrame* f, int toneMapping, float scale) {
    if (!f) return fail(RSTR_ERR_ARG, "null frame");
    size_t off = (size_t)(f->row0 - f->bufRow0) * f->W, n = (size_t)(f->row1 - f->row0) * f->W;
    stageBegin(f, RSTR_T_TONEMAP);
    launchTonemap(f->radiance + 3 * off, f->ldr + off, n, toneMapping, scale, f->stream);
    stageEnd(f, RSTR_T_TONEMAP);
    g_launches++;
    CU(cudaGetLastError());
    return RSTR_OK;
}

// saveImage (main.cpp:105-144): tone-map + gamma, mirror horizontally (img.setPixel(width - 1 - x, y, ...)), 8-bit PNG / JPEG
static int frameSaveImage(RstrFrame* f, const char* path, int toneMapping, bool jpg);
int rstr_frame_save_png(RstrFrame* f, const char* path, int toneMapping) { return frameSaveImage(f, path, toneMapping, false); }
int rstr_frame_save_jpg(RstrFrame* f, const char* path, int toneMapping) { return frameSaveImage(f, path, toneMapping, true); }
static int frameSaveImage(RstrFrame* f, const char* path, int toneMapping, bool jpg) {
    if (!f || !path) return fail(RSTR_ERR_ARG, "rstr_frame_save_png/jpg: bad argument");
    int rc = rstr_tonemap(f, toneMapping, 1.f);
    if (rc) return rc;
    const int W = f->W, H = f->row1 - f->row0;
    std::vector<uchar4> ldr((size_t)W * H);
    CU(cudaMemcpyAsync(ldr.data(), f->ldr + (size_t)(f->row0 - f->bufRow0) * W, ldr.size() * sizeof(uchar4), cudaMemcpyDeviceToHost, f->stream));
    CU(cudaStreamSynchronize(f->stream));
    std::vector<unsigned char> rgb((size_t)W * H * 3);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const uchar4 c = ldr[(size_t)y * W + x];
            unsigned char* o = &rgb[((size_t)y * W + (W - 1 - x)) * 3];
            o[0] = c.x; o[1] = c.y; o[2] = c.z;
        }
    std::string err;
    if (!(jpg ? writeJPG(path, W, H, rgb.data(), 90, err) : writePNG(path, W, H, rgb.data(), err))) return fail(RSTR_ERR_IO, err);
    return RSTR_OK;
}

// Image::saveJPG (image.cpp:59-75); quality 0 = the library default (90, which is also what the reference passes)
int rstr_image_write_jpg(const char* path, int width, int height, const unsigned char* rgb, int quality) {
    if (!path || width <= 0 || height <= 0 || !rgb) return fail(RSTR_ERR_ARG, "rstr_image_write_jpg: bad argument");
    std::string err;
    if (!writeJPG(path, width, height, rgb, quality, err)) return fail(RSTR_ERR_IO, err);
    return RSTR_OK;
}

// Image::Image(filename) (image.cpp:16-33).  rgbOut may be NULL to query the size only.
int rstr_image_load(const char* path, int flipY, int* width, int* height, float* rgbOut, size_t capacityBytes) {
    if (!path || !width || !height) return fail(RSTR_ERR_ARG, "rstr_image_load: bad argument");
    HostTexture t;
    std::string err;
    try {
        if (!loadImageRGB(path, flipY != 0, t, err)) return fail(RSTR_ERR_IO, err);
    } catch (const std::exception& e) {
        return fail(RSTR_ERR_IO, std::string("rstr_image_load: ") + e.what());
    }
    *width = t.w; *height = t.h;
    if (rgbOut) {
        if (capacityBytes < t.rgb.size() * 12) return fail(RSTR_ERR_ARG, "rstr_image_load: buffer too small");
        memcpy(rgbOut, t.rgb.data(), t.rgb.size() * 12);
    }
    return RSTR_OK;
}

// Image::savePNG (image.cpp:41-57)
int rstr_image_write_png(const char* path, int width, int height, const unsigned char* rgb) {
    if (!path || width <= 0 || height <= 0 || !rgb) return fail(RSTR_ERR_ARG, "rstr_image_write_png: bad argument");
    std::string err;
    if (!writePNG(path, width, height, rgb, err)) return fail(RSTR_ERR_IO, err);
    return RSTR_OK;
}

int rstr_render_frame_host(RstrFrame* f, const RstrCamera* cam, const RstrParams* prm, int looper, int iter, int toneMapping, void* hostLdr, size_t bytes) {
    int rc = rstr_gbuffer_render(f, cam);
    if (rc) return rc;
    rc = prm ? rstr_restir_direct(f, cam, prm, looper, iter) : rstr_pathtrace_direct(f, cam, looper, iter);
    if (rc) return rc;
    rc = rstr_tonemap(f, toneMapping, 1.f);
    if (rc) return rc;
    if ((rc = rstr_gbuffer_update(f, cam))) return rc;
    size_t off = (size_t)(f->row0 - f->bufRow0) * f->W, n = (size_t)(f->row1 - f->row0) * f->W;
    if (hostLdr) {
        if (bytes != n * sizeof(uchar4)) return fail(RSTR_ERR_ARG, "rstr_render_frame_host: size mismatch");
        CU(cudaMemcpyAsync(hostLdr, f->ldr + off, bytes, cudaMemcpyDeviceToHost, f->stream));
    }
    CU(cudaStreamSynchronize(f->stream));
    return RSTR_OK;
}

// Pipelined form of rstr_render_frame_host: enqueues frame k (render + tone-map into LDR slot k&1 + D2H on a copy
// stream) and returns without waiting; rstr_frame_wait_host(slot) blocks until that slot's image is in host memory.
int rstr_render_frame_host_async(RstrFrame* f, const RstrCamera* cam, const RstrParams* prm, int looper, int iter, int toneMapping,
                                 void* hostLdr, size_t bytes, int slot) {
    if (!f || slot < 0 || slot >= RSTR_LDR_SLOTS || !hostLdr) return fail(RSTR_ERR_ARG, "rstr_render_frame_host_async: bad argument");
    const size_t off = (size_t)(f->row0 - f->bufRow0) * f->W, n = (size_t)(f->row1 - f->row0) * f->W;
    if (bytes != n * sizeof(uchar4)) return fail(RSTR_ERR_ARG, "rstr_render_frame_host_async: size mismatch");
    if (!f->copyStream) {
        CU(cudaStreamCreateWithFlags(&f->copyStream, cudaStreamNonBlocking));
        for (int i = 0; i < RSTR_LDR_SLOTS; i++) {
            CU(cudaMalloc((void**)&f->ldrB[i], n * sizeof(uchar4)));
            CU(cudaEventCreateWithFlags(&f->evRendered[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&f->evCopied[i], cudaEventDisableTiming));
        }
    }
    int rc = rstr_gbuffer_render(f, cam);
    if (rc) return rc;
    rc = prm ? rstr_restir_direct(f, cam, prm, looper, iter) : rstr_pathtrace_direct(f, cam, looper, iter);
    if (rc) return rc;
    if (f->slotBusy[slot]) CU(cudaStreamWaitEvent(f->stream, f->evCopied[slot], 0));   // the slot's previous image must have left the device
    stageBegin(f, RSTR_T_TONEMAP);
    launchTonemap(f->radiance + 3 * off, f->ldrB[slot], n, toneMapping, 1.f, f->stream);
    stageEnd(f, RSTR_T_TONEMAP);
    g_launches++;
    CU(cudaGetLastError());
    if ((rc = rstr_gbuffer_update(f, cam))) return rc;
    CU(cudaEventRecord(f->evRendered[slot], f->stream));
    CU(cudaStreamWaitEvent(f->copyStream, f->evRendered[slot], 0));
    CU(cudaMemcpyAsync(hostLdr, f->ldrB[slot], bytes, cudaMemcpyDeviceToHost, f->copyStream));
    CU(cudaEventRecord(f->evCopied[slot], f->copyStream));
    f->slotBusy[slot] = true;
    return RSTR_OK;
}

int rstr_frame_wait_host(RstrFrame* f, int slot) {
    if (!f || slot < 0 || slot >= RSTR_LDR_SLOTS) return fail(RSTR_ERR_ARG, "rstr_frame_wait_host: bad argument");
    if (f->slotBusy[slot]) CU(cudaEventSynchronize(f->evCopied[slot]));
    return RSTR_OK;
}

int rstr_frame_sync(RstrFrame* f) {
    if (!f) return fail(RSTR_ERR_ARG, "null frame");
    int rc = rsFlushGBuffer(f);
    if (rc) return rc;
    CU(cudaStreamSynchronize(f->stream));
    return RSTR_OK;
}

static int ensureScratch(RstrFrame* f, size_t bytes) {
    if (f->scratchBytes >= bytes) return RSTR_OK;
    cudaFree(f->scratch); f->scratch = nullptr; f->scratchBytes = 0;
    CU(cudaMalloc(&f->scratch, bytes));
    f->scratchBytes = bytes;
    return RSTR_OK;
}

static int frameReadImpl(RstrFrame* f, int which, void* host, size_t bytes, bool toDevice);
int rstr_frame_read(RstrFrame* f, int which, void* host, size_t bytes) { return frameReadImpl(f, which, host, bytes, false); }
// same conversion, but into DEVICE memory and without synchronising (stream-ordered): strip gather to GPU 0
int rstr_frame_read_device(RstrFrame* f, int which, void* dev, size_t bytes) { return frameReadImpl(f, which, dev, bytes, true); }

static int frameReadImpl(RstrFrame* f, int which, void* host, size_t bytes, bool toDevice) {
    if (!f || !host) return fail(RSTR_ERR_ARG, "rstr_frame_read: bad argument");
    {
        int rc = rsFlushGBuffer(f);
        if (rc) return rc;
    }
    const size_t off = (size_t)(f->row0 - f->bufRow0) * f->W, n = (size_t)(f->row1 - f->row0) * f->W;
    size_t elem = 0;
    switch (which) {
    case RSTR_BUF_ALBEDO: case RSTR_BUF_NORMAL: case RSTR_BUF_RADIANCE: elem = 12; break;
    case RSTR_BUF_MATID: case RSTR_BUF_DEPTH: case RSTR_BUF_MOTION: case RSTR_BUF_LIGHT_INDEX: case RSTR_BUF_LDR: elem = 4; break;
    case RSTR_BUF_RESERVOIR: case RSTR_BUF_RESERVOIR_TEMP: elem = 36; break;
    default: return fail(RSTR_ERR_ARG, "rstr_frame_read: unknown buffer");
    }
    if (bytes != n * elem) return fail(RSTR_ERR_ARG, "rstr_frame_read: size mismatch");
    const void* src = nullptr;
    cudaStream_t st = f->stream;
    if (which == RSTR_BUF_RADIANCE) src = f->radiance + 3 * off;
    else if (which == RSTR_BUF_MATID) src = f->matId[f->cur] + off;
    else if (which == RSTR_BUF_LDR) src = f->ldr + off;
    else {
        int rc = ensureScratch(f, bytes);
        if (rc) return rc;
        const float4* g = f->geom[f->cur] + off;
        const float4* am = f->albedoMotion + off;
        // after the swap the history written by the last frame is resv[resvOut ^ 1]
        const ResvD* hist = f->resv[f->resvOut ^ 1] + off;
        switch (which) {
        case RSTR_BUF_ALBEDO: launchExportGeom(g, am, (float*)f->scratch, nullptr, nullptr, nullptr, n, st); break;
        case RSTR_BUF_NORMAL: launchExportGeom(g, am, nullptr, (float*)f->scratch, nullptr, nullptr, n, st); break;
        case RSTR_BUF_DEPTH: launchExportGeom(g, am, nullptr, nullptr, (float*)f->scratch, nullptr, n, st); break;
        case RSTR_BUF_MOTION: launchExportGeom(g, am, nullptr, nullptr, nullptr, (int*)f->scratch, n, st); break;
        case RSTR_BUF_RESERVOIR: launchExportResv(f->sc->dev, hist, (float*)f->scratch, nullptr, n, st); break;
        case RSTR_BUF_RESERVOIR_TEMP: launchExportResv(f->sc->dev, f->resvTemp + off, (float*)f->scratch, nullptr, n, st); break;
        case RSTR_BUF_LIGHT_INDEX: launchExportResv(f->sc->dev, hist, nullptr, (int*)f->scratch, n, st); break;
        }
        g_launches++;
        CU(cudaGetLastError());
        src = f->scratch;
    }
    CU(cudaMemcpyAsync(host, src, bytes, toDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    if (!toDevice) CU(cudaStreamSynchronize(st));
    return RSTR_OK;
}

int rstr_frame_stage_ms(RstrFrame* f, float* ms, int n) {
    if (!f || !ms) return fail(RSTR_ERR_ARG, "rstr_frame_stage_ms: bad argument");
    CU(cudaStreamSynchronize(f->stream));
    for (int s = 0; s < n && s < RSTR_T_COUNT; s++) {
        ms[s] = 0.f;
        if (f->ran[s]) CU(cudaEventElapsedTime(&ms[s], f->ev[2 * s], f->ev[2 * s + 1]));
    }
    return RSTR_OK;
}

// user timing marks on the frame's stream (bench.py: CUDA events on the stream the kernels are launched on)
int rstr_frame_mark(RstrFrame* f, int slot) {
    if (!f || slot < 0 || slot >= 8) return fail(RSTR_ERR_ARG, "rstr_frame_mark: bad argument");
    if (!f->marks[slot]) CU(cudaEventCreate(&f->marks[slot]));
    CU(cudaEventRecord(f->marks[slot], f->stream));
    return RSTR_OK;
}
int rstr_frame_elapsed_ms(RstrFrame* f, int slotA, int slotB, float* ms) {
    if (!f || !ms || slotA < 0 || slotA >= 8 || slotB < 0 || slotB >= 8 || !f->marks[slotA] || !f->marks[slotB]) return fail(RSTR_ERR_ARG, "rstr_frame_elapsed_ms: bad argument");
    CU(cudaEventSynchronize(f->marks[slotB]));
    CU(cudaEventElapsedTime(ms, f->marks[slotA], f->marks[slotB]));
    return RSTR_OK;
}
// run the frame's work on a caller-owned stream (e.g. torch's current stream, so NCCL calls order naturally)
int rstr_frame_set_stream(RstrFrame* f, void* stream) {
    if (!f) return fail(RSTR_ERR_ARG, "null frame");
    CU(cudaStreamSynchronize(f->stream));
    if (f->ownStream) { cudaStreamDestroy(f->stream); f->ownStream = false; }
    f->stream = (cudaStream_t)stream;
    return RSTR_OK;
}

uint64_t rstr_launch_count(void) { return g_launches.load(); }
void* rstr_frame_stream(RstrFrame* f) { return f ? (void*)f->stream : nullptr; }

int rstr_frame_halo_miss(RstrFrame* f, unsigned int* out) {
    if (!f || !out) return fail(RSTR_ERR_ARG, "rstr_frame_halo_miss: bad argument");
    CU(cudaMemcpyAsync(out, f->haloMiss, sizeof(unsigned int), cudaMemcpyDeviceToHost, f->stream));
    CU(cudaStreamSynchronize(f->stream));
    return RSTR_OK;
}

int rstr_frame_halo_miss_reset(RstrFrame* f) {
    if (!f) return fail(RSTR_ERR_ARG, "rstr_frame_halo_miss_reset: null frame");
    CU(cudaMemsetAsync(f->haloMiss, 0, sizeof(unsigned int), f->stream));
    return RSTR_OK;
}

int rstr_frame_motion_rows(RstrFrame* f, unsigned int* out, int reset) {
    if (!f || !out) return fail(RSTR_ERR_ARG, "rstr_frame_motion_rows: bad argument");
    int rc = rsFlushGBuffer(f);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out, f->haloMiss + 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, f->stream));
    if (reset) CU(cudaMemsetAsync(f->haloMiss + 1, 0, sizeof(unsigned int), f->stream));
    CU(cudaStreamSynchronize(f->stream));
    return RSTR_OK;
}

int rstr_frame_plane_row(RstrFrame* f, int plane, int row, void** devPtr, size_t* rowBytes) {
    if (!f || !devPtr || !rowBytes) return fail(RSTR_ERR_ARG, "rstr_frame_plane_row: bad argument");
    if (row < f->bufRow0 || row >= f->bufRow0 + f->bufRows) return fail(RSTR_ERR_ARG, "rstr_frame_plane_row: row not resident");
    {
        int rc = rsFlushGBuffer(f);
        if (rc) return rc;
    }
    size_t off = (size_t)(row - f->bufRow0) * f->W;
    switch (plane) {
    case RSTR_PLANE_GEOM_CUR: *devPtr = f->geom[f->cur] + off; *rowBytes = (size_t)f->W * sizeof(float4); break;
    case RSTR_PLANE_MATID_CUR: *devPtr = f->matId[f->cur] + off; *rowBytes = (size_t)f->W * sizeof(int); break;
    case RSTR_PLANE_RESV_HISTORY: *devPtr = f->resv[f->resvOut ^ 1] + off; *rowBytes = (size_t)f->W * sizeof(ResvD); break;
    case RSTR_PLANE_RESV_TEMP: *devPtr = f->resvTemp + off; *rowBytes = (size_t)f->W * sizeof(ResvD); break;
    case RSTR_PLANE_RESV_OUT: *devPtr = f->resv[f->resvOut] + off; *rowBytes = (size_t)f->W * sizeof(ResvD); break;
    case RSTR_PLANE_RESV_TEMP2: {
        int rc = rsEnsureTemp2(f);
        if (rc) return rc;
        *devPtr = f->resvTemp2 + off; *rowBytes = (size_t)f->W * sizeof(ResvD); break;
    }
    default: return fail(RSTR_ERR_ARG, "rstr_frame_plane_row: unknown plane");
    }
    return RSTR_OK;
}

int rstr_frame_copy_rows(RstrFrame* dst, RstrFrame* src, int plane, int row0, int row1) {
    if (!dst || !src || dst->W != src->W || row0 >= row1) return fail(RSTR_ERR_ARG, "rstr_frame_copy_rows: bad argument");
    void *ps = nullptr, *pd = nullptr;
    size_t rb = 0;
    int rc = rstr_frame_plane_row(src, plane, row0, &ps, &rb);
    if (rc) return rc;
    rc = rstr_frame_plane_row(dst, plane, row0, &pd, &rb);
    if (rc) return rc;
    if (row1 > src->bufRow0 + src->bufRows || row1 > dst->bufRow0 + dst->bufRows) return fail(RSTR_ERR_ARG, "rstr_frame_copy_rows: rows not resident");
    // order: after everything queued on src's stream, before anything queued later on dst's stream
    CU(cudaEventRecord(src->xfer, src->stream));
    CU(cudaStreamWaitEvent(dst->stream, src->xfer, 0));
    CU(cudaMemcpyAsync(pd, ps, rb * (size_t)(row1 - row0), cudaMemcpyDefault, dst->stream));
    // ... and src must not overwrite the rows (next frame's phase A) while dst's stream is still reading them
    CU(cudaEventRecord(dst->xfer, dst->stream));
    CU(cudaStreamWaitEvent(src->stream, dst->xfer, 0));
    return RSTR_OK;
}

void* rstr_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { g_err = "cudaHostAlloc failed"; return nullptr; }
    return p;
}
void rstr_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"
