// kernels_emu.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_host_shim.h).  The device code of restir_b200/csrc/kernels.cu + gi_kernels.inl +
// denoise.cu compiled by g++, with the same libm as the oracle, so the two must agree bit for bit.  Two levels:
//   per pixel    emu_gbuffer_render / emu_restir_indirect (modes 0-3): the bodies of the G-buffer and GI kernels, one "thread" per call
//   per kernel   emu_restir_indirect (modes 4-6), emu_di_*, emu_denoiser_*: the __global__ functions launched with emuLaunch as grids of
//                32-lane warps (fibers + exchanged warp intrinsics), in the order the library's launchers use
// Built by tests/emu/build.py into tests/emu/_build (git-ignored); never part of librestir_b200.so.
#include "cuda_host_shim.h"

#include "../../restir_b200/csrc/kernels.cu"
#include "../../restir_b200/csrc/denoise.cu"
#include "../../restir_b200/csrc/bvh_gpu.cu"

#include <algorithm>
#include <numeric>

#include <string>
#include <vector>

using namespace rs;

namespace {

struct EmuScene {
    HostScene hs;
    DevScene dev{};
    unsigned int fallback[4] = {0, 0, 0, 0};
};

struct EmuFrame {
    EmuScene* sc;
    int W, H;
    std::vector<float4> geom[2], albedoMotion;
    std::vector<int> matId[2];
    std::vector<float4> resv[2];
    std::vector<float> nsz[2], indirect, export17;
    unsigned int counters[4] = {0, 0, 0, 0};
    std::vector<int> queue;                      // fix-up queue of the kernels run under emuLaunch
    unsigned int queueCount[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int cur = 0, out = 0;
    bool first = true, haveLast = false;
    RstrCamera lastCamera{};
    unsigned long long undecided = 0;
};

CamDev toCamDev(const RstrCamera& c) {          // rsToCamDev (capi.cu)
    CamDev d;
    memcpy(d.position, c.position, 12); memcpy(d.right, c.right, 12); memcpy(d.up, c.up, 12); memcpy(d.view, c.view, 12);
    memcpy(d.rotInv, c.rotationMatInv, 36);
    d.aspect = (float)c.resolution[0] / c.resolution[1];
    d.tanFovY = tanf(radians(c.fov[1]));
    d.focalDist = c.focalDist;
    d.pixelSizeX = 1.f / (float)c.resolution[0];
    d.pixelSizeY = 1.f / (float)c.resolution[1];
    d.resX = (float)c.resolution[0]; d.resY = (float)c.resolution[1];
    return d;
}

FrameDev toFrameDev(EmuFrame* f) {
    FrameDev d{};
    d.W = f->W; d.H = f->H; d.rowLo = 0; d.rowHi = f->H; d.bufRow0 = 0; d.bufRows = f->H;
    d.geom[0] = f->geom[f->cur].data(); d.geom[1] = f->geom[f->cur ^ 1].data();
    d.matId[0] = f->matId[f->cur].data(); d.matId[1] = f->matId[f->cur ^ 1].data();
    d.albedoMotion = f->albedoMotion.data();
    d.haloMiss = f->counters; d.motionRows = f->counters + 1;
    f->queue.resize((size_t)f->W * f->H);
    d.queue = f->queue.data(); d.queueCount = f->queueCount;
    return d;
}

}  // namespace

extern "C" {

void* emu_scene_create(const RstrSceneDesc* desc) {                 // the HostScene fill of rstr_scene_create (capi.cu) + buildHostScene
    EmuScene* sc = new EmuScene;
    HostScene& hs = sc->hs;
    const int T = desc->numTris;
    hs.T = T;
    hs.vertices.resize(3 * (size_t)T); hs.normals.resize(3 * (size_t)T); hs.texcoords.assign(6 * (size_t)T, 0.f);
    memcpy(hs.vertices.data(), desc->vertices, 36 * (size_t)T);
    memcpy(hs.normals.data(), desc->normals, 36 * (size_t)T);
    if (desc->texcoords) memcpy(hs.texcoords.data(), desc->texcoords, 24 * (size_t)T);
    hs.materialIds.assign(desc->materialIds, desc->materialIds + T);
    hs.materials.assign(desc->materials, desc->materials + desc->numMaterials);
    hs.textures.resize(desc->numTextures);
    for (int t = 0; t < desc->numTextures; t++) {
        const RstrTexture& src = desc->textures[t];
        hs.textures[t].w = src.width; hs.textures[t].h = src.height;
        hs.textures[t].rgb.resize((size_t)src.width * src.height);
        memcpy(hs.textures[t].rgb.data(), src.rgb, 12 * hs.textures[t].rgb.size());
    }
    hs.envMapTexId = desc->envMap - 1;
    std::string err;
    if (!buildHostScene(hs, err)) { fprintf(stderr, "emu_scene_create: %s\n", err.c_str()); delete sc; return nullptr; }
    DevScene& d = sc->dev;                                          // ensureUploaded (capi.cu) with the host arrays themselves
    d.fallbackRays = sc->fallback;
    d.nodes = (const float4*)hs.packed.data(); d.triGeom = (const float4*)hs.fastTris.data(); d.triNorm = (const float4*)hs.triNorm.data();
    d.materials = hs.materials.data(); d.alias = (const float2*)hs.alias.data(); d.lights = (const float4*)hs.lights.data();
    d.fastNodes = (const float4*)hs.fastNodes.data(); d.primToFast = hs.primToFast.data(); d.rank = hs.rank.data();
    d.numTris = hs.T; d.numFastNodes = (int)hs.fastNodes.size(); d.fastRoot = hs.fastRoot; d.traversal = RS_TRAVERSAL_EXACT;
    memcpy(d.fastRootMin, hs.fastRootMin, 12); memcpy(d.fastRootMax, hs.fastRootMax, 12);
    d.numLights = (int)hs.alias.size();
    d.texData = (const float4*)hs.texData.data(); d.texInfo = (const int4*)hs.texInfo.data(); d.triUV = (const float4*)hs.triUV.data();
    d.anyMaps = hs.anyMaps ? 1 : 0;
    d.envTex = hs.envMapTexId; d.envLen = (int)hs.envAlias.size();
    d.envAlias = (const float2*)hs.envAlias.data(); d.envDir = (const float4*)hs.envDir.data();
    d.sumLightPowerInv = hs.sumLightPowerInv;
    d.rootRef = hs.rootRef;
    memcpy(d.rootMin, &hs.rootBox.pMin, 12); memcpy(d.rootMax, &hs.rootBox.pMax, 12);
    return sc;
}
void emu_scene_destroy(void* sc) { delete (EmuScene*)sc; }

void* emu_frame_create(void* sc, int W, int H) {
    EmuFrame* f = new EmuFrame;
    f->sc = (EmuScene*)sc; f->W = W; f->H = H;
    const size_t n = (size_t)W * H;
    for (int i = 0; i < 2; i++) {
        f->geom[i].assign(n, make_float4(0, 0, 0, 0)); f->matId[i].assign(n, 0);
        f->resv[i].assign(4 * n, make_float4(0, 0, 0, 0)); f->nsz[i].assign(n, 0.f);
    }
    f->albedoMotion.assign(n, make_float4(0, 0, 0, 0));
    f->indirect.assign(3 * n, 0.f);
    return f;
}
void emu_frame_destroy(void* f) { delete (EmuFrame*)f; }

// rstr_gbuffer_render in RS_TRAVERSAL_EXACT: k_gbuffer_exact's body per pixel
void emu_gbuffer_render(void* fv, const RstrCamera* cam) {
    EmuFrame* f = (EmuFrame*)fv;
    const FrameDev d = toFrameDev(f);
    const CamDev c = toCamDev(*cam), lc = toCamDev(f->haveLast ? f->lastCamera : *cam);
    const DevScene& s = f->sc->dev;
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < f->H; y++)
        for (int x = 0; x < f->W; x++) {
            threadIdx.x = 0;
            RS_DECLARE_STACK(stack);
            gbufferPixelExact(s, d, c, lc, x, y, stack);
        }
}
void emu_gbuffer_update(void* fv, const RstrCamera* cam) {
    EmuFrame* f = (EmuFrame*)fv;
    f->lastCamera = *cam; f->haveLast = true; f->cur ^= 1;
}
void emu_gi_reset(void* fv) { ((EmuFrame*)fv)->first = true; }

// rstr_restir_indirect.  tracedTree = 2 / 3: the staged / ray-queue pipeline's bodies; 4 / 5 / 6: the ray-queue / staged / one-kernel form's KERNELS as warps (below).  tracedTree = 0: k_restir_indirect_exact's body per pixel.  tracedTree = 1: the primary hit from the reference-order
// walk (standing in for the packet walk), then giAfterHit<false>: bounce rays through traceClosestFast, shadow rays through
// traceOccludedFast; an undecided pixel is recomputed like k_restir_indirect_fix does.
void emu_restir_indirect(void* fv, const RstrCamera* cam, int looper, int iter, int traceDepth, int reuse, int tracedTree) {
    EmuFrame* f = (EmuFrame*)fv;
    const FrameDev d = toFrameDev(f);
    const CamDev c = toCamDev(*cam);
    const DevScene& s = f->sc->dev;
    GIDev g{};
    g.resvOut = f->resv[f->out].data(); g.resvIn = f->resv[f->out ^ 1].data();
    g.nszOut = f->nsz[f->out].data(); g.nszIn = f->nsz[f->out ^ 1].data();
    g.indirect = f->indirect.data(); g.fallback = nullptr;
    g.maxDepth = traceDepth; g.reuse = reuse; g.first = f->first ? 1 : 0; g.iter = iter; g.bounceWalk = RS_TRAVERSAL_FAST;
    unsigned long long undecided = 0;
    if (tracedTree == 6) {                                           // the one-kernel form as warps: k_restir_indirect (packet walk included) + fix-up
        f->queueCount[0] = 0;
        emuLaunch((unsigned)((f->W + 15) / 16), (unsigned)((f->H + 7) / 8), [&] { k_restir_indirect(s, d, c, g, looper); });
        undecided = f->queueCount[0];
        emuLaunch(2, 1, [&] { k_restir_indirect_fix(s, d, c, g, looper); });
        f->undecided += undecided;
        f->out ^= 1;
        f->first = false;
        return;
    }
    if (tracedTree >= 2) {
        // the staged form: giStagePrimary per pixel (primary hit from the reference-order walk, as above), giStageBounce per live path and depth
        // through the queues, giStageResolve per pixel, marked pixels like k_restir_indirect_fix.  One-lane "warps": giAppendPath appends one record.
        const size_t n = (size_t)f->W * f->H;
        std::vector<float4> pix(7 * n), q0(RS_GI_PATH_F4 * n), q1(RS_GI_PATH_F4 * n);
        std::vector<int> status(n, 0);
        std::vector<unsigned int> counts(traceDepth + 2, 0u);
        g.pix = pix.data(); g.pixStatus = status.data(); g.pathQ[0] = q0.data(); g.pathQ[1] = q1.data(); g.pathCount = counts.data(); g.pixStride = n;
        f->queueCount[0] = 0;
        if (tracedTree >= 4)                                         // kernel-level modes: the real k_gi_primary (packet walk of the jittered rays)
            emuLaunch((unsigned)((f->W + 15) / 16), (unsigned)((f->H + 7) / 8), [&] { k_gi_primary(s, d, c, g, looper); });
#pragma omp parallel for schedule(dynamic, 4)
        for (int y = 0; y < (tracedTree >= 4 ? 0 : f->H); y++)
            for (int x = 0; x < f->W; x++) {
                threadIdx.x = 0;
                RS_DECLARE_STACK(stack);
                Rng rng;
                f3 o, dir;
                jitteredRay(d, c, looper, x, y, rng, o, dir);
                const RayT ray = makeRayT(o, dir);
                Hit h;
                traceClosestExact(s, ray, h, stack);
                GIPathRec rec;
                const bool live = giStagePrimary(s, d, g, x, y, rng, dir, h, rec);
                giAppendPath(g.pathQ[0], g.pathCount + 1, live, rec);
            }
        // tracedTree == 4: the ray-queue form's KERNELS (k_gi_head, k_gi_walk_shadow, k_gi_walk_closest, k_gi_tail, k_gi_resolve,
        // k_restir_indirect_fix) as grids of 32-lane warps (emuLaunch): lane refill, leaf wait, warp-aggregated appends and all
        if (tracedTree == 5) {                                       // the staged form's kernels as warps: k_gi_bounce per depth, k_gi_resolve, fix-up
            const unsigned blocks = (unsigned)((n + RS_BLOCK - 1) / RS_BLOCK);
            for (int depth = 1; depth <= traceDepth; depth++) emuLaunch(blocks, 1, [&] { k_gi_bounce(s, d, g, depth); });
            emuLaunch((unsigned)((f->W + 15) / 16), (unsigned)((f->H + 7) / 8), [&] { k_gi_resolve(d, g); });
            undecided = f->queueCount[0];
            emuLaunch(2, 1, [&] { k_restir_indirect_fix(s, d, c, g, looper); });
            f->undecided += undecided;
            f->out ^= 1;
            f->first = false;
            return;
        }
        if (tracedTree == 4) {
            std::vector<unsigned int> cl(n), sl(n), wc(4 * (traceDepth + 2), 0u);
            std::vector<float4> hit(n);
            std::vector<int> occ(n);
            g.closestList = cl.data(); g.shadowList = sl.data(); g.walkCount = wc.data(); g.hit = hit.data(); g.occ = occ.data();
            const unsigned blocks = (unsigned)((n + RS_BLOCK - 1) / RS_BLOCK);
            for (int depth = 1; depth <= traceDepth; depth++) {
                emuLaunch(blocks, 1, [&] { k_gi_head(s, d, g, depth); });
                if (depth > 1) emuLaunch(3, 1, [&] { k_gi_walk_shadow(s, g, depth); });
                emuLaunch(3, 1, [&] { k_gi_walk_closest(s, g, depth); });
                emuLaunch(blocks, 1, [&] { k_gi_tail(s, d, g, depth); });
            }
            emuLaunch((unsigned)((f->W + 15) / 16), (unsigned)((f->H + 7) / 8), [&] { k_gi_resolve(d, g); });
            undecided = f->queueCount[0];
            emuLaunch(2, 1, [&] { k_restir_indirect_fix(s, d, c, g, looper); });
            f->undecided += undecided;
            f->out ^= 1;
            f->first = false;
            return;
        }
        // tracedTree == 3: the ray-queue form -- k_gi_head / the two walkers (one ray per call here) / k_gi_tail per depth
        std::vector<unsigned int> closestList(tracedTree == 3 ? n : 0), shadowList(tracedTree == 3 ? n : 0), walkCount(4 * (traceDepth + 2), 0u);
        std::vector<float4> hit(tracedTree == 3 ? n : 0);
        std::vector<int> occ(tracedTree == 3 ? n : 0);
        for (int depth = 1; tracedTree == 3 && depth <= traceDepth; depth++) {
            const long long np = counts[depth];
            float4* q = g.pathQ[(depth - 1) & 1];
#pragma omp parallel for schedule(dynamic, 16)
            for (long long i = 0; i < np; i++) {
                threadIdx.x = 0;
                GIPathRec rec = giLoadPath(q, (size_t)i);
                const int flags = giQHead(s, d, g, depth, rec);
                float4* o = q + RS_GI_PATH_F4 * (size_t)i;
                o[1] = rec.b; o[2] = rec.c; o[3] = rec.d; o[4] = rec.e;
                giAppendSlot(closestList.data(), walkCount.data() + 4 * depth, (flags & GI_Q_ALIVE) != 0, (unsigned)i);
                giAppendSlot(shadowList.data(), walkCount.data() + 4 * depth + 2, (flags & GI_Q_SHADOW) != 0, (unsigned)i);
            }
            const long long nc = walkCount[4 * depth], ns = walkCount[4 * depth + 2];
#pragma omp parallel for schedule(dynamic, 16)
            for (long long k = 0; k < ns; k++) {
                threadIdx.x = 0;
                RS_DECLARE_STACK(stack);
                occ[shadowList[k]] = giQShadowOf(s, giLoadPath(q, shadowList[k]), stack);
            }
#pragma omp parallel for schedule(dynamic, 16)
            for (long long k = 0; k < nc; k++) {
                threadIdx.x = 0;
                RS_DECLARE_STACK(stack);
                RS_DECLARE_PACKET(pk, 1);
                (void)pk_tb; (void)pk_wst;
                hit[closestList[k]] = giQClosestOf(s, g, giLoadPath(q, closestList[k]), stack, pk_ta);
            }
#pragma omp parallel for schedule(dynamic, 16)
            for (long long i = 0; i < np; i++) {
                threadIdx.x = 0;
                const GIPathRec in = giLoadPath(q, (size_t)i);
                const int flags = __float_as_int(in.d.w);
                GIPathRec rec;
                int x, y;
                const int r = giQTail(s, d, g, depth, in, (flags & GI_Q_SHADOW) ? occ[i] : 0, (flags & GI_Q_ALIVE) ? hit[i] : make_float4(0.f, 0.f, 0.f, 0.f), rec, x, y);
                giAppendPath(g.pathQ[depth & 1], g.pathCount + depth + 1, r == 1, rec);
            }
        }
        for (int depth = 1; tracedTree == 2 && depth <= traceDepth; depth++) {
            const long long np = counts[depth];
#pragma omp parallel for schedule(dynamic, 16)
            for (long long i = 0; i < np; i++) {
                threadIdx.x = 0;
                RS_DECLARE_STACK(stack);
                RS_DECLARE_PACKET(pk, 1);
                (void)pk_tb; (void)pk_wst;
                const GIPathRec in = giLoadPath(g.pathQ[(depth - 1) & 1], (size_t)i);
                GIPathRec rec;
                int x, y;
                const int r = giStageBounce(s, d, g, depth, in, stack, pk_ta, rec, x, y);
                giAppendPath(g.pathQ[depth & 1], g.pathCount + depth + 1, r == 1, rec);
            }
        }
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : undecided)
        for (int y = 0; y < f->H; y++)
            for (int x = 0; x < f->W; x++) {
                threadIdx.x = 0;
                giStageResolve(d, g, x, y);
                if (status[(size_t)y * f->W + x] == 2) {
                    RS_DECLARE_STACK(stack);
                    undecided++;
                    giPixelExact(s, d, c, g, looper, x, y, stack);
                }
            }
        f->undecided += undecided;
        f->out ^= 1;
        f->first = false;
        return;
    }
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : undecided)
    for (int y = 0; y < f->H; y++)
        for (int x = 0; x < f->W; x++) {
            threadIdx.x = 0;
            RS_DECLARE_STACK(stack);
            if (!tracedTree) { giPixelExact(s, d, c, g, looper, x, y, stack); continue; }
            RS_DECLARE_PACKET(pk, 1);
            (void)pk_tb; (void)pk_wst;
            Rng rng;
            f3 o, dir;
            jitteredRay(d, c, looper, x, y, rng, o, dir);
            const RayT ray = makeRayT(o, dir);
            Hit h;
            traceClosestExact(s, ray, h, stack);
            if (!giAfterHit<false>(s, d, g, x, y, stack, pk_ta, rng, dir, h)) {
                undecided++;
                giPixelExact(s, d, c, g, looper, x, y, stack);
            }
        }
    f->undecided += undecided;
    f->out ^= 1;
    f->first = false;
}
unsigned long long emu_gi_undecided(void* fv) { return ((EmuFrame*)fv)->undecided; }
const float* emu_gi_indirect(void* fv) { return ((EmuFrame*)fv)->indirect.data(); }
// k_export_gi's body: the reservoirs the last call wrote, 17 x 4 bytes each (numSamples as int bits)
const float* emu_gi_reservoirs(void* fv) {
    EmuFrame* f = (EmuFrame*)fv;
    const size_t n = (size_t)f->W * f->H;
    f->export17.resize(17 * n);
    blockDim.x = 1; threadIdx.x = 0;
    for (size_t i = 0; i < n; i++) {
        blockIdx.x = (unsigned)i;
        k_export_gi(f->resv[f->out ^ 1].data(), f->nsz[f->out ^ 1].data(), f->export17.data(), n);
    }
    return f->export17.data();
}
const void* emu_frame_plane(void* fv, int which) {                   // 0 geom {n, depth}, 1 matId, 2 {albedo, motion} of the current frame
    EmuFrame* f = (EmuFrame*)fv;
    return which == 0 ? (const void*)f->geom[f->cur].data() : which == 1 ? (const void*)f->matId[f->cur].data() : (const void*)f->albedoMotion.data();
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ the direct path's kernels as warps
// One frame of the staged pipeline the way launchPhaseAStaged / launchRestirB (kernels.cu) and rstr_restir_phase_a / _b (capi.cu) run it:
// k_primary (packet walk of the centre + jittered rays of every 8x4 tile: G-buffer and the shaded-pixel queue), k_candidates, k_shadow
// (persistent, lane refill, cooperative drain), k_temporal, the fix-up kernel, k_restir_b per spatial pass, and the export kernels --
// each launched with emuLaunch as a grid of 32-lane warps.
struct EmuDI {
    EmuScene* sc;
    DevScene dev;
    int W, H;
    int row0, row1, bufRow0, bufRows;            // a horizontal strip [row0, row1) with its resident rows (the full frame: 0, H, 0, H)
    std::vector<float4> geom[2], albedoMotion;
    std::vector<int> matId[2], queue, shadeQueue;
    std::vector<float> radiance, export36;
    std::vector<ResvD> resv[2], resvTemp, resvTemp2;
    std::vector<HitRec> hit;
    std::vector<float2> hitMR;
    std::vector<float4> hitPos;                  // unbiased mode
    std::vector<int> exportI;
    unsigned int queueCount[4 + 2 * RS_MAX_BANDS];
    unsigned int counters[4] = {0, 0, 0, 0};
    int cur = 0, resvOut = 0;
    int bands = 1;                               // rstr_frame_set_bands: the staged pipeline cuts the rows into bands with a queue each
    bool first = true, haveLast = false, temp2Ready = false;
    RstrCamera lastCamera{};
};
static FrameDev diFrameDev(EmuDI* f) {                              // rsToFrameDev (capi.cu)
    FrameDev d{};
    d.W = f->W; d.H = f->H; d.rowLo = f->row0; d.rowHi = f->row1; d.bufRow0 = f->bufRow0; d.bufRows = f->bufRows;
    d.geom[0] = f->geom[f->cur].data(); d.geom[1] = f->geom[f->cur ^ 1].data();
    d.matId[0] = f->matId[f->cur].data(); d.matId[1] = f->matId[f->cur ^ 1].data();
    d.albedoMotion = f->albedoMotion.data(); d.radiance = f->radiance.data();
    d.resvOut = f->resv[f->resvOut].data(); d.resvIn = f->resv[f->resvOut ^ 1].data(); d.resvTemp = f->resvTemp.data();
    d.hit = f->hit.data(); d.hitMR = f->hitMR.empty() ? nullptr : f->hitMR.data(); d.hitPos = f->hitPos.empty() ? nullptr : f->hitPos.data();
    d.haloMiss = f->counters; d.motionRows = f->counters + 1;
    d.queue = f->queue.data(); d.queueCount = f->queueCount; d.shadeQueue = f->shadeQueue.data();
    return d;
}

extern "C" {

// rstr_frame_create: rows [row0, row1) of the W x H image plus `halo` resident rows on each side (row1 <= 0: the full frame)
void* emu_di_create_strip(void* scv, int W, int H, int row0, int row1, int halo) {
    EmuDI* f = new EmuDI;
    f->sc = (EmuScene*)scv; f->dev = f->sc->dev; f->dev.traversal = RS_TRAVERSAL_FAST;
    f->W = W; f->H = H;
    if (row1 <= 0) { row0 = 0; row1 = H; halo = 0; }
    f->row0 = row0; f->row1 = row1;
    f->bufRow0 = row0 - halo < 0 ? 0 : row0 - halo;
    f->bufRows = (row1 + halo > H ? H : row1 + halo) - f->bufRow0;
    const size_t n = (size_t)W * f->bufRows;
    ResvD zr;
    memset(&zr, 0, sizeof zr);
    zr.lightId = -1;                                                 // rstr_frame_create: a zero-filled reference reservoir has no sample
    HitRec zh;
    memset(&zh, 0, sizeof zh);
    for (int i = 0; i < 2; i++) { f->geom[i].assign(n, make_float4(0, 0, 0, 0)); f->matId[i].assign(n, 0); f->resv[i].assign(n, zr); }
    f->albedoMotion.assign(n, make_float4(0, 0, 0, 0));
    // negative control of tests/test_kernels_under_sanitizers.py: a radiance plane one pixel short must be REPORTED by the sanitizer
    const char* fault = getenv("EMU_INJECT_FAULT");
    f->radiance.assign(3 * n - (fault && !strcmp(fault, "short_plane") ? 3 : 0), 0.f);
    f->resvTemp.assign(n, zr); f->resvTemp2.assign(n, zr);
    f->hit.assign(n, zh);
    if (f->sc->hs.anyMRMaps) f->hitMR.assign(n, make_float2(0.f, 0.f));
    f->queue.assign(n, 0); f->shadeQueue.assign(n, 0);
    memset(f->queueCount, 0, sizeof f->queueCount);
    return f;
}
void* emu_di_create(void* scv, int W, int H) { return emu_di_create_strip(scv, W, H, 0, 0, 0); }
void emu_di_destroy(void* f) { delete (EmuDI*)f; }

// rstr_gbuffer_render + rstr_restir_direct.  pipeline 0: staged (launchPhaseAStaged), 1: fused (launchGBufferRestirA), 2: split (launchGBuffer,
// launchRestirA), 3: split with every ray in the reference's order (RS_TRAVERSAL_EXACT: k_gbuffer_exact, k_restir_a_exact)
void emu_di_phase_b_pass(void* fv, const RstrCamera* cam, const RstrParams* prm, int iter, int pass);
void emu_di_phase_a(void* fv, const RstrCamera* cam, const RstrParams* prm, int looper, int iter, int drain, int pipeline);
void emu_di_frame(void* fv, const RstrCamera* cam, const RstrParams* prm, int looper, int iter, int drain, int pipeline) {
    emu_di_phase_a(fv, cam, prm, looper, iter, drain, pipeline);
    const int passes = (prm->reuse & 2) ? (prm->spatialPasses < 1 ? 1 : prm->spatialPasses) : 0;
    for (int pass = 1; pass <= (passes ? passes : 1); pass++) emu_di_phase_b_pass(fv, cam, prm, iter, pass);
}
// rstr_gbuffer_render + rstr_restir_phase_a on the frame's own rows
void emu_di_phase_a(void* fv, const RstrCamera* cam, const RstrParams* prm, int looper, int iter, int drain, int pipeline) {
    EmuDI* f = (EmuDI*)fv;
    const DevScene& s = f->dev;
    const RstrParams p = *prm;
    const CamDev c = toCamDev(*cam), lc = toCamDev(f->haveLast ? f->lastCamera : *cam);
    const bool sp = (p.reuse & 2) != 0;
    const int first = f->first ? 1 : 0;
    if (p.unbiased && f->hitPos.empty()) f->hitPos.assign((size_t)f->W * f->H, make_float4(0, 0, 0, 0));      // rstr_restir_phase_a
    FrameDev d = diFrameDev(f);
    d.resvStage = f->resvTemp.data();
    memset(f->queueCount, 0, sizeof f->queueCount);
    d.shadeCount = f->queueCount + 4;
    const unsigned gx = (unsigned)((f->W + 15) / 16), gy = (unsigned)((f->row1 - f->row0 + 7) / 8), linear = gx * gy;
    if (pipeline == 0) {
        d.resvStage = sp ? f->resvTemp.data() : f->resvTemp2.data();
        if (!sp) f->temp2Ready = false;
        // launchPhaseAStaged: up to RS_MAX_BANDS horizontal bands of whole 8-row tiles, each with its own shaded-pixel queue
        const int rows = d.rowHi - d.rowLo;
        int bands = f->bands < 1 ? 1 : (f->bands > RS_MAX_BANDS ? RS_MAX_BANDS : f->bands);
        const int bandRows = ((rows + bands - 1) / bands + 7) & ~7;
        if (bandRows * (bands - 1) >= rows) bands = (rows + bandRows - 1) / bandRows;
        for (int b = 0; b < bands; b++) {
            FrameDev fb = d;
            fb.rowLo = d.rowLo + b * bandRows;
            fb.rowHi = fb.rowLo + bandRows < d.rowHi ? fb.rowLo + bandRows : d.rowHi;
            fb.shadeCount = f->queueCount + 4 + 2 * b;
            fb.shadeQueue = d.shadeQueue + (size_t)(fb.rowLo - d.rowLo) * d.W;
            const unsigned by = (unsigned)((fb.rowHi - fb.rowLo + 7) / 8), lin = gx * by;
            if (sp) emuLaunch(gx, by, [&] { k_primary<true>(s, fb, c, lc, looper, iter); });
            else emuLaunch(gx, by, [&] { k_primary<false>(s, fb, c, lc, looper, iter); });
            emuLaunch(lin, 1, [&] { k_candidates(s, fb, c, p, looper); });
            if (!p.unbiased) {
                if (drain) emuLaunch(3, 1, [&] { k_shadow<true>(s, fb); });
                else emuLaunch(3, 1, [&] { k_shadow<false>(s, fb); });
            }
            if (p.unbiased) {
                if (sp) emuLaunch(lin, 1, [&] { k_temporal_unb<true>(s, fb, p, iter, first); });
                else emuLaunch(lin, 1, [&] { k_temporal_unb<false>(s, fb, p, iter, first); });
            } else if (sp) emuLaunch(lin, 1, [&] { k_temporal<true>(s, fb, p, iter, first); });
            else emuLaunch(lin, 1, [&] { k_temporal<false>(s, fb, p, iter, first); });
        }
        if (sp) emuLaunch(2, 1, [&] { k_gbuffer_restir_a_fix<true>(s, d, c, lc, p, looper, iter, first); });
        else emuLaunch(2, 1, [&] { k_gbuffer_restir_a_fix<false>(s, d, c, lc, p, looper, iter, first); });
    } else if (pipeline == 1) {
        if (sp) {
            emuLaunch(gx, gy, [&] { k_gbuffer_restir_a<true>(s, d, c, lc, p, looper, iter, first); });
            emuLaunch(2, 1, [&] { k_gbuffer_restir_a_fix<true>(s, d, c, lc, p, looper, iter, first); });
        } else {
            emuLaunch(gx, gy, [&] { k_gbuffer_restir_a<false>(s, d, c, lc, p, looper, iter, first); });
            emuLaunch(2, 1, [&] { k_gbuffer_restir_a_fix<false>(s, d, c, lc, p, looper, iter, first); });
        }
    } else if (pipeline == 2) {
        emuLaunch(gx, gy, [&] { k_gbuffer(s, d, c, lc); });
        emuLaunch(2, 1, [&] { k_gbuffer_fix(s, d, c, lc); });
        f->queueCount[0] = 0;
        if (sp) {
            emuLaunch(gx, gy, [&] { k_restir_a<true>(s, d, c, p, looper, iter, first); });
            emuLaunch(2, 1, [&] { k_restir_a_fix<true>(s, d, c, p, looper, iter, first); });
        } else {
            emuLaunch(gx, gy, [&] { k_restir_a<false>(s, d, c, p, looper, iter, first); });
            emuLaunch(2, 1, [&] { k_restir_a_fix<false>(s, d, c, p, looper, iter, first); });
        }
    } else {
        emuLaunch(gx, gy, [&] { k_gbuffer_exact(s, d, c, lc); });
        if (sp) emuLaunch(gx, gy, [&] { k_restir_a_exact<true>(s, d, c, p, looper, iter, first); });
        else emuLaunch(gx, gy, [&] { k_restir_a_exact<false>(s, d, c, p, looper, iter, first); });
    }
}
// rstr_restir_phase_b_pass: spatial pass `pass` of the frame's own rows (between the passes a strip's halo rows of the published plane
// are exchanged by the caller); the last pass swaps the history
void emu_di_phase_b_pass(void* fv, const RstrCamera* cam, const RstrParams* prm, int iter, int pass) {
    EmuDI* f = (EmuDI*)fv;
    const DevScene& s = f->dev;
    const RstrParams p = *prm;
    const bool sp = (p.reuse & 2) != 0;
    FrameDev d = diFrameDev(f);
    const unsigned gx = (unsigned)((f->W + 15) / 16), gy = (unsigned)((f->row1 - f->row0 + 7) / 8);
    const int passes = sp ? (p.spatialPasses < 1 ? 1 : p.spatialPasses) : 0;
    if (passes) {
        if (passes > 1 && !f->temp2Ready) { f->resvTemp2 = f->resvTemp; f->temp2Ready = true; }
        ResvD* buf[2] = {f->resvTemp.data(), f->resvTemp2.data()};
        const ResvD* src = buf[(pass - 1) & 1];
        ResvD* dst = buf[pass & 1];
        // negative control of tests/test_kernels_under_sanitizers.py: spatial reuse IN PLACE -- the reference's own inter-block race
        // (restir.cu:192-196: __syncthreads() is not a grid barrier) -- must be REPORTED by TSan
        const char* fault = getenv("EMU_INJECT_FAULT");
        if (fault && !strcmp(fault, "in_place_spatial")) dst = (ResvD*)src;
        const int last = pass == passes ? 1 : 0;
        if (p.unbiased) emuLaunch(gx, gy, [&] { k_restir_b_unb(s, d, p, iter, src, dst, pass, last); });
        else emuLaunch(gx, gy, [&] { k_restir_b(s, d, p, iter, src, dst, pass, last); });
    }
    if (pass >= passes) {
        f->resvOut ^= 1;
        f->first = false;
    }
}
// rstr_frame_copy_rows: rows [r0, r1) of a plane from one strip's resident rows into another's.  0 geom_cur, 1 matid_cur, 2 resv_history,
// 3 resv_temp, 4 resv_temp2, 5 resv_out
int emu_di_copy_rows(void* dv, void* sv, int plane, int r0, int r1) {
    EmuDI *dst = (EmuDI*)dv, *src = (EmuDI*)sv;
    if (r0 < src->bufRow0 || r1 > src->bufRow0 + src->bufRows || r0 < dst->bufRow0 || r1 > dst->bufRow0 + dst->bufRows || r0 >= r1) return 1;
    const size_t so = (size_t)(r0 - src->bufRow0) * src->W, dof = (size_t)(r0 - dst->bufRow0) * dst->W, cnt = (size_t)(r1 - r0) * src->W;
    auto resv = [&](EmuDI* f) -> ResvD* {
        if ((plane == 4) && !f->temp2Ready) { f->resvTemp2 = f->resvTemp; f->temp2Ready = true; }     // rsEnsureTemp2
        return plane == 2 ? f->resv[f->resvOut ^ 1].data() : plane == 3 ? f->resvTemp.data() : plane == 4 ? f->resvTemp2.data() : f->resv[f->resvOut].data();
    };
    if (plane == 0) memcpy(dst->geom[dst->cur].data() + dof, src->geom[src->cur].data() + so, cnt * sizeof(float4));
    else if (plane == 1) memcpy(dst->matId[dst->cur].data() + dof, src->matId[src->cur].data() + so, cnt * sizeof(int));
    else memcpy(resv(dst) + dof, resv(src) + so, cnt * sizeof(ResvD));
    return 0;
}
unsigned emu_di_halo_miss(void* fv) { return ((EmuDI*)fv)->counters[0]; }
// rstr_gbuffer_render + rstr_pathtrace_direct (launchGBuffer, launchPTDirect)
void emu_di_ptdirect(void* fv, const RstrCamera* cam, int looper, int iter) {
    EmuDI* f = (EmuDI*)fv;
    const DevScene& s = f->dev;
    const CamDev c = toCamDev(*cam), lc = toCamDev(f->haveLast ? f->lastCamera : *cam);
    FrameDev d = diFrameDev(f);
    const unsigned gx = (unsigned)((f->W + 15) / 16), gy = (unsigned)((f->H + 7) / 8);
    memset(f->queueCount, 0, sizeof f->queueCount);
    emuLaunch(gx, gy, [&] { k_gbuffer(s, d, c, lc); });
    emuLaunch(2, 1, [&] { k_gbuffer_fix(s, d, c, lc); });
    f->queueCount[0] = 0;
    emuLaunch(gx, gy, [&] { k_ptdirect(s, d, c, looper, iter); });
    emuLaunch(2, 1, [&] { k_ptdirect_fix(s, d, c, looper, iter); });
}
void emu_di_set_bands(void* fv, int bands) { ((EmuDI*)fv)->bands = bands; }
void emu_di_update(void* fv, const RstrCamera* cam) {
    EmuDI* f = (EmuDI*)fv;
    f->lastCamera = *cam; f->haveLast = true; f->cur ^= 1;
}
unsigned emu_di_fixup_pixels(void* fv) { return ((EmuDI*)fv)->queueCount[0]; }
// rstr_frame_read: 0 albedo 1 normal 2 matid 3 depth 4 motion 5 radiance 6 reservoir 7 reservoir_temp 8 light_index
const void* emu_di_buffer(void* fv, int which) {
    EmuDI* f = (EmuDI*)fv;
    const size_t n = (size_t)f->W * f->bufRows;                     // the resident rows (the caller cuts the strip's own rows out)
    const unsigned blocks = (unsigned)((n + RS_BLOCK - 1) / RS_BLOCK);
    const float4* g = f->geom[f->cur].data();
    const float4* am = f->albedoMotion.data();
    if (which == 5) return f->radiance.data();
    if (which == 2) return f->matId[f->cur].data();
    f->export36.resize(9 * n); f->exportI.resize(n);
    float* of = f->export36.data();
    int* oi = f->exportI.data();
    const DevScene& s = f->dev;
    const ResvD* hist = f->resv[f->resvOut ^ 1].data();
    const ResvD* temp = f->resvTemp.data();
    switch (which) {
    case 0: emuLaunch(blocks, 1, [&] { k_export_geom(g, am, of, nullptr, nullptr, nullptr, n); }); return of;
    case 1: emuLaunch(blocks, 1, [&] { k_export_geom(g, am, nullptr, of, nullptr, nullptr, n); }); return of;
    case 3: emuLaunch(blocks, 1, [&] { k_export_geom(g, am, nullptr, nullptr, of, nullptr, n); }); return of;
    case 4: emuLaunch(blocks, 1, [&] { k_export_geom(g, am, nullptr, nullptr, nullptr, oi, n); }); return oi;
    case 6: emuLaunch(blocks, 1, [&] { k_export_resv(s, hist, of, nullptr, n); }); return of;
    case 7: emuLaunch(blocks, 1, [&] { k_export_resv(s, temp, of, nullptr, n); }); return of;
    case 8: emuLaunch(blocks, 1, [&] { k_export_resv(s, hist, nullptr, oi, n); }); return oi;
    }
    return nullptr;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ the image-space filters' kernels as warps
// rstr_denoiser_* (denoise.cu) on an EmuDI frame: the launch sequences of rstr_denoiser_filter (LeveledEAWFilter::filter, denoiser.cu:463-477;
// SpatioTemporalFilter::filter, :537-564), rstr_denoiser_modulate_albedo and rstr_denoiser_next_frame.
struct EmuDenoiser {
    EmuDI* f;
    int kind, W, H;
    float sigLumin, sigNormal, sigDepth;
    std::vector<float> buf[9];
    float *colorOut, *tempColor, *accumColor[2], *accumMoment[2], *variance, *tempVariance, *filteredVariance;
    bool firstTime = true;
    int frameIdx = 0;
};

extern "C" {

void* emu_denoiser_create(void* fv, int kind, float sigLumin, float sigNormal, float sigDepth) {
    EmuDenoiser* d = new EmuDenoiser;
    d->f = (EmuDI*)fv; d->kind = kind; d->W = d->f->W; d->H = d->f->H;
    d->sigLumin = sigLumin; d->sigNormal = sigNormal; d->sigDepth = sigDepth;
    const size_t n = (size_t)d->W * d->H;
    const size_t sizes[9] = {3 * n, 3 * n, 3 * n, 3 * n, 3 * n, 3 * n, n, n, n};
    for (int i = 0; i < 9; i++) d->buf[i].assign(sizes[i], 0.f);
    d->colorOut = d->buf[0].data(); d->tempColor = d->buf[1].data(); d->accumColor[0] = d->buf[2].data(); d->accumColor[1] = d->buf[3].data();
    d->accumMoment[0] = d->buf[4].data(); d->accumMoment[1] = d->buf[5].data();
    d->variance = d->buf[6].data(); d->tempVariance = d->buf[7].data(); d->filteredVariance = d->buf[8].data();
    return d;
}
void emu_denoiser_destroy(void* d) { delete (EmuDenoiser*)d; }
void emu_denoiser_filter(void* dv, const RstrCamera* cam) {
    EmuDenoiser* d = (EmuDenoiser*)dv;
    EmuDI* f = d->f;
    DenoiseG g;
    g.W = f->W; g.H = f->H;
    g.geom = f->geom[f->cur].data(); g.geomLast = f->geom[f->cur ^ 1].data();
    g.id = f->matId[f->cur].data(); g.idLast = f->matId[f->cur ^ 1].data();
    g.albedoMotion = f->albedoMotion.data();
    const CamDev c = toCamDev(*cam);
    const size_t n = (size_t)d->W * d->H;
    const unsigned gx = (unsigned)((d->W + 15) / 16), gy = (unsigned)((d->H + 7) / 8), lin = (unsigned)((n + 255) / 256);           // 256-thread blocks, as launched
    const float* colorIn = f->radiance.data();
    if (d->kind == RSTR_DENOISER_EAW) {
        emuLaunch(gx, gy, [&] { k_eaw(g, c, d->colorOut, colorIn, d->sigDepth, d->sigNormal, d->sigLumin, 0); });
        for (int level = 1; level <= 4; level++) {
            emuLaunch(gx, gy, [&] { k_eaw(g, c, d->tempColor, d->colorOut, d->sigDepth, d->sigNormal, d->sigLumin, level); });
            std::swap(d->colorOut, d->tempColor);
        }
    } else {
        const int fi = d->frameIdx;
        const int firstTime = d->firstTime ? 1 : 0;
        emuLaunch(lin, 1, [&] { k_temporal_accumulate(g, d->accumColor[fi], d->accumColor[fi ^ 1], d->accumMoment[fi], d->accumMoment[fi ^ 1], colorIn, firstTime); }, 8);
        d->firstTime = false;
        emuLaunch(lin, 1, [&] { k_estimate_variance(d->variance, d->accumMoment[fi], d->W, d->H); }, 8);
        auto wavelet = [&](float* out, const float* in, int level) {
            emuLaunch(lin, 1, [&] { k_filter_variance(d->filteredVariance, d->variance, d->W, d->H); }, 8);
            emuLaunch(gx, gy, [&] { k_eaw_svgf(g, c, out, in, d->tempVariance, d->variance, d->filteredVariance, d->sigDepth, d->sigNormal, d->sigLumin, level); });
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
}
void emu_denoiser_next_frame(void* dv) { ((EmuDenoiser*)dv)->frameIdx ^= 1; }
void emu_denoiser_modulate_albedo(void* dv) {
    EmuDenoiser* d = (EmuDenoiser*)dv;
    const size_t n = (size_t)d->W * d->H;
    const float4* am = d->f->albedoMotion.data();
    emuLaunch((unsigned)((n + 255) / 256), 1, [&] { k_modulate(d->colorOut, am, n); }, 8);
}
const float* emu_denoiser_color(void* dv) { return ((EmuDenoiser*)dv)->colorOut; }
const float* emu_denoiser_variance(void* dv) { return ((EmuDenoiser*)dv)->variance; }

}  // extern "C"

// ------------------------------------------------------------------------------------------------ the device-side tree build as warps / blocks
// rstr_scene_build_traced_gpu (bvh_gpu.cu) on an emulated scene: the same kernels in the same order; cub's radix sort and exclusive scan
// (library code) are std::stable_sort by key and a serial scan.  The new tree replaces the host-built one the scene's DevScene points at.
extern "C" int emu_scene_build_traced(void* scv, int mode) {
    EmuScene* sc = (EmuScene*)scv;
    HostScene& hs = sc->hs;
    const int T = hs.T;
    if (T < 8 || (mode != 0 && mode != 1)) return 1;
    const size_t N2 = 2 * (size_t)T - 1;
    std::vector<unsigned long long> keys(T), keysSorted(T);
    std::vector<unsigned int> slots(T), slotsSorted(T), arrived(T, 0u);
    std::vector<int> left(N2, 0), right(N2, 0), parent(N2, 0), count(N2, 0), firstPos(N2, 0), cidA(T), cidB(T), cidTmp(T), keep(T), pos(T), nearest(T);
    std::vector<float> cost(N2, 0.f);
    std::vector<GBox> box(N2);
    int scalars[4] = {0, 0, 0, 0};
    float bounds[6] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
    std::vector<FastNode> nodes(T - 1);
    std::vector<TriGeom> tgNew(T);
    std::vector<int> primToFast(T);
    const float4* tg = (const float4*)hs.fastTris.data();
    const int B = 256, G = (T + B - 1) / B;
    emuLaunchBlock((unsigned)std::min(G, 1184), [&] { k_bvh_scene_bounds(tg, T, bounds); });
    emuLaunchBlock((unsigned)G, [&] { k_bvh_morton(tg, T, bounds, keys.data(), slots.data()); });
    {                                                                // cub::DeviceRadixSort::SortPairs: stable, ascending keys
        std::vector<int> order(T);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return keys[a] < keys[b]; });
        for (int i = 0; i < T; i++) { keysSorted[i] = keys[order[i]]; slotsSorted[i] = slots[order[i]]; }
    }
    int root = T;
    if (mode == 1) {
        emuLaunchBlock((unsigned)G, [&] { k_bvh_radix_tree(keysSorted.data(), T, left.data(), right.data(), parent.data()); });
        emuLaunchBlock((unsigned)G, [&] { k_bvh_bottom_up(tg, slotsSorted.data(), T, left.data(), right.data(), parent.data(), box.data(), count.data(), cost.data(), arrived.data()); });
    } else {
        emuLaunchBlock((unsigned)G, [&] { k_ploc_init(tg, slotsSorted.data(), T, box.data(), count.data(), cost.data(), parent.data(), cidA.data()); });
        int n = T;
        int *cur = cidA.data(), *nxt = cidB.data();
        while (n > 1) {
            const int g = (n + B - 1) / B;
            emuLaunchBlock((unsigned)g, [&] { k_ploc_nearest(cur, n, box.data(), nearest.data()); });
            emuLaunchBlock((unsigned)g, [&] { k_ploc_merge(cur, n, nearest.data(), T, box.data(), count.data(), cost.data(), left.data(), right.data(), parent.data(), scalars, cidTmp.data(), keep.data()); });
            int run = 0;                                             // cub::DeviceScan::ExclusiveSum
            for (int i = 0; i < n; i++) { pos[i] = run; run += keep[i]; }
            emuLaunchBlock((unsigned)g, [&] { k_ploc_compact(cidTmp.data(), keep.data(), pos.data(), n, nxt, scalars + 1); });
            const int n2 = scalars[1];
            if (n2 >= n) return 2;
            n = n2;
            std::swap(cur, nxt);
        }
        root = cur[0];
    }
    const int G2 = (int)((N2 + B - 1) / B);
    emuLaunchBlock((unsigned)G2, [&] { k_bvh_first(T, (int)N2, left.data(), right.data(), parent.data(), count.data(), firstPos.data(), scalars + 2); });
    emuLaunchBlock((unsigned)G, [&] { k_bvh_emit(T, left.data(), right.data(), box.data(), count.data(), cost.data(), firstPos.data(), nodes.data()); });
    emuLaunchBlock((unsigned)G, [&] { k_bvh_reorder(tg, slotsSorted.data(), firstPos.data(), T, (float4*)tgNew.data(), primToFast.data()); });
    const int depth = scalars[2];
    if (depth + 1 > RS_PACKET_STACK) return 3;
    const GBox rootBox = box[root];
    hs.fastNodes.swap(nodes); hs.fastTris.swap(tgNew); hs.primToFast.swap(primToFast);      // swap the new tree in
    DevScene& d = sc->dev;
    d.fastNodes = (const float4*)hs.fastNodes.data(); d.triGeom = (const float4*)hs.fastTris.data(); d.primToFast = hs.primToFast.data();
    d.numFastNodes = T - 1; d.fastRoot = root - T;
    for (int a = 0; a < 3; a++) {
        const float pad = 4e-6f * fmaxf(fabsf(rootBox.lo[a]), fabsf(rootBox.hi[a])) + 1e-7f;
        d.fastRootMin[a] = rootBox.lo[a] - pad; d.fastRootMax[a] = rootBox.hi[a] + pad;
    }
    return 0;
}
