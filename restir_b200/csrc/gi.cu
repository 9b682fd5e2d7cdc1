// gi.cu -- C ABI of ReSTIR GI: ReSTIRIndirect (restir.cu:448-476) and the indirect-reservoir share of
// ReSTIRInit / ReSTIRFree / ReSTIRReset (restir.cu:491-496, 511-512, 516).  The kernels are in gi_kernels.inl (kernels.cu).
// The reference ships the call commented out (main.cpp:168) with Settings::traceDepth = 0; here it is reachable through rstr_gi_*.
#include <stdlib.h>
#include <string.h>

#include <utility>

#include "capi_internal.h"

using namespace rs;

struct RstrGI {
    RstrFrame* f = nullptr;
    int W = 0, H = 0;
    float4* resv[2] = {nullptr, nullptr};      // 64-byte records
    float* nsz[2] = {nullptr, nullptr};        // ns.z plane
    float* indirect = nullptr;                 // devIndirectIllum (own plane)
    float* scratch = nullptr;                  // export buffer (17 floats per pixel), allocated on first read
    unsigned int* fallback = nullptr;
    int out = 0;                               // which of resv[] is devIndTemporalReservoir (written by the next call)
    bool first = true;                         // ReSTIRFirstFrame
    int bounceWalk = RS_TRAVERSAL_FAST;
    int pipeline = RSTR_GI_PIPELINE_AUTO;
    // staged pipeline (allocated on first use): per-pixel hand-over planes, pixel status, two path queues, per-depth path counts
    float4* pix = nullptr;
    int* pixStatus = nullptr;
    float4* pathQ[2] = {nullptr, nullptr};
    unsigned int* pathCount = nullptr;
    // ray-queue pipeline (on top of the staged buffers): slot lists of the two walkers + their counters, hits, occlusion flags
    unsigned int* closestList = nullptr;
    unsigned int* shadowList = nullptr;
    unsigned int* walkCount = nullptr;
    float4* hit = nullptr;
    int* occ = nullptr;
};

#define RS_GI_MAX_DEPTH 64

static void giFree(RstrGI* g) {
    void* all[] = {g->resv[0], g->resv[1], g->nsz[0], g->nsz[1], g->indirect, g->scratch, g->fallback, g->pix, g->pixStatus, g->pathQ[0], g->pathQ[1], g->pathCount,
                   g->closestList, g->shadowList, g->walkCount, g->hit, g->occ};
    for (void* p : all) cudaFree(p);
    delete g;
}

extern "C" {

int rstr_gi_create(RstrFrame* f, RstrGI** out) {
    if (!f || !out) return rsFail(RSTR_ERR_ARG, "rstr_gi_create: bad argument");
    if (f->row0 != 0 || f->row1 != f->H || f->bufRow0 != 0) return rsFail(RSTR_ERR_ARG, "rstr_gi_create: ReSTIR GI works on full frames, not on strips");
    RstrGI* g = new RstrGI;
    g->f = f; g->W = f->W; g->H = f->H;
    const size_t n = (size_t)f->W * f->H;
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 2 && e == cudaSuccess; i++) {
        e = cudaMalloc(&g->resv[i], n * 64);
        if (e == cudaSuccess) e = cudaMemset(g->resv[i], 0, n * 64);                 // restir.cu:493-496
        if (e == cudaSuccess) e = cudaMalloc(&g->nsz[i], n * 4);
        if (e == cudaSuccess) e = cudaMemset(g->nsz[i], 0, n * 4);
    }
    if (e == cudaSuccess) e = cudaMalloc(&g->indirect, n * 12);
    if (e == cudaSuccess) e = cudaMemset(g->indirect, 0, n * 12);
    if (e == cudaSuccess) e = cudaMalloc(&g->fallback, 4);
    if (e == cudaSuccess) e = cudaMemset(g->fallback, 0, 4);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();       // the fills above are not ordered with the frame's (non-blocking) stream
    if (e != cudaSuccess) { giFree(g); return rsFail(RSTR_ERR_CUDA, std::string("rstr_gi_create: ") + cudaGetErrorString(e)); }
    if (const char* env = getenv("RSTR_GI_PIPELINE")) {                               // A/B switch for measurements; rstr_gi_set_pipeline overrides
        if (!strcmp(env, "staged")) g->pipeline = RSTR_GI_PIPELINE_STAGED;
        else if (!strcmp(env, "fused")) g->pipeline = RSTR_GI_PIPELINE_FUSED;
        else if (!strcmp(env, "queued")) g->pipeline = RSTR_GI_PIPELINE_QUEUED;
        else if (!strcmp(env, "auto")) g->pipeline = RSTR_GI_PIPELINE_AUTO;
    }
    *out = g;
    return RSTR_OK;
}

int rstr_gi_destroy(RstrGI* g) {
    if (!g) return RSTR_OK;
    cudaStreamSynchronize(g->f->stream);
    giFree(g);
    return RSTR_OK;
}

int rstr_gi_reset(RstrGI* g) {
    if (!g) return rsFail(RSTR_ERR_ARG, "rstr_gi_reset: null handle");
    g->first = true;
    return RSTR_OK;
}

int rstr_gi_set_bounce_walk(RstrGI* g, int traversal) {
    if (!g || (traversal != RS_TRAVERSAL_FAST && traversal != RS_TRAVERSAL_EXACT)) return rsFail(RSTR_ERR_ARG, "rstr_gi_set_bounce_walk: bad argument");
    g->bounceWalk = traversal;
    return RSTR_OK;
}

int rstr_gi_set_pipeline(RstrGI* g, int pipeline) {
    if (!g || (pipeline < RSTR_GI_PIPELINE_FUSED || pipeline > RSTR_GI_PIPELINE_AUTO)) return rsFail(RSTR_ERR_ARG, "rstr_gi_set_pipeline: bad argument");
    g->pipeline = pipeline;
    return RSTR_OK;
}

// hand-over planes and path queues of the staged pipeline: 112 + 4 + 2 x 96 bytes per pixel
static int giEnsureStaged(RstrGI* g) {
    if (g->pix) return RSTR_OK;
    const size_t n = (size_t)g->W * g->H;
    cudaError_t e = cudaMalloc(&g->pix, n * 7 * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc(&g->pixStatus, n * sizeof(int));
    for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaMalloc(&g->pathQ[i], n * 6 * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc(&g->pathCount, (RS_GI_MAX_DEPTH + 2) * sizeof(unsigned int));
    if (e != cudaSuccess) {
        void* all[] = {g->pix, g->pixStatus, g->pathQ[0], g->pathQ[1], g->pathCount};
        for (void* p : all) cudaFree(p);
        g->pix = nullptr; g->pixStatus = nullptr; g->pathQ[0] = g->pathQ[1] = nullptr; g->pathCount = nullptr;
        return rsFail(RSTR_ERR_CUDA, std::string("rstr_restir_indirect: staged buffers: ") + cudaGetErrorString(e));
    }
    return RSTR_OK;
}

// slot lists, hits and occlusion flags of the ray-queue pipeline: 4 + 4 + 16 + 4 bytes per pixel
static int giEnsureQueued(RstrGI* g) {
    if (g->hit) return RSTR_OK;
    const size_t n = (size_t)g->W * g->H;
    cudaError_t e = cudaMalloc(&g->closestList, n * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc(&g->shadowList, n * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc(&g->walkCount, 4 * (RS_GI_MAX_DEPTH + 2) * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc(&g->occ, n * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&g->hit, n * sizeof(float4));
    if (e != cudaSuccess) {
        void* all[] = {g->closestList, g->shadowList, g->walkCount, g->occ, g->hit};
        for (void* p : all) cudaFree(p);
        g->closestList = g->shadowList = g->walkCount = nullptr; g->occ = nullptr; g->hit = nullptr;
        return rsFail(RSTR_ERR_CUDA, std::string("rstr_restir_indirect: ray-queue buffers: ") + cudaGetErrorString(e));
    }
    return RSTR_OK;
}

int rstr_restir_indirect(RstrGI* g, const RstrCamera* cam, int looper, int iter, int traceDepth, int reuse, int target) {
    if (!g || !cam) return rsFail(RSTR_ERR_ARG, "rstr_restir_indirect: bad argument");
    RstrFrame* f = g->f;
    if (cam->resolution[0] != f->W || cam->resolution[1] != f->H) return rsFail(RSTR_ERR_ARG, "rstr_restir_indirect: camera resolution differs from the frame");
    if (traceDepth < 0 || traceDepth > RS_GI_MAX_DEPTH || iter < 0) return rsFail(RSTR_ERR_ARG, "rstr_restir_indirect: traceDepth must be 0..64 and iter >= 0");
    if (target != RSTR_GI_TARGET_OWN && target != RSTR_GI_TARGET_RADIANCE) return rsFail(RSTR_ERR_ARG, "rstr_restir_indirect: unknown target");
    { int rc = rsFlushGBuffer(f); if (rc) return rc; }
    const FrameDev d = rsToFrameDev(f, 0, f->H);
    GIDev gd{};
    gd.resvOut = g->resv[g->out]; gd.resvIn = g->resv[g->out ^ 1];
    gd.nszOut = g->nsz[g->out]; gd.nszIn = g->nsz[g->out ^ 1];
    gd.indirect = target == RSTR_GI_TARGET_RADIANCE ? f->radiance : g->indirect;
    gd.fallback = g->fallback;
    gd.maxDepth = traceDepth; gd.reuse = reuse; gd.first = g->first ? 1 : 0; gd.iter = iter;
    gd.bounceWalk = g->bounceWalk;
    // auto: like the direct path's rule (capi.cu): the ray-queue form pays for its extra launches once rays diverge, i.e. on real scenes
    // (200 k / 1 M triangles: 4.27 / 3.58 ms against 4.90 / 3.87 staged); a Cornell box is faster staged (0.88 against 1.07 ms).
    // profiles/r02_c35_gi_bench.jsonl
    const int pipeline = g->pipeline != RSTR_GI_PIPELINE_AUTO ? g->pipeline
                         : f->sc->dev.numFastNodes > 1024 ? RSTR_GI_PIPELINE_QUEUED : RSTR_GI_PIPELINE_STAGED;
    if (pipeline != RSTR_GI_PIPELINE_FUSED && f->sc->dev.traversal != RS_TRAVERSAL_EXACT) {       // the validation mode has one form
        int rc = giEnsureStaged(g);
        if (rc) return rc;
        gd.pix = g->pix; gd.pixStatus = g->pixStatus; gd.pathQ[0] = g->pathQ[0]; gd.pathQ[1] = g->pathQ[1]; gd.pathCount = g->pathCount;
        gd.pixStride = (size_t)g->W * g->H;
        if (pipeline == RSTR_GI_PIPELINE_QUEUED && g->bounceWalk == RS_TRAVERSAL_FAST) {      // the walkers walk the traced tree
            rc = giEnsureQueued(g);
            if (rc) return rc;
            gd.closestList = g->closestList; gd.shadowList = g->shadowList; gd.walkCount = g->walkCount; gd.hit = g->hit; gd.occ = g->occ;
        }
    }
    rsCountLaunches(launchRestirIndirect(f->sc->dev, d, rsToCamDev(*cam), gd, looper, rsSmCount(), f->stream));
    g->out ^= 1;                                                                     // std::swap(devIndTemporalReservoir, devIndLastTemporalReservoir), restir.cu:463
    g->first = false;
    CU(cudaGetLastError());
    return RSTR_OK;
}

int rstr_gi_read(RstrGI* g, float* indirectRgb, void* reservoirs) {
    if (!g) return rsFail(RSTR_ERR_ARG, "rstr_gi_read: null handle");
    const size_t n = (size_t)g->W * g->H;
    cudaStream_t st = g->f->stream;
    if (reservoirs) {
        if (!g->scratch) CU(cudaMalloc(&g->scratch, n * 68));
        launchExportGI(g->resv[g->out ^ 1], g->nsz[g->out ^ 1], g->scratch, n, st);   // the reservoirs the last call wrote
        rsCountLaunches(1);
        CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(st));
    if (indirectRgb) CU(cudaMemcpy(indirectRgb, g->indirect, n * 12, cudaMemcpyDeviceToHost));
    if (reservoirs) CU(cudaMemcpy(reservoirs, g->scratch, n * 68, cudaMemcpyDeviceToHost));
    return RSTR_OK;
}

int rstr_gi_indirect_device(RstrGI* g, float** dev) {
    if (!g || !dev) return rsFail(RSTR_ERR_ARG, "rstr_gi_indirect_device: bad argument");
    *dev = g->indirect;
    return RSTR_OK;
}

int rstr_gi_fallback_pixels(RstrGI* g, unsigned int* count, int reset) {
    if (!g || !count) return rsFail(RSTR_ERR_ARG, "rstr_gi_fallback_pixels: bad argument");
    CU(cudaStreamSynchronize(g->f->stream));
    CU(cudaMemcpy(count, g->fallback, 4, cudaMemcpyDeviceToHost));
    if (reset) CU(cudaMemset(g->fallback, 0, 4));
    return RSTR_OK;
}

}  // extern "C"
