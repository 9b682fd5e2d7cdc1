// capi_internal.h -- state behind the opaque handles of the C ABI, shared by capi.cu and strip_group.cu.
#pragma once

#include <stdio.h>
#include <string.h>

#include <atomic>
#include <string>
#include <vector>

#include "kernels.h"

int rsFail(int code, const std::string& msg);     // records the message for rstr_last_error() and returns code
void rsCountLaunches(int n);                       // rstr_launch_count bookkeeping
int rsSmCount();                                   // SMs of the current device (persistent kernels' grid)

#define CU(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char buf_[512];                                                                        \
            snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return rsFail(RSTR_ERR_CUDA, buf_);                                                    \
        }                                                                                          \
    } while (0)

struct RstrScene {
    rs::HostScene hs;
    rs::DevScene dev{};
    void* dNodes = nullptr; void* dTriGeom = nullptr; void* dTriNorm = nullptr;
    void* dFastNodes = nullptr; void* dPrimToFast = nullptr; void* dFallback = nullptr; void* dRank = nullptr;
    void* dMaterials = nullptr; void* dAlias = nullptr; void* dLights = nullptr;
    void* dTexData = nullptr; void* dTexInfo = nullptr; void* dTriUV = nullptr; void* dEnvAlias = nullptr; void* dEnvDir = nullptr;
    size_t deviceBytes = 0;
    bool uploaded = false;      // set only after every device array exists and DevScene is filled
    int traversalMode = RS_TRAVERSAL_FAST;
    int gpuTreeDepth = 0;       // > 0: the traced tree on the device was built by rstr_scene_build_traced_gpu (bvh_gpu.cu)
    float gpuTreeMs = 0.f;
};

// Planes that a neighbouring strip reads (halo rows) live in ONE device allocation, the "exchange slab": a 256-byte header of
// 32-bit flags followed by geom[2], matId[2], resv[2], resvTemp, resvTemp2.  One CUDA IPC handle then maps everything a
// peer GPU has to store into (strip_group.cu).
#define RS_SLAB_HEADER 256
enum { RS_XP_GEOM0 = 0, RS_XP_GEOM1, RS_XP_MATID0, RS_XP_MATID1, RS_XP_RESV0, RS_XP_RESV1, RS_XP_TEMP, RS_XP_TEMP2, RS_XP_COUNT };

struct RstrFrame {
    RstrScene* sc = nullptr;
    int W = 0, H = 0, row0 = 0, row1 = 0, halo = 0, bufRow0 = 0, bufRows = 0;
    size_t nBuf = 0;
    void* slab = nullptr; size_t slabBytes = 0; size_t slabOff[RS_XP_COUNT] = {};
    float4* geom[2] = {nullptr, nullptr};
    int* matId[2] = {nullptr, nullptr};
    float4* albedoMotion = nullptr;
    float* radiance = nullptr;
    rs::ResvD* resv[2] = {nullptr, nullptr};
    rs::ResvD* resvTemp = nullptr;
    rs::ResvD* resvTemp2 = nullptr;    // second publication buffer (spatialPasses > 1)
    bool temp2Ready = false;           // resvTemp2 holds a copy of resvTemp's initial content
    rs::HitRec* hit = nullptr;
    float4* hitPos = nullptr;      // unbiased mode: shaded points (allocated on first use)
    float2* hitMR = nullptr;       // allocated when the scene has metallic / roughness maps
    uchar4* ldr = nullptr;
    uchar4* ldrB[RSTR_LDR_SLOTS] = {};        // LDR frames in flight of the pipelined host call
    cudaStream_t copyStream = nullptr;
    cudaEvent_t evRendered[RSTR_LDR_SLOTS] = {}, evCopied[RSTR_LDR_SLOTS] = {};
    bool slotBusy[RSTR_LDR_SLOTS] = {};
    unsigned int* haloMiss = nullptr;
    unsigned long long* rowCost = nullptr;    // allocated by rstr_frame_row_cost
    int* queue = nullptr;
    int* shadeQueue = nullptr;     // staged phase A: compact list of shaded pixels
    rs::StagedStreams ss{};        // side streams / events of the banded staged pipeline (created on first use)
    int bands = 1;                 // row bands of the staged pipeline (bands > 1: a band's queue kernels overlap the next band's k_primary; measured slower)
    int staged = -1;               // phase A as the staged pipeline (kernels.cu): 1, as the single fused kernel: 0, by scene size: -1
    unsigned int* queueCount = nullptr;
    void* scratch = nullptr; size_t scratchBytes = 0;
    int cur = 0;        // GBuffer::frameIdx
    int resvOut = 0;    // which of resv[] is devDirectReservoir (written this frame)
    RstrCamera lastCamera{};
    bool haveLast = false;
    bool first = true;  // ReSTIRFirstFrame
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[2 * RSTR_T_COUNT] = {};
    cudaEvent_t xfer = nullptr;
    cudaEvent_t marks[8] = {};
    bool ownStream = true;
    bool fuse = true;              // G-buffer + phase A as one kernel when both are asked for with the same camera
    bool gbufPending = false;      // rstr_gbuffer_render was called and its launch is deferred to the next phase A (or flushed)
    RstrCamera pendCam{};
    rs::CamDev pendC{}, pendLC{};
    bool renderHalo = true;        // strip frames: G-buffer halo rows rendered locally (true) or received from the neighbours
    bool ran[RSTR_T_COUNT] = {};
};

rs::FrameDev rsToFrameDev(const RstrFrame* f, int rowLo, int rowHi);
rs::CamDev rsToCamDev(const RstrCamera& c);
extern "C" int rsFlushGBuffer(RstrFrame* f);     // launches a G-buffer render that rstr_gbuffer_render deferred
extern "C" int rsEnsureTemp2(RstrFrame* f);
int rsEnsureUploaded(RstrScene* sc);
         // uploads the host scene on first use (capi.cu)
