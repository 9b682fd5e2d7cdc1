// kernels.h -- launchers of the sm_100a kernels (kernels.cu); the int-returning ones report how many kernels they launched
#pragma once

#include "device_types.h"

namespace rs {

int launchGBuffer(const DevScene& s, const FrameDev& f, const CamDev& cam, const CamDev& lastCam, cudaStream_t st);
int launchRestirA(const DevScene& s, const FrameDev& f, const CamDev& cam, const RstrParams& p, int looper, int iter, int first, cudaStream_t st);
int launchGBufferRestirA(const DevScene& s, const FrameDev& f, const CamDev& cam, const CamDev& lastCam, const RstrParams& p, int looper, int iter, int first, cudaStream_t st);
// side streams / events of the banded staged pipeline (owned by the frame)
#define RS_MAX_BANDS 8
struct StagedStreams {
    int bands;
    cudaStream_t side[2];
    cudaEvent_t evPrimary[RS_MAX_BANDS], evDone[RS_MAX_BANDS];
};
int launchPhaseAStaged(const DevScene& s, const FrameDev& f, const CamDev& cam, const CamDev& lastCam, const RstrParams& p, int looper, int iter, int first, int numSMs,
                       cudaStream_t st, const StagedStreams& ss);
void launchRestirB(const DevScene& s, const FrameDev& f, const RstrParams& p, int iter, const ResvD* src, ResvD* dst, int pass, int last, cudaStream_t st);
int launchPTDirect(const DevScene& s, const FrameDev& f, const CamDev& cam, int looper, int iter, cudaStream_t st);
void launchTonemap(const float* radiance, uchar4* ldr, size_t n, int toneMapping, float scale, cudaStream_t st);
void launchExportGeom(const float4* geom, const float4* am, float* albedo, float* normal, float* depth, int* motion, size_t n, cudaStream_t st);
void launchExportResv(const DevScene& s, const ResvD* src, float* out36, int* lightIdx, size_t n, cudaStream_t st);
// ReSTIR GI (gi_kernels.inl)
int launchRestirIndirect(const DevScene& s, const FrameDev& f, const CamDev& cam, const GIDev& g, int looper, int numSMs, cudaStream_t st);
void launchExportGI(const float4* rec, const float* nsz, float* out17, size_t n, cudaStream_t st);

}  // namespace rs
