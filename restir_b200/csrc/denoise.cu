// denoise.cu -- the reference's image-space filters over a frame's G-buffer (denoiser.cu:25-567):
// the edge-avoiding a-trous wavelet filter (LeveledEAWFilter), its SVGF form with temporal accumulation and variance guidance
// (SpatioTemporalFilter), and the albedo modulation / image sums around them.  The reference creates these filters
// (main.cpp:78-80) but never calls them from runCuda (main.cpp:146-185): they are reachable here through rstr_denoiser_*.
//
// Every kernel is a gather over planes the frame already holds -- {normal, depth} as one float4, the material / primitive id,
// {albedo, motion} -- one thread per pixel, 25 (or 9) taps at stride 2^level: L2-resident stencils bounded by HBM / L2
// bandwidth, nothing to share between threads beyond what the caches do (taps of level >= 2 do not even share sectors).
// Positions are rebuilt from depth as the reference does (DENOISER_ENCODE_POSITION, common.h:10; Camera::getPosition).
#include "capi_internal.h"
#include "camera_dev.h"

using namespace rs;

namespace {

__constant__ float cGauss3[3][3] = {{.075f, .124f, .075f}, {.124f, .204f, .124f}, {.075f, .124f, .075f}};                  // denoiser.cu:11-15
__constant__ float cGauss5[5][5] = {{.0030f, .0133f, .0219f, .0133f, .0030f}, {.0133f, .0596f, .0983f, .0596f, .0133f},   // denoiser.cu:17-23
                                    {.0219f, .0983f, .1621f, .0983f, .0219f}, {.0133f, .0596f, .0983f, .0596f, .0133f},
                                    {.0030f, .0133f, .0219f, .0133f, .0030f}};

struct DenoiseG {                 // the G-buffer planes the filters read (full frames: plane index = y * W + x)
    int W, H;
    const float4* geom;           // {n.xyz, depth}, current frame
    const float4* geomLast;
    const int* id;                // gBuffer.primId()
    const int* idLast;
    const float4* albedoMotion;   // {albedo.xyz, motion (int bits)}
};

RS_D f3 load3(const float* p, size_t i) { return mk3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
RS_D void store3(float* p, size_t i, f3 v) { p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z; }

// waveletFilter (denoiser.cu:64-134)
__global__ void __launch_bounds__(128) k_eaw(const __grid_constant__ DenoiseG g, const __grid_constant__ CamDev cam, float* __restrict__ out,
                                             const float* __restrict__ in, float sigDepth, float sigNormal, float sigLumin, int level) {
    const int x = blockIdx.x * 16 + (threadIdx.x & 15), y = blockIdx.y * 8 + (threadIdx.x >> 4);
    if (x >= g.W || y >= g.H) return;
    const int step = 1 << level;
    const size_t p = (size_t)y * g.W + x;
    const int idP = g.id[p];
    const f3 colorP = load3(in, p);
    if (idP <= -1) { store3(out, p, colorP); return; }
    const float4 gp = g.geom[p];
    const f3 normP = mk3(gp.x, gp.y, gp.z), posP = cameraPosition(cam, x, y, gp.w);
    f3 sum = mk3(0.f);
    float sumWeight = 0.f;
#pragma unroll
    for (int i = -2; i <= 2; i++)
#pragma unroll
        for (int j = -2; j <= 2; j++) {
            const int qx = x + j * step, qy = y + i * step;
            if (qx >= g.W || qy >= g.H || qx < 0 || qy < 0) continue;
            const size_t q = (size_t)qy * g.W + qx;
            if (g.id[q] != idP) continue;
            const float4 gq = g.geom[q];
            const f3 normQ = mk3(gq.x, gq.y, gq.z), colorQ = load3(in, q), posQ = cameraPosition(cam, qx, qy, gq.w);
            const float wColor = fminf(1.f, expf(-dot(colorP - colorQ, colorP - colorQ) / sigLumin));
            const float wNorm = fminf(1.f, expf(-dot(normP - normQ, normP - normQ) / sigNormal));
            const float wPos = fminf(1.f, expf(-dot(posP - posQ, posP - posQ) / sigDepth));
            const float weight = wColor * wNorm * wPos * cGauss5[i + 2][j + 2];
            sum = sum + colorQ * weight;
            sumWeight += weight;
        }
    store3(out, p, sumWeight == 0.f ? colorP : sum / sumWeight);
}

// waveletFilter, SVGF form: colour and variance together (denoiser.cu:139-216)
__global__ void __launch_bounds__(128) k_eaw_svgf(const __grid_constant__ DenoiseG g, const __grid_constant__ CamDev cam, float* __restrict__ out,
                                                  const float* __restrict__ in, float* __restrict__ varOut, const float* __restrict__ varIn,
                                                  const float* __restrict__ varFiltered, float sigDepth, float sigNormal, float sigLumin, int level) {
    const int x = blockIdx.x * 16 + (threadIdx.x & 15), y = blockIdx.y * 8 + (threadIdx.x >> 4);
    if (x >= g.W || y >= g.H) return;
    const int step = 1 << level;
    const size_t p = (size_t)y * g.W + x;
    const int idP = g.id[p];
    const f3 colorP = load3(in, p);
    if (idP <= -1) { store3(out, p, colorP); varOut[p] = varIn[p]; return; }
    const float4 gp = g.geom[p];
    const f3 normP = mk3(gp.x, gp.y, gp.z), posP = cameraPosition(cam, x, y, gp.w);
    const float lumP = luminance(colorP);
    f3 sumColor = mk3(0.f);
    float sumVariance = 0.f, sumWeight = 0.f, sumWeight2 = 0.f;
#pragma unroll
    for (int i = -2; i <= 2; i++)
#pragma unroll
        for (int j = -2; j <= 2; j++) {
            const int qx = x + j * step, qy = y + i * step;
            if (qx >= g.W || qy >= g.H || qx < 0 || qy < 0) continue;
            const size_t q = (size_t)qy * g.W + qx;
            if (g.id[q] != idP) continue;
            const float4 gq = g.geom[q];
            const f3 normQ = mk3(gq.x, gq.y, gq.z), colorQ = load3(in, q), posQ = cameraPosition(cam, qx, qy, gq.w);
            const float varQ = varIn[q];
            const float wPos = expf(-dot(posP - posQ, posP - posQ) / sigDepth) + 1e-4f;
            const float wNorm = powf(satDot(normP, normQ), sigNormal) + 1e-4f;
            const float denom = sigLumin * sqrtf(fmaxf(varFiltered[q], 0.f)) + 1e-4f;
            const float wColor = expf(-fabsf(lumP - luminance(colorQ)) / denom) + 1e-4f;
            const float weight = wColor * wNorm * wPos * cGauss5[i + 2][j + 2];
            const float weight2 = weight * weight;
            sumColor = sumColor + colorQ * weight;
            sumVariance += varQ * weight2;
            sumWeight += weight;
            sumWeight2 += weight2;
        }
    store3(out, p, sumWeight < FLT_EPSILON ? colorP : sumColor / sumWeight);
    varOut[p] = sumWeight2 < FLT_EPSILON ? varIn[p] : sumVariance / sumWeight2;
}

// temporalAccumulate (denoiser.cu:250-305): exponential history (alpha .2) of colour and of the luminance moments, re-projected
__global__ void __launch_bounds__(256) k_temporal_accumulate(const __grid_constant__ DenoiseG g, float* __restrict__ colorOut, const float* __restrict__ colorLast,
                                                             float* __restrict__ momentOut, const float* __restrict__ momentLast,
                                                             const float* __restrict__ colorIn, int first) {
    const size_t p = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (p >= (size_t)g.W * g.H) return;
    const float Alpha = .2f;
    const int id = g.id[p];
    const int lastIdx = __float_as_int(g.albedoMotion[p].w);
    bool diff = first != 0;
    if (lastIdx < 0) diff = true;
    else if (id <= -1) diff = true;
    else if (g.idLast[lastIdx] != id) diff = true;
    else {
        const float4 a = g.geom[p], b = g.geomLast[lastIdx];
        if (fabsf(dot(mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z))) < .1f) diff = true;
    }
    const f3 color = load3(colorIn, p);
    const float lum = luminance(color);
    f3 accumColor, accumMoment;
    if (diff) {
        accumColor = color;
        accumMoment = mk3(lum, lum * lum, 0.f);
    } else {
        // (the reference reads devColorAccumIn[lastIdx] before the branch, lastIdx < 0 included; the value is unused then)
        const f3 lastColor = load3(colorLast, (size_t)lastIdx), lastMoment = load3(momentLast, (size_t)lastIdx);
        accumColor = mix(lastColor, color, Alpha);
        accumMoment = mk3(mixf(lastMoment.x, lum, Alpha), mixf(lastMoment.y, lum * lum, Alpha), lastMoment.z + 1.f);
    }
    store3(colorOut, p, accumColor);
    store3(momentOut, p, accumMoment);
}

// estimateVariance (denoiser.cu:307-343): temporal variance after 4 accumulated frames, else the 3x3 spatial estimate
__global__ void __launch_bounds__(256) k_estimate_variance(float* __restrict__ variance, const float* __restrict__ moment, int W, int H) {
    const size_t p = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (p >= (size_t)W * H) return;
    const int x = (int)(p % W), y = (int)(p / W);
    const f3 m = load3(moment, p);
    if (m.z > 3.5f) { variance[p] = m.y - m.x * m.x; return; }
    float sx = 0.f, sy = 0.f;
    int num = 0;
    for (int i = -1; i <= 1; i++)
        for (int j = -1; j <= 1; j++) {
            const int qx = x + j, qy = y + i;
            if (qx < 0 || qx >= W || qy < 0 || qy >= H) continue;
            const size_t q = (size_t)qy * W + qx;
            sx += moment[3 * q]; sy += moment[3 * q + 1];
            num++;
        }
    sx /= (float)num; sy /= (float)num;
    variance[p] = sy - sx * sx;
}

// filterVariance (denoiser.cu:345-370)
__global__ void __launch_bounds__(256) k_filter_variance(float* __restrict__ out, const float* __restrict__ in, int W, int H) {
    const size_t p = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (p >= (size_t)W * H) return;
    const int x = (int)(p % W), y = (int)(p / W);
    float sum = 0.f, sumWeight = 0.f;
#pragma unroll
    for (int i = -1; i <= 1; i++)
#pragma unroll
        for (int j = -1; j <= 1; j++) {
            const int qx = x + i, qy = y + j;
            if (qx < 0 || qx >= W || qy < 0 || qy >= H) continue;
            const float weight = cGauss3[i + 1][j + 1];
            sum += in[(size_t)qy * W + qx] * weight;
            sumWeight += weight;
        }
    out[p] = sum / sumWeight;
}

// modulate (denoiser.cu:218-228): LDRToHDR (mathUtil.h:40-43), times the albedo
__global__ void __launch_bounds__(256) k_modulate(float* __restrict__ image, const float4* __restrict__ albedoMotion, size_t n) {
    const size_t p = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (p >= n) return;
    f3 c = load3(image, p);
    c = c / 1.f;
    c = c / (mk3(1.f) - c + mk3(1e-4f));
    const float4 a = albedoMotion[p];
    store3(image, p, c * gmax(mk3(a.x, a.y, a.z), mk3(0.f)));
}
// add (denoiser.cu:230-248)
__global__ void __launch_bounds__(256) k_add(float* __restrict__ out, const float* __restrict__ a, const float* __restrict__ b, size_t n3) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n3) out[i] = a[i] + b[i];
}

}  // namespace

#ifndef RS_HOST_EMU      /* tests/emu compiles the kernels above with g++ and drives them itself; never defined in the product */
struct RstrDenoiser {
    RstrFrame* f = nullptr;
    int kind = 0;
    int W = 0, H = 0;
    float sigLumin = 0.f, sigNormal = 0.f, sigDepth = 0.f;
    // LeveledEAWFilter: devColorOut (the caller's in the reference) / devTempImg.  SpatioTemporalFilter: devColorOut, devTempColor,
    // devAccumColor[2], devAccumMoment[2], devVariance, devTempVariance, devFilteredVariance (denoiser.h:33-66)
    float* colorOut = nullptr; float* tempColor = nullptr;
    float* accumColor[2] = {nullptr, nullptr}; float* accumMoment[2] = {nullptr, nullptr};
    float* variance = nullptr; float* tempVariance = nullptr; float* filteredVariance = nullptr;
    bool firstTime = true;
    int frameIdx = 0;
};

static void denoiserFree(RstrDenoiser* d) {
    float* all[] = {d->colorOut, d->tempColor, d->accumColor[0], d->accumColor[1], d->accumMoment[0], d->accumMoment[1], d->variance, d->tempVariance, d->filteredVariance};
    for (float* p : all) cudaFree(p);
    delete d;
}
static DenoiseG denoiseG(const RstrFrame* f) {
    DenoiseG g;
    g.W = f->W; g.H = f->H;
    g.geom = f->geom[f->cur]; g.geomLast = f->geom[f->cur ^ 1];
    g.id = f->matId[f->cur]; g.idLast = f->matId[f->cur ^ 1];
    g.albedoMotion = f->albedoMotion;
    return g;
}

extern "C" {

int rstr_denoiser_create(RstrFrame* f, int kind, RstrDenoiser** out) {
    if (!f || !out || (kind != RSTR_DENOISER_EAW && kind != RSTR_DENOISER_SVGF)) return rsFail(RSTR_ERR_ARG, "rstr_denoiser_create: bad argument");
    if (f->row0 != 0 || f->row1 != f->H || f->bufRow0 != 0) return rsFail(RSTR_ERR_ARG, "rstr_denoiser_create: the filters work on full frames, not on strips");
    RstrDenoiser* d = new RstrDenoiser;
    d->f = f; d->kind = kind; d->W = f->W; d->H = f->H;
    const size_t n = (size_t)f->W * f->H;
    cudaError_t e = cudaMalloc(&d->colorOut, n * 12);
    if (e == cudaSuccess) e = cudaMalloc(&d->tempColor, n * 12);
    if (kind == RSTR_DENOISER_EAW) {
        d->sigLumin = 64.f; d->sigNormal = .2f; d->sigDepth = 1.f;                   // LeveledEAWFilter::create, denoiser.cu:455
    } else {
        d->sigLumin = 4.f; d->sigNormal = 128.f; d->sigDepth = 1.f;                  // SpatioTemporalFilter::create, denoiser.cu:488
        for (int i = 0; i < 2 && e == cudaSuccess; i++) {
            e = cudaMalloc(&d->accumColor[i], n * 12);
            if (e == cudaSuccess) e = cudaMalloc(&d->accumMoment[i], n * 12);
            // the reference leaves these uninitialised; the first frame never reads them (temporalAccumulate's `first`)
            if (e == cudaSuccess) e = cudaMemset(d->accumColor[i], 0, n * 12);
            if (e == cudaSuccess) e = cudaMemset(d->accumMoment[i], 0, n * 12);
        }
        if (e == cudaSuccess) e = cudaMalloc(&d->variance, n * 4);
        if (e == cudaSuccess) e = cudaMalloc(&d->tempVariance, n * 4);
        if (e == cudaSuccess) e = cudaMalloc(&d->filteredVariance, n * 4);
    }
    if (e != cudaSuccess) { denoiserFree(d); return rsFail(RSTR_ERR_CUDA, std::string("rstr_denoiser_create: ") + cudaGetErrorString(e)); }
    *out = d;
    return RSTR_OK;
}

int rstr_denoiser_destroy(RstrDenoiser* d) {
    if (!d) return RSTR_OK;
    cudaStreamSynchronize(d->f->stream);
    denoiserFree(d);
    return RSTR_OK;
}

int rstr_denoiser_set_sigmas(RstrDenoiser* d, float sigLumin, float sigNormal, float sigDepth) {
    if (!d) return rsFail(RSTR_ERR_ARG, "rstr_denoiser_set_sigmas: null denoiser");
    d->sigLumin = sigLumin; d->sigNormal = sigNormal; d->sigDepth = sigDepth;
    return RSTR_OK;
}

int rstr_denoiser_filter(RstrDenoiser* d, const RstrCamera* cam) {
    if (!d || !cam) return rsFail(RSTR_ERR_ARG, "rstr_denoiser_filter: bad argument");
    RstrFrame* f = d->f;
    { int rc = rsFlushGBuffer(f); if (rc) return rc; }
    const DenoiseG g = denoiseG(f);
    const CamDev c = rsToCamDev(*cam);
    const size_t n = (size_t)d->W * d->H;
    const dim3 grid((d->W + 15) / 16, (d->H + 7) / 8);
    const unsigned lin = (unsigned)((n + 255) / 256);
    cudaStream_t st = f->stream;
    const float* colorIn = f->radiance;
    int launches = 0;
    if (d->kind == RSTR_DENOISER_EAW) {                                              // LeveledEAWFilter::filter, denoiser.cu:463-477
        k_eaw<<<grid, 128, 0, st>>>(g, c, d->colorOut, colorIn, d->sigDepth, d->sigNormal, d->sigLumin, 0);
        for (int level = 1; level <= 4; level++) {
            k_eaw<<<grid, 128, 0, st>>>(g, c, d->tempColor, d->colorOut, d->sigDepth, d->sigNormal, d->sigLumin, level);
            std::swap(d->colorOut, d->tempColor);
        }
        launches = 5;
    } else {                                                                         // SpatioTemporalFilter::filter, denoiser.cu:537-564
        const int fi = d->frameIdx;
        k_temporal_accumulate<<<lin, 256, 0, st>>>(g, d->accumColor[fi], d->accumColor[fi ^ 1], d->accumMoment[fi], d->accumMoment[fi ^ 1], colorIn, d->firstTime ? 1 : 0);
        d->firstTime = false;
        k_estimate_variance<<<lin, 256, 0, st>>>(d->variance, d->accumMoment[fi], d->W, d->H);
        auto wavelet = [&](float* out, const float* in, int level) {
            k_filter_variance<<<lin, 256, 0, st>>>(d->filteredVariance, d->variance, d->W, d->H);
            k_eaw_svgf<<<grid, 128, 0, st>>>(g, c, out, in, d->tempVariance, d->variance, d->filteredVariance, d->sigDepth, d->sigNormal, d->sigLumin, level);
        };
        wavelet(d->colorOut, d->accumColor[fi], 0);
        std::swap(d->colorOut, d->accumColor[fi]);          // the level-0 result becomes the history the next frame accumulates into
        std::swap(d->tempVariance, d->variance);
        wavelet(d->colorOut, d->accumColor[fi], 1);
        std::swap(d->tempVariance, d->variance);
        for (int level = 2; level <= 4; level++) {
            wavelet(d->tempColor, d->colorOut, level);
            std::swap(d->tempColor, d->colorOut);
            std::swap(d->tempVariance, d->variance);
        }
        launches = 12;
    }
    rsCountLaunches(launches);
    CU(cudaGetLastError());
    return RSTR_OK;
}

int rstr_denoiser_next_frame(RstrDenoiser* d) {                                      // SpatioTemporalFilter::nextFrame, denoiser.cu:566-568
    if (!d) return rsFail(RSTR_ERR_ARG, "rstr_denoiser_next_frame: null denoiser");
    d->frameIdx ^= 1;
    return RSTR_OK;
}

int rstr_denoiser_read(RstrDenoiser* d, float* rgb, float* variance) {
    if (!d || !rgb) return rsFail(RSTR_ERR_ARG, "rstr_denoiser_read: bad argument");
    if (variance && d->kind != RSTR_DENOISER_SVGF) return rsFail(RSTR_ERR_ARG, "rstr_denoiser_read: only the SVGF filter carries a variance");
    const size_t n = (size_t)d->W * d->H;
    CU(cudaStreamSynchronize(d->f->stream));
    CU(cudaMemcpy(rgb, d->colorOut, n * 12, cudaMemcpyDeviceToHost));
    if (variance) CU(cudaMemcpy(variance, d->variance, n * 4, cudaMemcpyDeviceToHost));
    return RSTR_OK;
}

// modulateAlbedo(devImage, gBuffer) on the filtered image (denoiser.cu:405-411)
int rstr_denoiser_modulate_albedo(RstrDenoiser* d) {
    if (!d) return rsFail(RSTR_ERR_ARG, "rstr_denoiser_modulate_albedo: null denoiser");
    const size_t n = (size_t)d->W * d->H;
    k_modulate<<<(unsigned)((n + 255) / 256), 256, 0, d->f->stream>>>(d->colorOut, d->f->albedoMotion, n);
    rsCountLaunches(1);
    CU(cudaGetLastError());
    return RSTR_OK;
}

// addImage(devOut, devIn1, devIn2) (denoiser.cu:420-425) over two filters' images: a += b
int rstr_denoiser_add_image(RstrDenoiser* a, const RstrDenoiser* b) {
    if (!a || !b || a->W != b->W || a->H != b->H || a->f->stream != b->f->stream) return rsFail(RSTR_ERR_ARG, "rstr_denoiser_add_image: bad argument");
    const size_t n3 = (size_t)a->W * a->H * 3;
    k_add<<<(unsigned)((n3 + 255) / 256), 256, 0, a->f->stream>>>(a->colorOut, a->colorOut, b->colorOut, n3);
    rsCountLaunches(1);
    CU(cudaGetLastError());
    return RSTR_OK;
}

}  // extern "C"
#endif  // RS_HOST_EMU
