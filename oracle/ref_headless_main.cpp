/*
 * ref_headless_main.cpp -- headless driver of THE REFERENCE'S OWN CUDA BUILD.  TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * Linked by `make -C oracle ref_cuda` with the reference's gbuffer.cu, denoiser.cu (GBuffer::create/destroy), restir.cu
 * (compiled from a mktemp copy carrying the mechanical patch of SURVEY.md App. D: scope braces around
 * restir.cu:140-226 for nvcc's "goto bypasses initialisation", the unused GI kernel #if 0'd, the four literals made
 * -D overridable) and its host .cpp files, all compiled where they lie under /root/reference.  Replays
 * runCuda() (main.cpp:146-185) without GL: fixed animation clock t_k = k * animateSpeed / 60.
 *
 *   ref_headless <scene.txt> <frames> <warmup> <reuse 0..3> [dump_prefix dump_frame]
 * With REF_DENOISE=1 in the environment the reference's own filters (denoiser.cu, linked unmodified) run on every frame's
 * devDirectIllum as well -- SpatioTemporalFilter::filter + nextFrame each frame, LeveledEAWFilter::filter on the dump frame --
 * outside the timed region, and their images are dumped too (eaw.bin, svgf.bin, svgf_var.bin).
 * prints one JSON line: mean ms/frame of GBuffer::render + ReSTIRDirect (CUDA events; the reference's device-wide
 * sync after every launch, cudaUtil.h:15, is left in as shipped).
 */
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "scene.h"
#include "gbuffer.h"
#include "restir.h"
#include "denoiser.h"

extern Reservoir<DirectLiSample>* devLastDirectReservoir;   /* restir.cu:9, made extern by build_ref_cuda.sh */

static void dump(const std::string& path, const void* dev, size_t bytes) {
    std::vector<char> h(bytes);
    cudaMemcpy(h.data(), dev, bytes, cudaMemcpyDeviceToHost);
    FILE* f = fopen(path.c_str(), "wb");
    fwrite(h.data(), 1, bytes, f);
    fclose(f);
}

#ifdef __CUDACC__
/* probe: the reference's device functions evaluated on the GPU for one fixed input (debugging aid, REF_PROBE=1) */
__global__ void probeKernel(DevScene* scene, float* out) {
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
#endif

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: %s scene.txt frames warmup reuse [dump_prefix dump_frame]\n", argv[0]); return 2; }
    int frames = atoi(argv[2]), warmup = atoi(argv[3]);
    Settings::reservoirReuse = atoi(argv[4]);
    std::string dumpPrefix = argc > 6 ? argv[5] : "";
    int dumpFrame = argc > 6 ? atoi(argv[6]) : -1;
    FILE* out = stdout;
    stdout = stderr;                        /* the reference logs with std::cout; keep the JSON line alone */
    Scene* scene = new Scene(argv[1]);
    scene->buildDevData();
    State::scene = scene;
    Camera& cam = scene->camera;
    const int w = cam.resolution.x, h = cam.resolution.y;
    glm::vec3* devDirectIllum = cudaMalloc<glm::vec3>(w * h);
    cudaMemset(devDirectIllum, 0, sizeof(glm::vec3) * w * h);
    if (getenv("REF_PROBE")) {
        float* d; cudaMalloc(&d, 64);
        probeKernel<<<1, 1>>>(scene->devScene, d);
        float hst[16]; cudaMemcpy(hst, d, 64, cudaMemcpyDeviceToHost);
        fprintf(out, "probe p=%g Li.x=%g wi.y=%g dist=%g sumInv=%g L=%g w=%g g.x=%g bsdf=%g cos=%g sizeofDevScene=%g type=%g\n", hst[0], hst[1], hst[2], hst[3], hst[4], hst[5], hst[6], hst[7], hst[8], hst[9], hst[10], hst[11]);
    }
    GBuffer gBuffer;
    gBuffer.create(w, h);
    ReSTIRInit();
    const bool denoise = getenv("REF_DENOISE") != nullptr;
    LeveledEAWFilter eawFilter;                                             /* main.cpp:33-34, 78-79 */
    SpatioTemporalFilter svgfFilter;
    glm::vec3 *devEawOut = nullptr, *devSvgfOut = nullptr;
    if (denoise) {
        eawFilter.create(w, h, 5);
        svgfFilter.create(w, h, 5);
        devEawOut = cudaMalloc<glm::vec3>(w * h);
        devSvgfOut = cudaMalloc<glm::vec3>(w * h);
        /* the reference leaves its history buffers uninitialised; the first frame does not use them, but reads them */
        for (int i = 0; i < 2; i++) { cudaMemset(svgfFilter.devAccumColor[i], 0, sizeof(glm::vec3) * w * h); cudaMemset(svgfFilter.devAccumMoment[i], 0, sizeof(glm::vec3) * w * h); }
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const glm::vec3 camOrigPos = cam.position;
    double total = 0.0; int timed = 0;
    for (int k = 0; k < warmup + frames; k++) {
        float t = float(k) / 60.f * Settings::animateSpeed;                 /* main.cpp:150 with a fixed clock */
        cam.position = camOrigPos + glm::vec3(glm::cos(t), 0.f, glm::sin(t)) * Settings::animateRadius;
        cam.update();
        if (k == 0) gBuffer.lastCamera = cam;                               /* uninitialised in the reference (gbuffer.h:56) */
        cudaEventRecord(e0);
        gBuffer.render(scene->devScene, cam);
        ReSTIRDirect(devDirectIllum, 0, gBuffer);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (k >= warmup) { total += ms; timed++; }
        if (denoise) svgfFilter.filter(devSvgfOut, devDirectIllum, gBuffer, cam);
        if (denoise && k == dumpFrame) {
            eawFilter.filter(devEawOut, devDirectIllum, gBuffer, cam);
            cudaDeviceSynchronize();
            dump(dumpPrefix + "eaw.bin", devEawOut, sizeof(glm::vec3) * w * h);
            dump(dumpPrefix + "svgf.bin", devSvgfOut, sizeof(glm::vec3) * w * h);
            dump(dumpPrefix + "svgf_var.bin", svgfFilter.devVariance, sizeof(float) * w * h);
        }
        if (denoise) svgfFilter.nextFrame();
        if (k == dumpFrame) {
            dump(dumpPrefix + "radiance.bin", devDirectIllum, sizeof(glm::vec3) * w * h);
            dump(dumpPrefix + "matid.bin", gBuffer.devPrimId[gBuffer.frameIdx], sizeof(int) * w * h);
            dump(dumpPrefix + "depth.bin", gBuffer.devDepth[gBuffer.frameIdx], sizeof(float) * w * h);
            dump(dumpPrefix + "normal.bin", gBuffer.devNormal[gBuffer.frameIdx], sizeof(glm::vec3) * w * h);
            dump(dumpPrefix + "motion.bin", gBuffer.devMotion, sizeof(int) * w * h);
            dump(dumpPrefix + "reservoir.bin", devLastDirectReservoir, 36 * (size_t)w * h);   /* after the swap at restir.cu:434 */
        }
        gBuffer.update(cam);
        cam.position = camOrigPos;
    }
    fprintf(out, "{\"impl\": \"reference_cuda\", \"ms_per_frame\": %.6f, \"frames\": %d, \"warmup\": %d, \"width\": %d, \"height\": %d, \"reuse\": %d}\n",
            total / (timed ? timed : 1), timed, warmup, w, h, Settings::reservoirReuse);
    fflush(out);
    return 0;
}
