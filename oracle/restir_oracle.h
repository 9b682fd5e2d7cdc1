/*
 * restir_oracle.h -- C API of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT).
 *
 * The oracle is a from-scratch CPU restatement of the HummaWhite/ReSTIR direct-
 * illumination hot path (see restir_oracle.cpp for per-function reference
 * file:line citations).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product library
 * (librestir_b200.so) never links or calls anything in oracle/.
 *
 * The very same API is exported by oracle/_ref/libref_harness.so, which is the
 * reference's OWN headers (scene.h, bvh.cpp, sampler.h, material.h, restir.h ...)
 * compiled with g++ from /root/reference by oracle/Makefile; that build pins
 * this restatement (tests/test_oracle_vs_ref.py, tests/golden/).
 */
#ifndef RESTIR_ORACLE_H
#define RESTIR_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference Camera POD, 196 bytes (sceneStructs.h:22-126) */
typedef struct OrcCamera {
    int   resolution[2];      /*   0 */
    float position[3];        /*   8 */
    float rotation[3];        /*  20  yaw, pitch, roll in degrees */
    float view[3];            /*  32 */
    float up[3];              /*  44 */
    float right[3];           /*  56 */
    float fov[2];             /*  68  degrees; y = HALF vertical angle */
    float pixelLength[2];     /*  76 */
    float rotationMatInv[9];  /*  84  column-major mat3 */
    float viewProjection[16]; /* 120  unused by the hot path */
    float lensRadius;         /* 184 */
    float focalDist;          /* 188 */
    float tanFovY;            /* 192 */
} OrcCamera;

/* reference Material POD, 44 bytes (material.h:258-267) */
typedef struct OrcMaterial {
    int   type;               /* 0 Lambertian 1 MetallicWorkflow 2 Dielectric 3 Disney 4 Light */
    float baseColor[3];
    float metallic, roughness, ior;
    int   baseColorMapId, metallicMapId, roughnessMapId, normalMapId; /* -1 none */
} OrcMaterial;

/* knobs that are literals in restir.cu; defaults reproduce the reference */
typedef struct OrcParams {
    int   numCandidates;   /* restir.cu:3    32 */
    int   temporalCap;     /* restir.cu:183  20 */
    int   numSpatial;      /* restir.cu:93    5 */
    float spatialRadius;   /* restir.cu:49    5 */
    int   reuse;           /* common.h:36-43 bit0 temporal, bit1 spatial */
    int   spatialPasses;   /* restir.cu:196-209: 1 = as shipped; 2..3 = the commented-out extra pass(es), preClampedMerge<4> */
    int   unbiased;        /* 0 = the reference's reuse.  1 = restatement of the product's ADDITIONAL unbiased mode (not in the reference):
                              light point + contribution weight W in the reservoir, target re-evaluated at the receiver, 1/Z, one
                              visibility test for the final sample (include/restir_b200.h, RstrParams::unbiased).  Only the port
                              implements it; the reference harness reads the fields above only. */
} OrcParams;

/* buffer selectors for orc_frame_buffer */
enum {
    ORC_BUF_ALBEDO = 0,     /* P x 3 f32 */
    ORC_BUF_NORMAL = 1,     /* P x 3 f32, current frame */
    ORC_BUF_MATID = 2,      /* P i32 ("primId" buffer), current frame */
    ORC_BUF_DEPTH = 3,      /* P f32 */
    ORC_BUF_MOTION = 4,     /* P i32 */
    ORC_BUF_RADIANCE = 5,   /* P x 3 f32 (directIllum) */
    ORC_BUF_RESERVOIR = 6,  /* P x 36 B: history written by the last restir_direct */
    ORC_BUF_RESERVOIR_TEMP = 7, /* P x 36 B */
    ORC_BUF_LIGHT_INDEX = 8 /* P i32: light id of the sample held by the history reservoir (-1 none) */
};

typedef struct OrcScene OrcScene;
typedef struct OrcFrame OrcFrame;

/* Scene from flattened world-space triangle soup (what Scene::buildDevData produces):
 * builds light list, alias table, BVH + 6 MTBVH orderings. */
OrcScene* orc_scene_create(int numTris, const float* vertices, const float* normals,
                           const float* texcoords, const int* materialIds,
                           int numMaterials, const OrcMaterial* materials);
/* Textures (linear RGB float, width x height x 3, row-major; image.h:7-39) referenced by the materials' map ids, and
 * the environment map (texture index or -1).  Call once, right after orc_scene_create: an environment map becomes the
 * last entry of the light sampler (scene.cpp:136-157).  Returns 0, or -1 for an out-of-range texture id. */
int  orc_scene_set_textures(OrcScene*, int numTextures, const int* widths, const int* heights, const float* const* rgb, int envMapTexId);
const void* orc_scene_env_alias(const OrcScene*, int* lengthOut, float* sumAllOut);   /* envMapSampler table, W*H x {f32, i32} */
void orc_scene_destroy(OrcScene*);
int  orc_scene_bvh_size(const OrcScene*);
const float* orc_scene_boxes(const OrcScene*);          /* bvhSize x 6 f32 */
const int*   orc_scene_mtbvh(const OrcScene*, int i);   /* bvhSize x 3 i32 {prim, box, miss} */
int  orc_scene_num_lights(const OrcScene*);
const int*   orc_scene_light_prim_ids(const OrcScene*);
const float* orc_scene_light_radiance(const OrcScene*); /* L x 3 */
const void*  orc_scene_alias_table(const OrcScene*);    /* L x {f32 prob, i32 failId} */
float orc_scene_sum_light_power(const OrcScene*);
int  orc_scene_bvh_depth(const OrcScene*);

void orc_camera_update(OrcCamera*);                     /* Camera::update(), sceneStructs.h:88 */

OrcFrame* orc_frame_create(const OrcScene*, int w, int h);
void orc_frame_destroy(OrcFrame*);
void orc_frame_reset(OrcFrame*);                        /* ReSTIRReset, restir.cu:516 */
void orc_gbuffer_render(OrcFrame*, const OrcCamera*);   /* GBuffer::render, gbuffer.cu:80 */
void orc_gbuffer_update(OrcFrame*, const OrcCamera*);   /* GBuffer::update, gbuffer.cu:75 */
void orc_restir_direct(OrcFrame*, const OrcCamera*, const OrcParams*, int looper, int iter);
void orc_pathtrace_direct(OrcFrame*, const OrcCamera*, int looper, int iter);
const void* orc_frame_buffer(OrcFrame*, int which);
void orc_set_threads(int n);                             /* 0 = all */
int  orc_get_threads(void);
/* traversal statistics of the last gbuffer_render (node visits, triangle tests) */
void orc_last_trace_stats(OrcFrame*, uint64_t* nodesVisited, uint64_t* trisTested, uint64_t* rays);

/* RNG known-answer helper: n draws of sample1D for (looper, pixel index) */
void orc_rng_draws(int looper, int index, int n, float* out);

/* alias table of arbitrary weights (sampler.h:79-121); returns sumAll */
float orc_alias_build(int n, const float* values, void* outTable);

/* closest-hit / any-hit probes for unit tests */
int  orc_intersect(const OrcScene*, const float* origin, const float* dir, float* outPosNormUv8, int* outMatId);
int  orc_occluded(const OrcScene*, const float* x, const float* y);

/* ReSTIR GI (ReSTIRIndirect / ReSTIRIndirectKernel, restir.cu:232-476): one call = one frame of indirect illumination for the frame's
 * current G-buffer (call between orc_gbuffer_render and orc_gbuffer_update).  maxDepth = Settings::traceDepth, reuse bit 0 = temporal. */
typedef struct OrcGI OrcGI;
OrcGI* orc_gi_create(OrcFrame*);
void orc_gi_destroy(OrcGI*);
void orc_restir_indirect(OrcGI*, const OrcCamera*, int looper, int iter, int maxDepth, int reuse);
const float* orc_gi_indirect(OrcGI*);     /* P x 3 f32 (devIndirectIllum) */
const float* orc_gi_reservoirs(OrcGI*);   /* P x 17 f32: Lo xv nv xs ns, M, weight of the reservoirs the last call wrote */

/* The reference's image-space filters (denoiser.cu:25-567): kind 1 = LeveledEAWFilter, 2 = SpatioTemporalFilter, applied to the
 * frame's radiance with its current G-buffer.  Port only (restated): see restir_oracle.cpp. */
typedef struct OrcDenoiser OrcDenoiser;
OrcDenoiser* orc_denoiser_create(OrcFrame*, int kind);
void orc_denoiser_destroy(OrcDenoiser*);
void orc_denoiser_set_sigmas(OrcDenoiser*, float sigLumin, float sigNormal, float sigDepth);
void orc_denoiser_filter(OrcDenoiser*, const OrcCamera*);
void orc_denoiser_next_frame(OrcDenoiser*);
void orc_denoiser_modulate_albedo(OrcDenoiser*);
const float* orc_denoiser_color(OrcDenoiser*);      /* P x 3 f32 */
const float* orc_denoiser_variance(OrcDenoiser*);   /* P f32 (kind 2) */

#ifdef __cplusplus
}
#endif
#endif
