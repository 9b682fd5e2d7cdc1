/*
 * restir_b200.h -- C ABI of the B200-native ReSTIR direct-illumination pipeline.
 *
 * Drop-in boundary for the frame hot path of HummaWhite/ReSTIR.  Each entry point names the
 * reference interface it replaces (paths relative to the reference's src/).  The reference has no
 * plugin registry: the path sits behind C++ free functions / methods over global state
 * (State::scene, State::looper, Settings::reservoirReuse, file-static reservoir pointers); here that
 * state is explicit in opaque handles and every call returns 0 on success or a negative code, with
 * the message available from rstr_last_error().  Nothing in this library exit()s (the reference's
 * checkCUDAError does, cudaUtil.h:13-31) and nothing falls back to the CPU: without a usable CUDA
 * device every compute call fails with RSTR_ERR_CUDA.
 *
 * Host-facing layouts kept from the reference (all little-endian, fp32 / int32):
 *   Camera 196 B (sceneStructs.h:22-126), Material 44 B (material.h:258-267),
 *   reservoir 36 B AoS {Li, wi, dist, numSamples, weight} (restir.h:7-11,114-116),
 *   G-buffer arrays albedo vec3 / normal vec3 / "primId" = material id / depth / motion (gbuffer.h:15-58),
 *   boundingBoxes 24 B and MTBVHNode 12 B x 6 orderings (bvh.h:15-171), alias entries 8 B (sampler.h:63-67).
 * Device-internal layouts are different; rstr_*_read converts.
 */
#ifndef RESTIR_B200_H
#define RESTIR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSTR_OK 0
#define RSTR_ERR_ARG (-1)    /* bad argument / unsupported feature */
#define RSTR_ERR_CUDA (-2)   /* CUDA runtime error (message in rstr_last_error) */
#define RSTR_ERR_IO (-3)     /* scene file / OBJ could not be read or parsed */
#define RSTR_ERR_LIMIT (-4)  /* BVH deeper than the traversal stack, etc. */

/* Camera POD -- sceneStructs.h:22-126 (offsets: SURVEY.md App. E) */
typedef struct RstrCamera {
    int   resolution[2];      /*   0 */
    float position[3];        /*   8 */
    float rotation[3];        /*  20  yaw, pitch, roll (degrees) */
    float view[3];            /*  32 */
    float up[3];              /*  44 */
    float right[3];           /*  56 */
    float fov[2];             /*  68  degrees; y is the HALF vertical angle (sceneStructs.h:72) */
    float pixelLength[2];     /*  76 */
    float rotationMatInv[9];  /*  84  column-major */
    float viewProjection[16]; /* 120  not used by the DI path */
    float lensRadius;         /* 184 */
    float focalDist;          /* 188 */
    float tanFovY;            /* 192 */
} RstrCamera;

/* Material POD -- material.h:258-267 */
typedef struct RstrMaterial {
    int   type;               /* 0 Lambertian, 1 MetallicWorkflow, 2 Dielectric, 3 Disney, 4 Light (material.h:114-120) */
    float baseColor[3];       /* for Light: emitted radiance */
    float metallic, roughness, ior;
    int   baseColorMapId, metallicMapId, roughnessMapId, normalMapId;  /* index into RstrSceneDesc::textures; -1 none
                                                                          (material.h:11); baseColorMapId -2 = procedural
                                                                          pattern (material.h:13, scene.h:68-76) */
} RstrMaterial;

/* Texture in the reference's host form (image.h:7-39): width x height texels, 3 x f32 linear RGB, row-major */
typedef struct RstrTexture {
    int width, height;
    const float* rgb;
} RstrTexture;

/* Knobs that are literals in restir.cu; rstr_params_default() reproduces the reference. */
typedef struct RstrParams {
    int   numCandidates;   /* restir.cu:3    #define ReservoirSize 32 */
    int   temporalCap;     /* restir.cu:183  preClampedMerge<20> */
    int   numSpatial;      /* restir.cu:93   5 neighbours */
    float spatialRadius;   /* restir.cu:49   5 px */
    int   reuse;           /* Settings::reservoirReuse (common.h:36-43): bit0 temporal, bit1 spatial */
    int   spatialPasses;   /* 1 = as shipped (restir.cu:196-199); 2..3 add the commented-out pass restir.cu:201-209
                              (publish, barrier, new 5-neighbour aggregate, preClampedMerge<4>); values < 1 mean 1 */
    int   unbiased;        /* 0 = the reference's reuse (a reused sample keeps the direction / distance it had at the pixel that
                              drew it, neighbours merge with their weight sums, 1/M normalisation, visibility tested for the RIS
                              winner only: restir.cu:172-199, restir.h:61-70 -- biased).  1 = ADDITIONAL mode the reference does not
                              have (north star: "unbiased-MIS / visibility re-check"): reservoirs carry the light POINT and its
                              unbiased contribution weight W; every reused sample's target function is re-evaluated at the receiving
                              pixel; spatial merges are normalised by 1/Z, Z = sum of M over the merged pixels whose own target is
                              non-zero for the chosen sample; visibility is tested for the final sample at the receiving pixel.
                              The accumulated image converges to pathTraceDirect's.  Triangle lights only; single-GPU frames. */
} RstrParams;

#define RSTR_REUSE_NONE 0
#define RSTR_REUSE_TEMPORAL 1
#define RSTR_REUSE_SPATIAL 2
#define RSTR_REUSE_SPATIOTEMPORAL 3

/* Flattened world-space scene, i.e. the arrays Scene::buildDevData() produces (scene.cpp:159-190). */
typedef struct RstrSceneDesc {
    int numTris;
    const float* vertices;     /* 3T x 3 */
    const float* normals;      /* 3T x 3 */
    const float* texcoords;    /* 3T x 2, may be NULL */
    const int*   materialIds;  /* T */
    int numMaterials;
    const RstrMaterial* materials;
    int numTextures;             /* Scene::textures (scene.cpp:362-375); 0 = none */
    const RstrTexture* textures;
    int envMap;                  /* 0 = none, else 1 + index of the environment map texture (Scene::envMapTexId,
                                    scene.cpp:122-128): sampled as the last entry of the light sampler (scene.cpp:136-152,
                                    scene.h:364-375, 400-403) and shown behind the scene (gbuffer.cu:59-62, restir.cu:134) */
} RstrSceneDesc;

typedef struct RstrSceneInfo {
    int numTris, numLights, bvhSize, bvhDepth, numMaterials;
    float sumLightPower;
    double buildSeconds;       /* host: reference-identical BVH + tables + traced BVH */
    size_t deviceBytes;
    int tracedBvhDepth;        /* depth of the binned-SAH tree the kernels trace */
    double tracedBuildSeconds;
    int numEmissiveTris;       /* numLights counts the light sampler's entries: emissive triangles + 1 for an environment map */
    int numTextures, envWidth, envHeight;
    int tracedNodes, tracedRoot;  /* node count / root reference of the traced tree (RSTR_SCENE_TRACED_NODES) */
} RstrSceneInfo;

typedef struct RstrScene RstrScene;
typedef struct RstrFrame RstrFrame;

/* scene arrays readable through rstr_scene_read (reference host layouts) */
enum {
    RSTR_SCENE_BOXES = 0,          /* (2T-1) x 24 B           bvh.cpp:54   Scene::boundingBoxes */
    RSTR_SCENE_MTBVH0 = 1,         /* +i: (2T-1) x 12 B, i=0..5  bvh.cpp:133-201 Scene::BVHNodes[i] */
    RSTR_SCENE_LIGHT_PRIM_IDS = 7, /* L x i32                 scene.cpp:182 */
    RSTR_SCENE_LIGHT_RADIANCE = 8, /* L x 12 B                scene.cpp:183 */
    RSTR_SCENE_ALIAS = 9,          /* L x {f32 prob, i32 failId}  sampler.h:79-121 */
    RSTR_SCENE_VERTICES = 10,      /* 3T x 12 B */
    RSTR_SCENE_NORMALS = 11,
    RSTR_SCENE_TEXCOORDS = 12,     /* 3T x 8 B */
    RSTR_SCENE_MATERIAL_IDS = 13,  /* T x i32 */
    RSTR_SCENE_MATERIALS = 14,     /* numMaterials x 44 B */
    RSTR_SCENE_ENV_ALIAS = 15,     /* envW*envH x {f32 prob, i32 failId}  envMapSampler, scene.cpp:147 */
    RSTR_SCENE_TRACED_NODES = 16,  /* the binned-SAH tree the kernels walk: tracedNodes x 64 B {lmin[3],lmax[3],rmin[3],rmax[3],left,right,pad[2]};
                                      child >= 0 node, < 0 leaf = 0x80000000 | (count-1) << 27 | first triangle of TRACED_TRIS */
    RSTR_SCENE_TRACED_TRIS = 17,   /* T x 48 B {v0,v1,v2,matId,origPrim,pad} in leaf order */
    RSTR_SCENE_TEXTURE0 = 32       /* +i: texture i, width x height x 12 B (size from rstr_scene_texture_info)  Scene::textures */
};

/* frame buffers readable through rstr_frame_read (reference layouts, full image rows of this frame/strip) */
enum {
    RSTR_BUF_ALBEDO = 0,          /* P x 12 B  GBuffer::devAlbedo */
    RSTR_BUF_NORMAL = 1,          /* P x 12 B  GBuffer::normal() */
    RSTR_BUF_MATID = 2,           /* P x i32   GBuffer::primId()  (material id, -1 miss, -2 light; gbuffer.cu:29,42,65) */
    RSTR_BUF_DEPTH = 3,           /* P x f32   GBuffer::depth() */
    RSTR_BUF_MOTION = 4,          /* P x i32   GBuffer::devMotion */
    RSTR_BUF_RADIANCE = 5,        /* P x 12 B  devDirectIllum */
    RSTR_BUF_RESERVOIR = 6,       /* P x 36 B  history reservoirs written by the last rstr_restir_direct */
    RSTR_BUF_RESERVOIR_TEMP = 7,  /* P x 36 B  devDirectTemp */
    RSTR_BUF_LIGHT_INDEX = 8,     /* P x i32   light id held by each history reservoir (-1 = none) */
    RSTR_BUF_LDR = 9              /* P x uchar4  output of rstr_tonemap */
};

const char* rstr_last_error(void);
/* identifies the build: first 12 hex digits of the SHA-1 over the library's sources (restir_b200/build.py), so that a
 * measurement can be tied to the code that produced it */
const char* rstr_build_id(void);
/* binds the calling thread to a device (the reference pins device 0, preview.cpp:112) */
int rstr_init(int device);
void rstr_params_default(RstrParams*);

/* replaces Scene::buildDevData() + DevScene::create() (scene.cpp:159-215, 435-509): light list, alias
 * table, BVH (bit-identical node/primitive order to BVHBuilder::build, bvh.cpp:10-201), upload. */
int rstr_scene_create(const RstrSceneDesc*, RstrScene**);
/* replaces Scene::Scene(filename) (scene.cpp:96-131) + buildDevData; fills the scene file's camera */
int rstr_scene_load_file(const char* path, RstrScene**, RstrCamera* cameraOut);
/* replaces Scene::clear() + DevScene::destroy() (scene.cpp:217-220, 511-532) */
int rstr_scene_destroy(RstrScene*);
int rstr_scene_info(const RstrScene*, RstrSceneInfo*);
/* 0 (default): rays are traced through a binned-SAH tree; rays whose outcome depends on the reference's visiting order
 * (two hits within a few ulps of each other) or on its non-conservative near-axis box test (bvh.h:91-123) are
 * detected and re-traced with the reference-order walk.  1: every ray walks the reference tree in the reference's
 * order with its exact box predicate (validation mode, ~2x slower). */
int rstr_scene_set_traversal(RstrScene*, int mode);
/* SURVEY section 8 f3 (the role of BVHBuilder::build, bvh.cpp:10-131, for the tree that is TRACED): rebuilds the traced tree
 * of the scene on the device and swaps it in; every frame of the scene traces it from then on.  mode 0: PLOC (Morton order,
 * bottom-up merging of the nearest clusters by surface area, SAH leaf cut; ~ the host tree's quality); mode 1: the plain
 * Morton radix tree (fastest build, slower to trace).  Results do not change: which hit a ray reports is decided per
 * triangle; the reference-identical tree stays on the host for rstr_scene_read and for the
 * visiting-rank table.  Synchronises the device.  *milliseconds (may be NULL): device time of the build. */
int rstr_scene_build_traced_gpu(RstrScene*, int mode, float* milliseconds);
/* pixels recomputed with the reference-order walk since the scene was created / last reset (instrumentation);
 * count[4] = {G-buffer pixels, ReSTIR phase-A pixels, PTDirect pixels, unused} */
int rstr_scene_fallback_rays(RstrScene*, unsigned long long* count4, int reset);
/* size of texture i and whether it is the environment map (Scene::textures / envMapTexId, scene.h:487-491) */
int rstr_scene_texture_info(const RstrScene*, int index, int* width, int* height, int* isEnvMap);
int rstr_scene_read(const RstrScene*, int which, void* host, size_t bytes);

/* replaces Camera::update() (sceneStructs.h:88-102) */
int rstr_camera_update(RstrCamera*);
/* runCuda()'s camera animation (main.cpp:149-153 + update()): position = base + (cos t, 0, sin t) * radius with the wall
 * clock replaced by t = frame / fps * speed (Settings::animateSpeed 2.7, animateRadius 1, common.cpp:9-10) */
int rstr_camera_orbit(const RstrCamera* base, int frame, float speed, float radius, float fps, RstrCamera* out);

/* replaces GBuffer::create (denoiser.cu:373) + ReSTIRInit (restir.cu:478) + the devDirectIllum allocation
 * (main.cpp:38).  The strip variant owns image rows [row0,row1) of a W x H image plus `halo` rows on each
 * side (multi-GPU strips); pixel indices and RNG streams stay global. */
int rstr_frame_create(RstrScene*, int width, int height, RstrFrame**);
int rstr_frame_create_strip(RstrScene*, int width, int height, int row0, int row1, int halo, RstrFrame**);
/* replaces GBuffer::destroy (denoiser.cu:391) + ReSTIRFree (restir.cu:506) */
int rstr_frame_destroy(RstrFrame*);
/* replaces ReSTIRReset (restir.cu:516) */
int rstr_frame_reset(RstrFrame*);

/* replaces GBuffer::render(devScene, cam) (gbuffer.cu:80-85) */
int rstr_gbuffer_render(RstrFrame*, const RstrCamera*);
/* replaces GBuffer::update(cam) (gbuffer.cu:75-78) */
int rstr_gbuffer_update(RstrFrame*, const RstrCamera*);
/* replaces ReSTIRDirect(devDirectIllum, iter, gBuffer) (restir.cu:418-446); looper = State::looper */
int rstr_restir_direct(RstrFrame*, const RstrCamera*, const RstrParams*, int looper, int iter);
/* replaces pathTraceDirect(devDirectIllum, iter) (pathtrace.cu:457-473) */
int rstr_pathtrace_direct(RstrFrame*, const RstrCamera*, int looper, int iter);
/* replaces copyImageToPBO(devPBO, devImage, w, h, toneMapping, scale) (pathtrace.cu:108-113);
 * toneMapping: 0 none, 1 filmic, 2 ACES (common.h:18-22) */
int rstr_tonemap(RstrFrame*, int toneMapping, float scale);
/* saveImage (main.cpp:105-144): tone-map + gamma the radiance image, mirror it horizontally (main.cpp:126) and write an
 * 8-bit RGB PNG (Image::savePNG, image.cpp:41-57).  `path` is the complete file name. */
int rstr_frame_save_png(RstrFrame*, const char* path, int toneMapping);
int rstr_frame_save_jpg(RstrFrame*, const char* path, int toneMapping);   /* saveImage(jpg = true): Image::saveJPG, quality 90 (image.cpp:59-75) */
/* Image::Image(filename) (image.cpp:16-33): PNG, JPEG, BMP, TGA or Radiance .hdr -> width x height x 3 f32, linear (8-bit samples / 255,
 * stbi_ldr_to_hdr_gamma(1), scene.cpp:97); flipY = stbi_set_flip_vertically_on_load (true for material textures, false for
 * the environment map, scene.cpp:98,124).  Call with rgbOut = NULL to get the size.  Host only, no GPU needed. */
int rstr_image_load(const char* path, int flipY, int* width, int* height, float* rgbOut, size_t capacityBytes);
/* Image::savePNG (image.cpp:41-57): width x height x 3 bytes, top row first.  Host only. */
int rstr_image_write_png(const char* path, int width, int height, const unsigned char* rgb);
/* Image::saveJPG (image.cpp:59-75): the file stbi_write_jpg writes, byte for byte; quality 1..100, 0 = 90.  Host only. */
int rstr_image_write_jpg(const char* path, int width, int height, const unsigned char* rgb, int quality);

/* One whole frame of runCuda (main.cpp:146-185) from HOST inputs to a HOST result:
 * gbuffer_render + restir_direct (or pathtrace_direct when params == NULL) + tonemap + gbuffer_update,
 * then the LDR image (P x uchar4) is copied to hostLdr (pinned or pageable).  Synchronous. */
int rstr_render_frame_host(RstrFrame*, const RstrCamera*, const RstrParams*, int looper, int iter,
                           int toneMapping, void* hostLdr, size_t bytes);

#define RSTR_LDR_SLOTS 3
/* Pipelined form: enqueue frame k into LDR slot (0 .. RSTR_LDR_SLOTS-1) -- render, tone-map, D2H of the image on a copy stream -- and
 * return at once; rstr_frame_wait_host(slot) blocks until that slot's image has arrived in hostLdr.  Lets the copy of
 * frame k overlap the rendering of frame k+1 (the reference's GL interop path has no copy at all). */
int rstr_render_frame_host_async(RstrFrame*, const RstrCamera*, const RstrParams*, int looper, int iter,
                                 int toneMapping, void* hostLdr, size_t bytes, int slot);
int rstr_frame_wait_host(RstrFrame*, int slot);

int rstr_frame_sync(RstrFrame*);
int rstr_frame_read(RstrFrame*, int which, void* host, size_t bytes);
/* same conversion into DEVICE memory, stream-ordered and without host synchronisation (strip gather to GPU 0) */
int rstr_frame_read_device(RstrFrame*, int which, void* dev, size_t bytes);

/* ---- instrumentation / plumbing ---- */
/* device time (ms, CUDA events on the frame's stream) of the stages of the last frame */
enum { RSTR_T_GBUFFER = 0, RSTR_T_RIS = 1, RSTR_T_SPATIAL = 2, RSTR_T_PTDIRECT = 3, RSTR_T_TONEMAP = 4, RSTR_T_COUNT = 5 };
int rstr_frame_stage_ms(RstrFrame*, float* ms, int n);
/* timing marks: cudaEventRecord on the frame's stream into slot 0..7, and the elapsed device time between two */
int rstr_frame_mark(RstrFrame*, int slot);
int rstr_frame_elapsed_ms(RstrFrame*, int slotA, int slotB, float* ms);
/* make the frame enqueue its work on a caller-owned cudaStream_t (NULL = the legacy default stream) */
int rstr_frame_set_stream(RstrFrame*, void* stream);
/* number of kernel launches issued by this library so far (all frames of this process) */
uint64_t rstr_launch_count(void);
/* the frame's CUDA stream (cudaStream_t) for callers that enqueue their own work */
void* rstr_frame_stream(RstrFrame*);

/* number of temporal / spatial neighbour reads that fell outside the rows resident in a strip frame (must stay 0
 * for results identical to the single-GPU frame; widen `halo` otherwise) */
int rstr_frame_halo_miss(RstrFrame*, unsigned int* count);
/* zero that counter (stream-ordered): e.g. after warm-up frames that jumped between camera poses */
int rstr_frame_halo_miss_reset(RstrFrame*);
/* per-frame bound for the temporal halo: the largest |row(motion) - row| over the pixels whose reprojection
 * (gbuffer.cu:49-55) landed on screen, over every G-buffer rendered since the frame was created / the last reset.  A strip
 * frame needs halo > this value (and > ceil(spatialRadius)) for rstr_frame_halo_miss to stay 0. */
int rstr_frame_motion_rows(RstrFrame*, unsigned int* maxRows, int reset);
/* page-locked host memory for rstr_render_frame_host / rstr_frame_read targets */
void* rstr_host_alloc(size_t bytes);
void rstr_host_free(void*);

/* Halo planes of a strip frame, for the neighbour exchange.  `plane` selects one of
 * the device-internal per-pixel arrays; the returned pointer addresses row `row` (global image row) and
 * *rowBytes is the size of one image row of that plane. */
enum {
    RSTR_PLANE_GEOM_CUR = 0,   /* float4 {n.xyz, depth}, current frame */
    RSTR_PLANE_MATID_CUR = 1,  /* int */
    RSTR_PLANE_RESV_HISTORY = 2, /* 32 B reservoirs written by the last restir_direct */
    RSTR_PLANE_RESV_TEMP = 3,  /* 32 B post-temporal reservoirs (input of spatial pass 1; published by even passes) */
    RSTR_PLANE_RESV_TEMP2 = 4, /* 32 B reservoirs published by odd spatial passes (input of pass 2; spatialPasses > 1 only) */
    RSTR_PLANE_RESV_OUT = 5    /* the history reservoirs THIS frame's phase A wrote (restir.cu:211-212; phase B does not touch them):
                                  between phase A and phase B this is the plane RESV_HISTORY will name after phase B, so its halo
                                  rows can travel in the same exchange as RESV_TEMP */
};
int rstr_frame_plane_row(RstrFrame*, int plane, int row, void** devPtr, size_t* rowBytes);
/* copy rows [row0,row1) of `plane` from src to dst (same device or a peer device), ordered after the work queued on
 * src's stream and before work queued later on dst's stream: the single-process form of the halo exchange */
int rstr_frame_copy_rows(RstrFrame* dst, RstrFrame* src, int plane, int row0, int row1);
/* split form of rstr_restir_direct for strips: phase A (candidates + shadow + temporal), then the caller
 * exchanges RESV_TEMP / GEOM_CUR / MATID_CUR halo rows, then phase B (spatial + shade). */
int rstr_restir_phase_a(RstrFrame*, const RstrCamera*, const RstrParams*, int looper, int iter);
int rstr_restir_phase_b(RstrFrame*, const RstrCamera*, const RstrParams*, int looper, int iter);
/* One spatial pass (1-based) of phase B, for strip frames with spatialPasses > 1: pass p reads the reservoirs published
 * by pass p-1 (RESV_TEMP for p = 1, then RESV_TEMP2, RESV_TEMP, ...) in a neighbourhood, so the caller exchanges the
 * halo rows of that plane between two passes.  The last pass shades and closes the frame like rstr_restir_phase_b. */
int rstr_restir_phase_b_pass(RstrFrame*, const RstrCamera*, const RstrParams*, int looper, int iter, int pass);
/* Strip frames: by default rstr_gbuffer_render also renders the halo rows locally (no G-buffer traffic between
 * GPUs).  With renderHalo = 0 only the strip's own rows are rendered and the caller exchanges the GEOM_CUR / MATID_CUR
 * halo rows together with RESV_TEMP before phase B (cheaper when strips are thin: 20 B per halo pixel over NVLink
 * instead of a primary ray each). */
int rstr_frame_set_halo_render(RstrFrame*, int renderHalo);
/* By default rstr_gbuffer_render defers its launch: when the next call on the frame is rstr_restir_direct / _phase_a
 * with the same camera, ONE kernel renders the G-buffer and runs phase A, walking the tree once for the pixel's two
 * primary rays (centre ray, gbuffer.cu:11-23; jittered ray, restir.cu:129).  Results are identical to the two separate
 * kernels; any other use of the G-buffer in between (reads, plane pointers, gbuffer_update, sync, PTDirect) launches the
 * plain G-buffer kernel first.  enable = 0 turns the fusion off (A/B measurements, tests). */
int rstr_frame_set_fusion(RstrFrame*, int enable);
/* How the fused G-buffer + phase A work is launched: staged = 1 runs it as a pipeline of kernels cut where the shape of
 * the parallelism changes -- one packet walk per 8x4 tile for the primary rays, then the 32 light candidates, the shadow rays
 * (persistent warps refilled from a queue) and the temporal merge over the compacted list of shaded pixels; staged = 0 runs
 * everything in one kernel, one thread per pixel; staged = -1 (default) picks the pipeline unless the scene is tiny (a
 * traced tree of at most 1024 nodes, e.g. a Cornell box, where no ray diverges and the extra launches cost more than
 * they save).  Identical results in every mode (A/B measurements, tests). */
int rstr_frame_set_pipeline(RstrFrame*, int staged);
/* The staged pipeline can cut its rows into `bands` horizontal bands (1 .. 8), each with its own queue: a band's queue kernels run on
 * a side stream under the next band's k_primary.  Default 1 = strictly one kernel after the other: on B200 the overlap LOSES
 * 15 % (2.61 -> 3.00 ms on the 1M-triangle scene with 2 or 4 bands: the persistent shadow kernel and the packet walk fight for
 * the same SMs and L1), so it stays an A/B switch. */
int rstr_frame_set_bands(RstrFrame*, int bands);
/* Cost profile for placing the strip cuts.  enable != 0 starts (or restarts from zero) the
 * accumulation: the G-buffer and phase-A kernels add the SM cycles every block (16x8 pixels) held its SM slot to one
 * counter per group of 8 image rows.  If cyclesPerRowGroup != NULL the counters (numGroups = ceil(height / 8)) are read
 * first.  enable == 0 stops profiling and frees the counters. */
int rstr_frame_row_cost(RstrFrame*, int enable, double* cyclesPerRowGroup, int numGroups);

/* ---- multi-GPU strips: one process per GPU, halo rows and the final gather as peer stores over NVLink ----
 * Nothing in the reference to replace (it pins device 0, preview.cpp:112); this is the data plane BASELINE.json's north star
 * asks for.  Each rank creates a strip frame (rstr_frame_create_strip, rows [bounds[r], bounds[r+1]), halo >
 * max(ceil(spatialRadius), rstr_frame_motion_rows)), then a group on it; the ranks swap their RSTR_STRIP_HANDLE_BYTES blobs
 * once (any transport: MPI, a pipe, torch.distributed) and connect.  From then on rstr_strip_group_frame renders this rank's
 * strip of one frame -- identical, bit for bit, to the same rows of the single-GPU frame -- with the halo rows stored
 * directly into the neighbours' memory by a kernel on this rank's stream and awaited by flag on the neighbours' streams
 * (no host synchronisation, no library collective), and rstr_strip_group_present tone-maps the strip straight into rank
 * 0's full-frame LDR slot and brings the assembled frame to the host on rank 0.  Every rank must issue the same sequence
 * of group calls.  Ranks living in one process (tests) are connected through their device pointers instead of CUDA IPC. */
typedef struct RstrStripGroup RstrStripGroup;
#define RSTR_STRIP_HANDLE_BYTES 512
int rstr_strip_group_create(RstrFrame* stripFrame, int rank, int world, RstrStripGroup**);
int rstr_strip_group_handle(RstrStripGroup*, void* blob /* RSTR_STRIP_HANDLE_BYTES */);
int rstr_strip_group_connect(RstrStripGroup*, const void* blobsInRankOrder /* world x RSTR_STRIP_HANDLE_BYTES */);
/* gbuffer_render + restir_direct + gbuffer_update of runCuda (main.cpp:164-183) for this rank's strip, halo exchanges included */
int rstr_strip_group_frame(RstrStripGroup*, const RstrCamera*, const RstrParams*, int looper, int iter);
/* the building block of the above: push this rank's edge rows of the planes in planeMask (bit i = RSTR_PLANE_i) into the
 * neighbours' halo rows, then make the frame's stream wait for the neighbours' rows of the same exchange */
int rstr_strip_group_exchange(RstrStripGroup*, unsigned int planeMask);
/* the two halves of _exchange and of _frame.  Only needed when several ranks share ONE process (tests): all ranks' pushes must
 * be issued before any rank's wait, because the streams of one process share hardware work queues. */
int rstr_strip_group_push(RstrStripGroup*, unsigned int planeMask);
int rstr_strip_group_wait(RstrStripGroup*);
int rstr_strip_group_ack(RstrStripGroup*);     /* behind the kernels that read the halo rows: the neighbours may overwrite them */
int rstr_strip_group_frame_begin(RstrStripGroup*, const RstrCamera*, const RstrParams*, int looper, int iter);   /* G-buffer, phase A, push */
int rstr_strip_group_frame_end(RstrStripGroup*, const RstrCamera*, const RstrParams*, int looper, int iter);     /* wait, phase B (one pass), gbuffer_update */
/* copyImageToPBO for strips: tone-map into slot (0 .. RSTR_LDR_SLOTS-1) of rank 0's full frame; on rank 0 the assembled
 * W x H x uchar4 frame is copied to hostLdr (may be NULL) on a copy stream; rstr_strip_group_wait_host(slot) blocks until it is there */
int rstr_strip_group_present(RstrStripGroup*, int toneMapping, void* hostLdr, size_t bytes, int slot);
int rstr_strip_group_wait_host(RstrStripGroup*, int slot);
/* *flag = 1 when a wait on a peer timed out (about 10 s): the peer died or the ranks issued different call sequences */
int rstr_strip_group_error(RstrStripGroup*, int* flag);
int rstr_strip_group_destroy(RstrStripGroup*);

/* ---- the reference's image-space filters (denoiser.h:14-73, denoiser.cu:25-567).  The reference creates them
 * (main.cpp:78-80) and exposes their knobs in its GUI (preview.cpp:253-287) but never calls them from runCuda; here they are
 * applied to a full frame's radiance plane (devDirectIllum) with the frame's current G-buffer.
 *   RSTR_DENOISER_EAW   LeveledEAWFilter: five edge-avoiding a-trous wavelet passes (strides 1, 2, 4, 8, 16; sigmas 64 / .2 / 1)
 *   RSTR_DENOISER_SVGF  SpatioTemporalFilter: re-projected exponential history of colour and luminance moments, variance
 *                       estimate, five variance-guided a-trous passes (sigmas 4 / 128 / 1); the level-0 result feeds the history */
typedef struct RstrDenoiser RstrDenoiser;
enum { RSTR_DENOISER_EAW = 1, RSTR_DENOISER_SVGF = 2 };
int rstr_denoiser_create(RstrFrame* fullFrame, int kind, RstrDenoiser**);          /* LeveledEAWFilter::create / SpatioTemporalFilter::create */
int rstr_denoiser_destroy(RstrDenoiser*);                                          /* ...::destroy */
int rstr_denoiser_set_sigmas(RstrDenoiser*, float sigLumin, float sigNormal, float sigDepth);   /* EAWaveletFilter's members (the GUI sliders) */
/* ...::filter(devColorOut, devColorIn = the frame's radiance, gBuffer, cam): after rstr_restir_direct / rstr_pathtrace_direct and
 * before rstr_gbuffer_update of the same frame */
int rstr_denoiser_filter(RstrDenoiser*, const RstrCamera*);
int rstr_denoiser_next_frame(RstrDenoiser*);                                       /* SpatioTemporalFilter::nextFrame */
int rstr_denoiser_modulate_albedo(RstrDenoiser*);                                  /* modulateAlbedo(devColorOut, gBuffer) */
int rstr_denoiser_add_image(RstrDenoiser* a, const RstrDenoiser* b);               /* addImage(a.devColorOut, b.devColorOut) */
/* devColorOut as W x H x 3 floats (and, SVGF, the filtered variance as W x H floats; may be NULL) */
int rstr_denoiser_read(RstrDenoiser*, float* rgb, float* variance);

/* ---- ReSTIR GI (restir.h:13-27, 133; restir.cu:242-416, 448-476).  The reference ships ReSTIRIndirect commented out
 * at its call site (main.cpp:168) and with Settings::traceDepth = 0 (common.cpp:3); here it is a handle on a full frame:
 *   rstr_gi_create    the indirect share of ReSTIRInit (restir.cu:491-496: devIndTemporalReservoir, devIndLastTemporalReservoir, zeroed)
 *                     plus a devIndirectIllum plane of its own
 *   rstr_gi_destroy   ... of ReSTIRFree (restir.cu:511-512)
 *   rstr_gi_reset     ReSTIRReset (restir.cu:516): the next call does not read the history
 *   rstr_restir_indirect   ReSTIRIndirect(devIndirectIllum, iter, gBuffer) with State::looper, Settings::traceDepth and
 *                     Settings::reservoirReuse (bit 0 = temporal) as arguments: between rstr_gbuffer_render and rstr_gbuffer_update of
 *                     the frame.  target RSTR_GI_TARGET_OWN accumulates into the handle's own plane, RSTR_GI_TARGET_RADIANCE into the
 *                     frame's radiance plane (the reference's commented call passes devDirectIllum), where rstr_tonemap, the filters
 *                     and rstr_frame_save_* find it.
 * A pixel whose jittered ray leaves the scene or hits an emitter writes 0 (the reference shades a history sample with uninitialised
 * locals there).  */
typedef struct RstrGI RstrGI;
enum { RSTR_GI_TARGET_OWN = 0, RSTR_GI_TARGET_RADIANCE = 1 };
int rstr_gi_create(RstrFrame* fullFrame, RstrGI**);
int rstr_gi_destroy(RstrGI*);
int rstr_gi_reset(RstrGI*);
int rstr_restir_indirect(RstrGI*, const RstrCamera*, int looper, int iter, int traceDepth, int reuse, int target);
/* the own plane as W x H x 3 floats (may be NULL) and the reservoirs the last call wrote in the reference's layout,
 * Reservoir<IndirectLiSample> = 68 bytes: Lo xv nv xs ns (5 x vec3), numSamples (int), weight (may be NULL) */
int rstr_gi_read(RstrGI*, float* indirectRgb, void* reservoirs);
int rstr_gi_indirect_device(RstrGI*, float** devIndirectIllum);                    /* device pointer of the own plane */
/* which tree the bounce rays walk when the scene's traversal mode is the traced one: 0 the traced tree (default), 1 the reference tree in
 * the reference's order; the hits are the same either way (validation switch) */
int rstr_gi_set_bounce_walk(RstrGI*, int traversal);
/* how the frame is launched in the traced mode: one kernel per frame (fused); a wavefront -- primary rays, one launch per iteration of
 * the path loop over the compacted live paths, resolve (staged); the staged form with the walks of every iteration as kernels of their own
 * over ray lists, persistent warps whose lanes fetch the next ray when they finish one (queued); or queued on real scenes and staged on
 * tiny ones (auto, the default: the rule of the direct path's pipelines).  Same result bit for bit in every form; the environment
 * variable RSTR_GI_PIPELINE=fused|staged|queued|auto sets the default of new handles (A/B measurements). */
enum { RSTR_GI_PIPELINE_FUSED = 0, RSTR_GI_PIPELINE_STAGED = 1, RSTR_GI_PIPELINE_QUEUED = 2, RSTR_GI_PIPELINE_AUTO = 3 };
int rstr_gi_set_pipeline(RstrGI*, int pipeline);
int rstr_gi_fallback_pixels(RstrGI*, unsigned int* count, int reset);              /* pixels recomputed with the reference-order walk */

#ifdef __cplusplus
}
#endif
#endif
