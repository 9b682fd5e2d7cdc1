// device_types.h -- device-side scene / frame descriptors shared by kernels.cu and capi.cu.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "scene_host.h"

namespace rs {

#define RS_STACK_DEPTH 96      /* per-lane traversal stack entries (reference-order walk, shadow rays); deeper trees are rejected at scene creation */
#define RS_PACKET_STACK 64     /* per-warp stack entries of the packet walk of the traced tree (one entry per level at most) */

#define RS_TRAVERSAL_FAST 0     /* binned-SAH tree; reference-order walk only for near-axis and near-tie rays */
#define RS_TRAVERSAL_EXACT 1    /* reference-order walk of the reference tree for every ray */

struct DevScene {
    const float4* nodes;       // PackedNode[] (reference tree), 4 x float4 each
    const float4* fastNodes;   // FastNode[] (traced tree), 4 x float4 each
    const float4* triGeom;     // TriGeom[] in the traced tree's leaf order, 3 x float4 each
    const float4* triNorm;     // TriNorm[] in original primitive order, 3 x float4 each
    const int* primToFast;     // original primitive id -> triGeom index
    const int* rank;           // [6][numTris] visiting rank in the reference's orderings (near-tie resolution)
    int numTris;
    int numFastNodes;
    int fastRoot;
    int traversal;             // RS_TRAVERSAL_*
    unsigned int* fallbackRays;  // [3] pixels recomputed with the reference-order walk: G-buffer, ReSTIR phase A, PTDirect
    float fastRootMin[3], fastRootMax[3];
    const RstrMaterial* materials;
    const float2* alias;       // AliasEntry[] as {prob, failId bits}
    const float4* lights;      // LightRec[], 4 x float4 each
    int numLights;
    int rootRef;               // packed index of the root, or ~primId when the scene is a single triangle
    float rootMin[3], rootMax[3];
    // textures / environment map (all null / 0 / -1 for scenes without them: the kernels then skip that code)
    const float4* texData;     // float4 texels of all textures
    const int4* texInfo;       // per texture {w, h, first texel, 0}
    const float4* triUV;       // TriUV[] in original primitive order, 2 x float4 each
    int anyMaps;               // some material has a map id != -1
    int envTex;                // texture index of the environment map, or -1
    int envLen;                // envW * envH
    const float2* envAlias;    // envMapSampler
    const float4* envDir;      // per texel: direction of its centre
    float sumLightPowerInv;
};

// Reservoir in HBM: 32 B, one aligned sector, 2 x LDG/STG.128.  Li is not stored: it is the per-light constant
// lights[lightId].Le (the reference stores Li = lightUnitRadiance[lightId], scene.h:420), lightId < 0 <=> Li = 0.
struct alignas(16) ResvD {
    float wi[3];
    float dist;
    float weight;
    int M;
    int lightId;
    int pad;
};
static_assert(sizeof(ResvD) == 32, "ResvD");

// per-pixel state handed from phase A to phase B when spatial reuse is on (32 B)
struct alignas(16) HitRec {
    float n[3];
    int matId;       // >= 0: shaded pixel; -1: phase A already wrote the radiance (miss / emitter)
    float wo[3];
    uint32_t rng;
};
static_assert(sizeof(HitRec) == 32, "HitRec");

struct CamDev {
    float position[3];
    float right[3], up[3], view[3];
    float rotInv[9];           // column-major
    float aspect, tanFovY, focalDist;
    float pixelSizeX, pixelSizeY;
    float resX, resY;
};

struct FrameDev {
    int W, H;                  // full image
    int rowLo, rowHi;          // rows this launch computes
    int bufRow0, bufRows;      // rows resident in the per-pixel planes (strip + halo)
    // G-buffer planes (index = (y - bufRow0) * W + x)
    float4* geom[2];           // {n.xyz, depth}; [cur], [last]
    int* matId[2];
    float4* albedoMotion;      // {albedo.xyz, motion (int bits)}
    float4* radiance4;         // unused
    float* radiance;           // 3 floats per pixel (devDirectIllum layout)
    ResvD* resvOut;            // history written this frame
    const ResvD* resvIn;       // history of the previous frame
    ResvD* resvTemp;
    ResvD* resvStage;          // staged phase A: the reservoir between k_candidates and k_temporal (resvTemp, or resvTemp2 when spatial reuse is off)
    HitRec* hit;
    float4* hitPos;            // unbiased mode: the shaded point {pos.xyz, -} (allocated on first use)
    float2* hitMR;             // {metallic, roughness} of the shaded point; only allocated for scenes with such maps
    int* queue;                // pixels deferred to the reference-order fix-up kernel
    unsigned int* queueCount;  // [0] fix-up queue length; [4 + 2b], [5 + 2b]: band b's shaded-pixel queue length and k_shadow's cursor into it
    unsigned int* shadeCount;  // staged phase A, this launch's band: {queue length, k_shadow's cursor}
    int* shadeQueue;           // staged phase A, this launch's band: pixels (global index) whose jittered ray hit a shaded surface
    unsigned int* haloMiss;    // count of neighbour / reprojection reads that fell outside the resident rows
    unsigned int* motionRows;  // running max |row(motion) - row| of the reprojected pixels (bound for the temporal halo)
    unsigned long long* rowCost;  // optional [ceil(H / 8)] cycle accumulators (rstr_frame_row_cost), else null
};

// ReSTIR GI (gi_kernels.inl): the indirect reservoirs and devIndirectIllum of a full frame.  A reservoir is one 64-byte record
// {Lo, weight} {xv, numSamples} {nv, ns.x} {xs, ns.y} plus ns.z in a 4-byte plane (68 B per pixel, as Reservoir<IndirectLiSample>).
struct GIDev {
    float4* resvOut;           // devIndTemporalReservoir: written this frame
    const float4* resvIn;      // devIndLastTemporalReservoir: the previous frame's
    float* nszOut;
    const float* nszIn;
    float* indirect;           // 3 floats per pixel (devIndirectIllum layout)
    unsigned int* fallback;    // pixels recomputed with the reference-order walk
    int maxDepth;              // Settings::traceDepth
    int reuse;                 // bit 0: temporal (ReservoirReuse::Temporal)
    int first;                 // ReSTIRFirstFrame
    int iter;
    int bounceWalk;            // RS_TRAVERSAL_*: which tree the bounce rays walk in the traced mode (same hits either way)
    // staged form (k_gi_primary / k_gi_bounce / k_gi_resolve); null in the one-kernel form
    float4* pix;               // [7][pixStride] what the stages hand on per pixel: primary surface, xv / nv, xs / ns, final Lo + RNG state
    int* pixStatus;            // 0: the jittered ray left the scene or hit an emitter, 1: shaded, 2: undecided ray (fix-up queue)
    float4* pathQ[2];          // live paths between two bounces, 96 B each, ping-pong
    unsigned int* pathCount;   // [d]: paths entering bounce d (1-based); zeroed per frame
    size_t pixStride;          // pixels of the frame
    // ray-queue form (k_gi_head / k_gi_walk_shadow / k_gi_walk_closest / k_gi_tail per depth); null otherwise
    unsigned int* closestList; // slots of the live-path queue whose bounce ray is to be walked
    unsigned int* shadowList;  // ... whose next-event segment is to be tested
    unsigned int* walkCount;   // [4 * depth + k]: closest rays, the closest walker's cursor, shadow rays, the shadow walker's cursor
    float4* hit;               // per slot: {bx, by, prim, undecided} of the bounce ray
    int* occ;                  // per slot: 1 occluded, 0 free, -1 undecided
};

}  // namespace rs

// A plain load that only decides whether an atomic is worth issuing (a running maximum, the first value of a CAS loop): racing with other
// blocks' atomics is harmless -- a stale value costs one redundant atomic or one more CAS iteration.  The ThreadSanitizer build of tests/emu
// reads it atomically so that its report list stays empty for real findings.
#ifdef RS_HOST_EMU
#define RS_PEEK(p) __atomic_load_n(p, __ATOMIC_RELAXED)
#else
#define RS_PEEK(p) (*(p))
#endif
