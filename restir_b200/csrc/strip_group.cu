// strip_group.cu -- the multi-GPU data plane of the ReSTIR DI frame: horizontal image strips, one process per GPU.
//
// Nothing in the reference to match (preview.cpp:112 pins device 0); SURVEY section 5 / 8e and BASELINE.json's north star are the
// spec: "halo rows of reservoirs (spatial radius plus reprojection margin) are exchanged peer-to-peer over NVLink each
// frame, and strips are gathered to GPU 0 for display".
//
// Every rank owns rows [row0,row1) of the image plus `halo` rows of its neighbours' planes on each side (capi.cu).  The
// planes a neighbour reads live in one allocation per frame (the exchange slab, capi_internal.h) whose CUDA IPC handle
// is handed to the neighbours ONCE (rstr_strip_group_handle / _connect: the caller moves 512-byte blobs between the
// processes with whatever it has -- MPI, torch.distributed, a pipe).  After that no library or host code is on the data path:
//   * k_halo_push (sender's stream, right behind phase A) stores the sender's edge rows of G-buffer / reservoir planes
//     straight into the neighbours' halo rows over NVLink with 16-byte stores, then releases a per-edge sequence flag in the
//     NEIGHBOUR's slab header (system-scope fence + store);
//   * k_flag_wait (receiver's stream, in front of phase B) spins on the flags in its OWN header until both neighbours' rows
//     of this exchange have landed;
//   * after the kernels that read the halo rows, the receiver acknowledges the sequence number in the SENDERS' headers; the next
//     push into the same rows waits for that acknowledgement (it is long there by then: a whole phase A lies in between);
//   * the tone-mapped strip is written by k_tonemap straight into rank 0's full-frame LDR slot, rank 0 waits for one
//     flag per rank and copies the frame to pinned host memory on a copy stream.
// Spin loops carry a clock64() time-out that raises an error flag instead of hanging the GPU when a peer dies.
#include <unistd.h>

#include <algorithm>

#include "capi_internal.h"

using namespace rs;

#define RS_STRIP_MAGIC 0x52535447u      /* "RSTG" */
#define RS_SPIN_TIMEOUT_CYCLES (20ll * 1000 * 1000 * 1000)   /* ~10 s at 2 GHz */

// 32-bit flags in the slab header
enum {
    RS_FL_ARRIVED_UP = 0,     // written by rank-1: its rows of exchange #seq are in my upper halo
    RS_FL_ARRIVED_DOWN = 1,   // written by rank+1
    RS_FL_ACK_UP = 2,         // written by rank-1: it has finished reading what I pushed in exchange #seq
    RS_FL_ACK_DOWN = 3,       // written by rank+1
    RS_FL_ERROR = 4,          // a spin loop timed out
    RS_FL_LDR_ACK = 5,        // written by rank 0: frames up to #seq have left its LDR slots
    RS_FL_PUSH_BLOCKS = 6,    // local: blocks of the running push kernel that are done
    RS_FL_LDR_ARRIVED = 16    // + rank, in rank 0's header: strip of frame #seq is in the LDR slot
};

struct StripHandle {            // what a rank publishes; padded to RSTR_STRIP_HANDLE_BYTES
    uint32_t magic, bytes;
    int rank, world, W, H, row0, row1, halo, bufRow0, bufRows, device, pid, pad;
    uint64_t slabPtr, slabBytes, slabOff[RS_XP_COUNT];
    uint64_t ldrPtr, ldrBytes;             // rank 0: RSTR_LDR_SLOTS full-frame LDR images
    cudaIpcMemHandle_t slabIpc, ldrIpc;
};
static_assert(sizeof(StripHandle) <= RSTR_STRIP_HANDLE_BYTES, "StripHandle");

struct Peer {
    bool present = false, mapped = false, ldrMapped = false;
    StripHandle h{};
    char* slab = nullptr;
    uchar4* ldr = nullptr;
    unsigned int* flags() const { return (unsigned int*)slab; }
};

struct RstrStripGroup {
    RstrFrame* f = nullptr;
    int rank = 0, world = 1;
    std::vector<Peer> peers;         // by rank; [rank] describes this rank itself (not mapped)
    uchar4* ldrFull = nullptr;       // rank 0: RSTR_LDR_SLOTS x W x H
    unsigned seq = 0;                // halo exchanges issued so far
    unsigned acked = 0;              // last exchange this rank has acknowledged to its neighbours
    unsigned pushed = 0, waited = 0; // last exchange pushed / waited for
    unsigned ldrSeq = 0;             // frames presented so far
    cudaStream_t copyStream = nullptr;
    cudaEvent_t evTone[RSTR_LDR_SLOTS] = {}, evCopied[RSTR_LDR_SLOTS] = {};
    bool slotBusy[RSTR_LDR_SLOTS] = {};
    bool connected = false;
};

// ------------------------------------------------------------------------------------------------ kernels
struct PushSeg { const char* src; char* dst; unsigned long long bytes; };
#define RS_PUSH_MAX_SEGS 12
struct PushArgs {
    PushSeg seg[RS_PUSH_MAX_SEGS];
    int n;
    unsigned int seq;
    unsigned int* arrived[2];                  // in the neighbours' headers (null: no neighbour on that side)
    const volatile unsigned int* ack[2];       // in my header
    unsigned int ackSeq;                       // the neighbours must have acknowledged this exchange before their halo rows are overwritten
    unsigned int* local;                       // my header
};

__device__ __forceinline__ bool spinUntil(const volatile unsigned int* flag, unsigned int value, unsigned int* err) {
    const long long t0 = clock64();
    while ((int)(*flag - value) < 0) {
        if (clock64() - t0 > RS_SPIN_TIMEOUT_CYCLES) { atomicExch(err, 1u); return false; }
        __nanosleep(200);
    }
    __threadfence_system();
    return true;
}

__global__ void __launch_bounds__(256) k_halo_push(const PushArgs a) {
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; s++)
            if (a.arrived[s] && a.ackSeq) spinUntil(a.ack[s], a.ackSeq, a.local + RS_FL_ERROR);
    }
    __syncthreads();
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    for (int i = 0; i < a.n; i++) {
        const PushSeg sg = a.seg[i];
        if ((((size_t)sg.src | (size_t)sg.dst | sg.bytes) & 15) == 0) {
            const uint4* s4 = (const uint4*)sg.src;
            uint4* d4 = (uint4*)sg.dst;
            const size_t n16 = sg.bytes / 16;
            size_t k = tid;
            for (; k + 3 * nth < n16; k += 4 * nth) {                  // four independent 16-byte loads in flight per thread
                const uint4 v0 = s4[k], v1 = s4[k + nth], v2 = s4[k + 2 * nth], v3 = s4[k + 3 * nth];
                d4[k] = v0; d4[k + nth] = v1; d4[k + 2 * nth] = v2; d4[k + 3 * nth] = v3;
            }
            for (; k < n16; k += nth) d4[k] = s4[k];
        } else {
            const unsigned int* s1 = (const unsigned int*)sg.src;      // every plane is a multiple of 4 bytes per pixel
            unsigned int* d1 = (unsigned int*)sg.dst;
            for (size_t k = tid; k < sg.bytes / 4; k += nth) d1[k] = s1[k];
        }
    }
    // the last block to finish releases the flags: its fence orders every block's stores (each fenced before counting)
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(a.local + RS_FL_PUSH_BLOCKS, 1u);
        if (done == gridDim.x - 1) {
            a.local[RS_FL_PUSH_BLOCKS] = 0;
            __threadfence_system();
            for (int s = 0; s < 2; s++)
                if (a.arrived[s]) *(volatile unsigned int*)a.arrived[s] = a.seq;
        }
    }
}

// waits until every listed flag has reached `value` (flags in this GPU's memory, written by peers)
struct WaitArgs { const volatile unsigned int* flag[8]; int n; unsigned int value; unsigned int* err; };
__global__ void k_flag_wait(const WaitArgs a) {
    if (threadIdx.x < a.n) spinUntil(a.flag[threadIdx.x], a.value, a.err);
}
// stores `value` into every listed flag (flags in peers' memory)
struct SetArgs { unsigned int* flag[8]; int n; unsigned int value; };
__global__ void k_flag_set(const SetArgs a) {
    __threadfence_system();
    if (threadIdx.x < a.n) *(volatile unsigned int*)a.flag[threadIdx.x] = a.value;
}

// ------------------------------------------------------------------------------------------------ helpers
static int xpOfPlane(const RstrFrame* f, int plane) {
    switch (plane) {
    case RSTR_PLANE_GEOM_CUR: return f->cur ? RS_XP_GEOM1 : RS_XP_GEOM0;
    case RSTR_PLANE_MATID_CUR: return f->cur ? RS_XP_MATID1 : RS_XP_MATID0;
    case RSTR_PLANE_RESV_HISTORY: return (f->resvOut ^ 1) ? RS_XP_RESV1 : RS_XP_RESV0;
    case RSTR_PLANE_RESV_OUT: return f->resvOut ? RS_XP_RESV1 : RS_XP_RESV0;
    case RSTR_PLANE_RESV_TEMP: return RS_XP_TEMP;
    case RSTR_PLANE_RESV_TEMP2: return RS_XP_TEMP2;
    }
    return -1;
}
static size_t xpPixelBytes(int xp) { return xp <= RS_XP_GEOM1 ? sizeof(float4) : (xp <= RS_XP_MATID1 ? sizeof(int) : sizeof(ResvD)); }

static void fillHandle(const RstrStripGroup* g, StripHandle& h) {
    const RstrFrame* f = g->f;
    memset(&h, 0, sizeof h);
    h.magic = RS_STRIP_MAGIC; h.bytes = sizeof(StripHandle);
    h.rank = g->rank; h.world = g->world; h.W = f->W; h.H = f->H; h.row0 = f->row0; h.row1 = f->row1; h.halo = f->halo;
    h.bufRow0 = f->bufRow0; h.bufRows = f->bufRows;
    cudaGetDevice(&h.device);
    h.pid = (int)getpid();
    h.slabPtr = (uint64_t)f->slab; h.slabBytes = f->slabBytes;
    for (int i = 0; i < RS_XP_COUNT; i++) h.slabOff[i] = f->slabOff[i];
    h.ldrPtr = (uint64_t)g->ldrFull; h.ldrBytes = g->ldrFull ? (uint64_t)RSTR_LDR_SLOTS * f->W * f->H * sizeof(uchar4) : 0;
}

extern "C" {

int rstr_strip_group_create(RstrFrame* f, int rank, int world, RstrStripGroup** out) {
    if (!f || !out || world < 1 || rank < 0 || rank >= world) return rsFail(RSTR_ERR_ARG, "rstr_strip_group_create: bad argument");
    if (world > 1 && f->row1 - f->row0 < f->halo)
        return rsFail(RSTR_ERR_ARG, "rstr_strip_group_create: a strip must be at least `halo` rows tall (halo rows come from the adjacent ranks only)");
    RstrStripGroup* g = new RstrStripGroup;
    g->f = f; g->rank = rank; g->world = world;
    g->peers.resize(world);
    if (rank == 0) {
        const size_t bytes = (size_t)RSTR_LDR_SLOTS * f->W * f->H * sizeof(uchar4);
        if (cudaMalloc((void**)&g->ldrFull, bytes) != cudaSuccess || cudaMemset(g->ldrFull, 0, bytes) != cudaSuccess) {
            delete g;
            return rsFail(RSTR_ERR_CUDA, "rstr_strip_group_create: cudaMalloc of the gather slots failed");
        }
    }
    cudaError_t e = cudaStreamCreateWithFlags(&g->copyStream, cudaStreamNonBlocking);
    for (int i = 0; i < RSTR_LDR_SLOTS && e == cudaSuccess; i++) {
        e = cudaEventCreateWithFlags(&g->evTone[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->evCopied[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) { rstr_strip_group_destroy(g); return rsFail(RSTR_ERR_CUDA, std::string("rstr_strip_group_create: ") + cudaGetErrorString(e)); }
    // the flags must read 0 before any peer can see them
    cudaMemsetAsync(f->slab, 0, RS_SLAB_HEADER, f->stream);
    cudaStreamSynchronize(f->stream);
    *out = g;
    return RSTR_OK;
}

int rstr_strip_group_handle(RstrStripGroup* g, void* blob) {
    if (!g || !blob) return rsFail(RSTR_ERR_ARG, "rstr_strip_group_handle: bad argument");
    StripHandle h;
    fillHandle(g, h);
    CU(cudaIpcGetMemHandle(&h.slabIpc, g->f->slab));
    if (g->ldrFull) CU(cudaIpcGetMemHandle(&h.ldrIpc, g->ldrFull));
    memset(blob, 0, RSTR_STRIP_HANDLE_BYTES);
    memcpy(blob, &h, sizeof h);
    return RSTR_OK;
}

// `blobs`: world x RSTR_STRIP_HANDLE_BYTES in rank order.  Maps the neighbours' slabs (rank 0: every rank's; every rank: rank 0's
// gather slots).  Ranks of the same process (tests) are reached through their raw pointers.
int rstr_strip_group_connect(RstrStripGroup* g, const void* blobs) {
    if (!g || !blobs) return rsFail(RSTR_ERR_ARG, "rstr_strip_group_connect: bad argument");
    if (g->connected) return rsFail(RSTR_ERR_ARG, "rstr_strip_group_connect: already connected");
    const RstrFrame* f = g->f;
    int myDev = 0;
    CU(cudaGetDevice(&myDev));
    for (int r = 0; r < g->world; r++) {
        Peer& p = g->peers[r];
        memcpy(&p.h, (const char*)blobs + (size_t)r * RSTR_STRIP_HANDLE_BYTES, sizeof(StripHandle));
        const StripHandle& h = p.h;
        if (h.magic != RS_STRIP_MAGIC || h.rank != r || h.world != g->world || h.W != f->W || h.H != f->H || h.halo != f->halo)
            return rsFail(RSTR_ERR_ARG, "rstr_strip_group_connect: handle of another group / image / halo");
        p.present = true;
    }
    for (int r = 0; r + 1 < g->world; r++)
        if (g->peers[r].h.row1 != g->peers[r + 1].h.row0) return rsFail(RSTR_ERR_ARG, "rstr_strip_group_connect: strips are not contiguous");
    if (g->peers[0].h.row0 != 0 || g->peers[g->world - 1].h.row1 != f->H) return rsFail(RSTR_ERR_ARG, "rstr_strip_group_connect: strips do not cover the image");
    for (int r = 0; r < g->world; r++) {
        if (r == g->rank) { g->peers[r].slab = (char*)f->slab; g->peers[r].ldr = g->ldrFull; continue; }
        Peer& p = g->peers[r];
        const bool neighbour = r == g->rank - 1 || r == g->rank + 1;
        const bool needSlab = neighbour || g->rank == 0 || r == 0;      // rank 0 acknowledges to everybody; everybody signals rank 0
        const bool needLdr = r == 0;
        const bool sameProcess = p.h.pid == (int)getpid();
        if (sameProcess && p.h.device != myDev) {
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, myDev, p.h.device));
            if (!can) return rsFail(RSTR_ERR_CUDA, "rstr_strip_group_connect: no peer access between the GPUs of two strips");
            cudaError_t e = cudaDeviceEnablePeerAccess(p.h.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
            (void)cudaGetLastError();
        }
        if (needSlab) {
            if (sameProcess) p.slab = (char*)p.h.slabPtr;
            else { CU(cudaIpcOpenMemHandle((void**)&p.slab, p.h.slabIpc, cudaIpcMemLazyEnablePeerAccess)); p.mapped = true; }
        }
        if (needLdr && p.h.ldrBytes) {
            if (sameProcess) p.ldr = (uchar4*)p.h.ldrPtr;
            else { CU(cudaIpcOpenMemHandle((void**)&p.ldr, p.h.ldrIpc, cudaIpcMemLazyEnablePeerAccess)); p.ldrMapped = true; }
        }
    }
    g->connected = true;
    return RSTR_OK;
}

int rstr_strip_group_destroy(RstrStripGroup* g) {
    if (!g) return RSTR_OK;
    if (g->f && g->f->stream) cudaStreamSynchronize(g->f->stream);
    if (g->copyStream) cudaStreamSynchronize(g->copyStream);
    for (Peer& p : g->peers) {
        if (p.mapped) cudaIpcCloseMemHandle(p.slab);
        if (p.ldrMapped) cudaIpcCloseMemHandle(p.ldr);
    }
    for (int i = 0; i < RSTR_LDR_SLOTS; i++) {
        if (g->evTone[i]) cudaEventDestroy(g->evTone[i]);
        if (g->evCopied[i]) cudaEventDestroy(g->evCopied[i]);
    }
    if (g->copyStream) cudaStreamDestroy(g->copyStream);
    cudaFree(g->ldrFull);
    delete g;
    return RSTR_OK;
}

// Acknowledge to the neighbours that everything queued so far on this rank's stream has finished with the halo rows of
// exchange #seq (called behind the kernels that read them).
static int ackNeighbours(RstrStripGroup* g) {
    if (g->acked == g->seq) return RSTR_OK;
    SetArgs a{};
    if (g->rank > 0) a.flag[a.n++] = g->peers[g->rank - 1].flags() + RS_FL_ACK_DOWN;              // I am its lower neighbour
    if (g->rank + 1 < g->world) a.flag[a.n++] = g->peers[g->rank + 1].flags() + RS_FL_ACK_UP;
    a.value = g->seq;
    if (a.n) { k_flag_set<<<1, 32, 0, g->f->stream>>>(a); rsCountLaunches(1); CU(cudaGetLastError()); }
    g->acked = g->seq;
    return RSTR_OK;
}

// Tell the neighbours that everything queued so far on this rank's stream has finished with the halo rows they pushed
// (issued behind the kernels that read them; rstr_strip_group_push issues it itself if the caller did not).
int rstr_strip_group_ack(RstrStripGroup* g) {
    if (!g) return rsFail(RSTR_ERR_ARG, "rstr_strip_group_ack: null group");
    if (g->world == 1 || !g->connected) return RSTR_OK;
    return ackNeighbours(g);
}

// First half of a halo exchange: this rank's edge rows of the selected planes (bit i = RSTR_PLANE_i) go into the
// neighbours' halo rows and the arrival flags are released behind them.
int rstr_strip_group_push(RstrStripGroup* g, unsigned int planeMask) {
    if (!g || !g->connected) return rsFail(RSTR_ERR_ARG, "rstr_strip_group_push: group not connected");
    if (g->pushed != g->waited) return rsFail(RSTR_ERR_ARG, "rstr_strip_group_push: the previous exchange has not been waited for");
    RstrFrame* f = g->f;
    int rc = rsFlushGBuffer(f);
    if (rc) return rc;
    if (g->world == 1) return RSTR_OK;
    if ((planeMask & (1u << RSTR_PLANE_RESV_TEMP2)) && (rc = rsEnsureTemp2(f))) return rc;
    if ((rc = ackNeighbours(g))) return rc;          // the previous exchange's rows have been read by everything queued before this call
    const unsigned seq = ++g->seq;
    g->pushed = seq;
    PushArgs a{};
    a.seq = seq; a.ackSeq = seq - 1; a.local = (unsigned int*)f->slab;
    a.ack[0] = a.local + RS_FL_ACK_UP; a.ack[1] = a.local + RS_FL_ACK_DOWN;
    const int W = f->W;
    size_t total = 0;
    for (int side = 0; side < 2; side++) {
        const int r = side == 0 ? g->rank - 1 : g->rank + 1;
        if (r < 0 || r >= g->world) continue;
        const Peer& p = g->peers[r];
        a.arrived[side] = p.flags() + (side == 0 ? RS_FL_ARRIVED_DOWN : RS_FL_ARRIVED_UP);     // I am its lower / upper neighbour
        // rows I own that the neighbour keeps as halo
        const int lo = std::max(f->row0, p.h.bufRow0), hi = std::min(f->row1, p.h.bufRow0 + p.h.bufRows);
        if (lo >= hi) continue;
        for (int plane = 0; plane < 6; plane++) {
            if (!(planeMask & (1u << plane))) continue;
            const int xp = xpOfPlane(f, plane);
            const size_t px = xpPixelBytes(xp);
            if (a.n >= RS_PUSH_MAX_SEGS) return rsFail(RSTR_ERR_ARG, "rstr_strip_group_exchange: too many planes");
            PushSeg& sg = a.seg[a.n++];
            sg.src = (const char*)f->slab + f->slabOff[xp] + (size_t)(lo - f->bufRow0) * W * px;
            sg.dst = p.slab + p.h.slabOff[xp] + (size_t)(lo - p.h.bufRow0) * W * px;
            sg.bytes = (size_t)(hi - lo) * W * px;
            total += sg.bytes;
        }
    }
    // enough blocks to fill the NVLink pipes, few enough to leave the SMs to the frame kernels of other streams
    const int blocks = (int)std::min<size_t>(128, std::max<size_t>(1, total / (256 * 16 * 4)));
    k_halo_push<<<blocks, 256, 0, f->stream>>>(a);
    rsCountLaunches(1);
    CU(cudaGetLastError());
    return RSTR_OK;
}

// Second half: the frame's stream waits until the neighbours' rows of the exchange pushed last have arrived here.
int rstr_strip_group_wait(RstrStripGroup* g) {
    if (!g || !g->connected) return rsFail(RSTR_ERR_ARG, "rstr_strip_group_wait: group not connected");
    if (g->world == 1 || g->waited == g->pushed) return RSTR_OK;
    RstrFrame* f = g->f;
    unsigned int* local = (unsigned int*)f->slab;
    WaitArgs w{};
    if (g->rank > 0) w.flag[w.n++] = local + RS_FL_ARRIVED_UP;
    if (g->rank + 1 < g->world) w.flag[w.n++] = local + RS_FL_ARRIVED_DOWN;
    w.value = g->pushed; w.err = local + RS_FL_ERROR;
    k_flag_wait<<<1, 32, 0, f->stream>>>(w);
    g->waited = g->pushed;
    rsCountLaunches(1);
    CU(cudaGetLastError());
    return RSTR_OK;
}

// One halo exchange = push + wait.  (Ranks that share ONE process, as in the tests, must issue every rank's push before any
// rank's wait -- rstr_strip_group_push / _wait -- because streams of one process share hardware work queues: a waiting
// kernel may sit in front of the push it waits for.  One process per GPU, the deployment, has no such coupling.)
int rstr_strip_group_exchange(RstrStripGroup* g, unsigned int planeMask) {
    int rc = rstr_strip_group_push(g, planeMask);
    if (rc) return rc;
    return rstr_strip_group_wait(g);
}

// One whole frame of a strip (runCuda's sequence, main.cpp:164-183): G-buffer + phase A on the strip's own rows, ONE exchange
// carrying every halo row anybody needs -- what this frame's spatial pass reads (current G-buffer rows, post-temporal
// reservoirs) and the history reservoirs phase A just wrote, which the NEXT frame's temporal step reads -- then phase B;
// with spatialPasses > 1 the plane each pass publishes is exchanged before the next pass.
// _frame_begin = up to and including the push; _frame_end = the wait and everything behind it.
static unsigned frameMask(const RstrParams* prm) {
    const bool spatial = (prm->reuse & RSTR_REUSE_SPATIAL) != 0, temporal = (prm->reuse & RSTR_REUSE_TEMPORAL) != 0;
    unsigned mask = 0;
    if (spatial) mask |= (1u << RSTR_PLANE_GEOM_CUR) | (1u << RSTR_PLANE_MATID_CUR) | (1u << RSTR_PLANE_RESV_TEMP);
    if (temporal) mask |= (1u << RSTR_PLANE_GEOM_CUR) | (1u << RSTR_PLANE_MATID_CUR) | (1u << RSTR_PLANE_RESV_OUT);
    return mask;
}

int rstr_strip_group_frame_begin(RstrStripGroup* g, const RstrCamera* cam, const RstrParams* prm, int looper, int iter) {
    if (!g || !cam || !prm) return rsFail(RSTR_ERR_ARG, "rstr_strip_group_frame_begin: bad argument");
    RstrFrame* f = g->f;
    if (g->world > 1 && !g->connected) return rsFail(RSTR_ERR_ARG, "rstr_strip_group_frame_begin: group not connected");
    if (g->world > 1) f->renderHalo = false;        // halo rows of the G-buffer come from the neighbours
    int rc = rstr_gbuffer_render(f, cam);
    if (rc) return rc;
    if ((rc = rstr_restir_phase_a(f, cam, prm, looper, iter))) return rc;
    if (g->world > 1 && frameMask(prm) && (rc = rstr_strip_group_push(g, frameMask(prm)))) return rc;
    return RSTR_OK;
}

int rstr_strip_group_frame_end(RstrStripGroup* g, const RstrCamera* cam, const RstrParams* prm, int looper, int iter) {
    if (!g || !cam || !prm) return rsFail(RSTR_ERR_ARG, "rstr_strip_group_frame_end: bad argument");
    RstrFrame* f = g->f;
    const int passes = (prm->reuse & RSTR_REUSE_SPATIAL) ? (prm->spatialPasses < 1 ? 1 : prm->spatialPasses) : 0;
    int rc = rstr_strip_group_wait(g);
    if (rc) return rc;
    if (passes > 1 && g->world > 1)
        return rsFail(RSTR_ERR_ARG, "rstr_strip_group_frame_end: with spatialPasses > 1 use rstr_strip_group_frame, or rstr_restir_phase_b_pass + rstr_strip_group_push / _wait per pass");
    for (int pass = 1; pass <= (passes ? passes : 1); pass++)
        if ((rc = rstr_restir_phase_b_pass(f, cam, prm, looper, iter, pass))) return rc;
    if ((rc = rstr_strip_group_ack(g))) return rc;
    return rstr_gbuffer_update(f, cam);
}

int rstr_strip_group_frame(RstrStripGroup* g, const RstrCamera* cam, const RstrParams* prm, int looper, int iter) {
    int rc = rstr_strip_group_frame_begin(g, cam, prm, looper, iter);
    if (rc) return rc;
    RstrFrame* f = g->f;
    const int passes = (prm->reuse & RSTR_REUSE_SPATIAL) ? (prm->spatialPasses < 1 ? 1 : prm->spatialPasses) : 0;
    if ((rc = rstr_strip_group_wait(g))) return rc;
    for (int pass = 1; pass <= (passes ? passes : 1); pass++) {
        if ((rc = rstr_restir_phase_b_pass(f, cam, prm, looper, iter, pass))) return rc;
        if ((rc = rstr_strip_group_ack(g))) return rc;
        if (g->world > 1 && pass < passes && (rc = rstr_strip_group_exchange(g, 1u << ((pass & 1) ? RSTR_PLANE_RESV_TEMP2 : RSTR_PLANE_RESV_TEMP)))) return rc;
    }
    return rstr_gbuffer_update(f, cam);
}

// Display step of a strip (copyImageToPBO, pathtrace.cu:108-113, + the gather of the north star): tone-map this rank's rows
// straight into rank 0's full-frame LDR slot `slot`, signal rank 0; rank 0 waits for every rank's signal and copies the
// frame to hostLdr (W x H x uchar4) on a copy stream.  Returns at once; rank 0 reads the image after rstr_strip_group_wait_host.
int rstr_strip_group_present(RstrStripGroup* g, int toneMapping, void* hostLdr, size_t bytes, int slot) {
    if (!g || slot < 0 || slot >= RSTR_LDR_SLOTS) return rsFail(RSTR_ERR_ARG, "rstr_strip_group_present: bad argument");
    if (g->world > 1 && !g->connected) return rsFail(RSTR_ERR_ARG, "rstr_strip_group_present: group not connected");
    RstrFrame* f = g->f;
    const size_t P = (size_t)f->W * f->H;
    if (g->rank == 0 && hostLdr && bytes != P * sizeof(uchar4)) return rsFail(RSTR_ERR_ARG, "rstr_strip_group_present: size mismatch");
    int rc = rsFlushGBuffer(f);
    if (rc) return rc;
    const unsigned seq = ++g->ldrSeq;
    unsigned int* local = (unsigned int*)f->slab;
    uchar4* dstBase = g->rank == 0 ? g->ldrFull : g->peers[0].ldr;
    if (!dstBase) return rsFail(RSTR_ERR_ARG, "rstr_strip_group_present: rank 0's gather slots are not mapped");
    // the slot's previous frame (#seq - RSTR_LDR_SLOTS) must have left rank 0
    if (seq > RSTR_LDR_SLOTS) {
        if (g->rank == 0) { if (g->slotBusy[slot]) CU(cudaStreamWaitEvent(f->stream, g->evCopied[slot], 0)); }
        else {
            WaitArgs w{};
            w.flag[w.n++] = local + RS_FL_LDR_ACK; w.value = seq - RSTR_LDR_SLOTS; w.err = local + RS_FL_ERROR;
            k_flag_wait<<<1, 32, 0, f->stream>>>(w);
            rsCountLaunches(1);
        }
    }
    const size_t off = (size_t)(f->row0 - f->bufRow0) * f->W, n = (size_t)(f->row1 - f->row0) * f->W;
    launchTonemap(f->radiance + 3 * off, dstBase + (size_t)slot * P + (size_t)f->row0 * f->W, n, toneMapping, 1.f, f->stream);
    rsCountLaunches(1);
    if (g->rank != 0) {
        SetArgs s{};
        s.flag[s.n++] = g->peers[0].flags() + RS_FL_LDR_ARRIVED + g->rank; s.value = seq;
        k_flag_set<<<1, 32, 0, f->stream>>>(s);
        rsCountLaunches(1);
        CU(cudaGetLastError());
        return RSTR_OK;
    }
    // rank 0: wait for the other strips (flags in my header), then D2H on the copy stream
    for (int r0 = 1; r0 < g->world; r0 += 8) {
        WaitArgs w{};
        for (int r = r0; r < g->world && r < r0 + 8; r++) w.flag[w.n++] = local + RS_FL_LDR_ARRIVED + r;
        w.value = seq; w.err = local + RS_FL_ERROR;
        k_flag_wait<<<1, 32, 0, f->stream>>>(w);
        rsCountLaunches(1);
    }
    CU(cudaEventRecord(g->evTone[slot], f->stream));
    CU(cudaStreamWaitEvent(g->copyStream, g->evTone[slot], 0));
    if (hostLdr) CU(cudaMemcpyAsync(hostLdr, g->ldrFull + (size_t)slot * P, P * sizeof(uchar4), cudaMemcpyDeviceToHost, g->copyStream));
    // tell every rank that frames up to #seq have left the slots
    for (int r0 = 1; r0 < g->world; r0 += 8) {
        SetArgs s{};
        for (int r = r0; r < g->world && r < r0 + 8; r++) s.flag[s.n++] = g->peers[r].flags() + RS_FL_LDR_ACK;
        s.value = seq;
        k_flag_set<<<1, 32, 0, g->copyStream>>>(s);
        rsCountLaunches(1);
    }
    CU(cudaEventRecord(g->evCopied[slot], g->copyStream));
    g->slotBusy[slot] = true;
    CU(cudaGetLastError());
    return RSTR_OK;
}

int rstr_strip_group_wait_host(RstrStripGroup* g, int slot) {
    if (!g || slot < 0 || slot >= RSTR_LDR_SLOTS) return rsFail(RSTR_ERR_ARG, "rstr_strip_group_wait_host: bad argument");
    if (g->slotBusy[slot]) CU(cudaEventSynchronize(g->evCopied[slot]));
    return RSTR_OK;
}

// 0 = healthy; 1 = a wait on a peer timed out (the frame's results are invalid)
int rstr_strip_group_error(RstrStripGroup* g, int* flag) {
    if (!g || !flag) return rsFail(RSTR_ERR_ARG, "rstr_strip_group_error: bad argument");
    unsigned int v = 0;
    CU(cudaMemcpyAsync(&v, (unsigned int*)g->f->slab + RS_FL_ERROR, sizeof v, cudaMemcpyDeviceToHost, g->f->stream));
    CU(cudaStreamSynchronize(g->f->stream));
    *flag = (int)v;
    return RSTR_OK;
}

}  // extern "C"
