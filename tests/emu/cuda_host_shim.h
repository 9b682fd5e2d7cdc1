// cuda_host_shim.h -- TEST INFRASTRUCTURE ONLY.  Lets g++ compile restir_b200/csrc/kernels.cu (with RS_HOST_EMU) so that the device
// functions of the kernels can be run, one "thread" per call, on the CPU and compared with the oracle bit for bit (tests/test_device_code_on_host.py).
// Nothing here is linked into, loaded by or reachable from the product (restir_b200/): the product's only compute path is the CUDA library.
#pragma once
#include <cuda_runtime.h>   // vector types / make_float4 / cudaStream_t for the host; __device__ & co become ignored attributes
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef EMU_STD_THREADS
#include <atomic>
#include <thread>
#endif

#define RS_HOST_EMU 1
// kernels and __noinline__ device functions must not be emitted unless the emulation calls them: several contain PTX inline assembly
#undef __global__
#define __global__ static inline
#undef __noinline__
#define __noinline__ inline
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __grid_constant__
#define __grid_constant__
// __shared__ arrays: one copy per OS thread -- what the 32 fibers of a warp share under emuRunWarp (the kernels emulated that way keep
// per-lane columns [..][threadIdx.x] in them); a plain per-call array in the one-thread-per-call emulation
#undef __shared__
#define __shared__ static thread_local

struct EmuIdx { unsigned x = 0, y = 0, z = 0; };
static thread_local EmuIdx threadIdx, blockIdx, blockDim, gridDim;

template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline int __float2int_rz(float x) {            // cvt.rzi.s32.f32: saturating, NaN -> 0
    if (x != x) return 0;
    if (x >= 2147483648.f) return 2147483647;
    if (x <= -2147483648.f) return (-2147483647 - 1);
    return (int)x;
}
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
// CUDA's atomics are relaxed and the kernels order them with __threadfence() where it matters (bvh_gpu.cu's bottom-up pass: "the second child
// to arrive finishes the node").  TSan does not model fences, so in its build the read-modify-writes themselves carry acquire / release --
// the standard reading of the fence + atomic idiom; elsewhere they stay relaxed.
#if defined(__SANITIZE_THREAD__)
#define EMU_RMW __ATOMIC_ACQ_REL
#else
#define EMU_RMW __ATOMIC_RELAXED
#endif
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, EMU_RMW); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, EMU_RMW); }
static inline unsigned atomicMax(unsigned* p, unsigned v) {
    unsigned o = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (o < v && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return o;
}
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, EMU_RMW); }
static inline int atomicCAS(int* p, int expected, int desired) { __atomic_compare_exchange_n(p, &expected, desired, false, EMU_RMW, __ATOMIC_RELAXED); return expected; }
static inline int atomicMax(int* p, int v) {
    int o = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (o < v && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return o;
}
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline long long clock64() { return 0; }
static inline void __trap() { abort(); }
// ---- warps.  Outside emuRunWarp a "warp" has one lane (the per-pixel emulation: one call = one thread).  Inside emuRunWarp the 32 lanes of
// a warp are 32 fibers (own stacks, switched by hand) on one OS thread: a lane runs until its next warp-level intrinsic, deposits its value and yields; when
// every live lane has deposited, the lanes are resumed in turn and combine the values.  All lanes of convergent code reach the same
// sequence of intrinsics, so pass k of the scheduler is synchronisation point k of every lane; the deposits are double-buffered by the
// parity of k (a resumed lane may reach point k + 1 and deposit again before its neighbours have read point k).  Lanes that have
// returned from the kernel contribute 0 / are skipped, as exited threads do on the GPU.
#define EMU_FIBER_STACK (512 * 1024)
#if defined(__x86_64__)
// switch stacks: callee-saved registers of the System V ABI on the old stack, its pointer to *from, the same from `to` (swapcontext costs
// a sigprocmask system call per switch, and a packet walk synchronises its lanes a dozen times per node)
typedef void* EmuCtx;
__attribute__((naked, noinline)) static void emuSwitch(EmuCtx* from, EmuCtx to) {
    asm volatile(
        "pushq %rbp\n\tpushq %rbx\n\tpushq %r12\n\tpushq %r13\n\tpushq %r14\n\tpushq %r15\n\t"
        "movq %rsp, (%rdi)\n\t"
        "movq %rsi, %rsp\n\t"
        "popq %r15\n\tpopq %r14\n\tpopq %r13\n\tpopq %r12\n\tpopq %rbx\n\tpopq %rbp\n\t"
        "ret\n\t");
}
static inline void emuMakeCtx(EmuCtx* ctx, char* stack, size_t size, void (*entry)()) {
    void** sp = (void**)(((size_t)stack + size) & ~(size_t)15);
    *--sp = nullptr;                             // the entry's "return address": it never returns
    *--sp = (void*)entry;                        // popped by the first switch's ret: entry runs with rsp + 8 aligned to 16, as after a call
    for (int i = 0; i < 6; i++) *--sp = nullptr; // rbp rbx r12 r13 r14 r15
    *ctx = sp;
}
#else
#include <ucontext.h>
typedef ucontext_t EmuCtx;
static inline void emuSwitch(EmuCtx* from, EmuCtx& to) { swapcontext(from, &to); }
#endif
#if defined(__SANITIZE_THREAD__)
// TSan keeps a shadow call stack per context: it has to be told about every switch (and orders the two sides of a switch, so the lanes of
// a warp, which run strictly one after the other, are never reported against each other; warps on different OS threads are)
extern "C" {
void* __tsan_get_current_fiber(void);
void* __tsan_create_fiber(unsigned flags);
void __tsan_destroy_fiber(void* fiber);
void __tsan_switch_to_fiber(void* fiber, unsigned flags);
}
#define EMU_TSAN_SWITCH(fiber) __tsan_switch_to_fiber(fiber, 0)
#else
#define EMU_TSAN_SWITCH(fiber) ((void)0)
#endif
#define EMU_MAX_LANES 256                        /* 32: a warp; up to 256: a whole block whose only synchronisation is __syncthreads (emuLaunchBlock) */
struct EmuWarp {
    void* tsanSched;
    void* tsanLane[EMU_MAX_LANES];
    EmuCtx sched, ctx[EMU_MAX_LANES];
    bool done[EMU_MAX_LANES];
    int lane;                                    // the lane that is running
    int nLanes;
    bool block;                                  // the lanes are a thread block: __syncthreads is their synchronisation point
    unsigned sync[EMU_MAX_LANES];                // synchronisation points passed, per lane
    unsigned long long val[2][EMU_MAX_LANES];
    bool has[2][EMU_MAX_LANES];                  // the lane has deposited at this point (false: it had exited before)
    void (*body)(void*);
    void* arg;
};
static thread_local EmuWarp* emuWarp = nullptr;
static thread_local char* emuStacks = nullptr;
// deposit `mine`, wait for the warp, return the buffer of this synchronisation point (entries of exited lanes are 0)
static inline const unsigned long long* emuExchange(unsigned long long mine) {
    EmuWarp* w = emuWarp;
    const int l = w->lane;
    const unsigned k = w->sync[l]++ & 1u;
    w->val[k][l] = mine;
    w->has[k][l] = true;
    EMU_TSAN_SWITCH(w->tsanSched);
    emuSwitch(&w->ctx[l], w->sched);
    return emuWarp->val[k];
}
static void emuTrampoline() {
    EmuWarp* w = emuWarp;
    w->body(w->arg);
    w = emuWarp;
    // what the lane deposited at its last synchronisation point is still being read by the lanes resumed after it in this pass (the scheduler
    // clears it after the pass); at the next point it counts as exited
    const int l = w->lane;
    const unsigned next = w->sync[l] & 1u;
    w->done[l] = true;
    w->val[next][l] = 0; w->has[next][l] = false;
#if defined(__x86_64__)
    for (;;) { EMU_TSAN_SWITCH(w->tsanSched); emuSwitch(&w->ctx[l], w->sched); }    // never resumed: the scheduler skips finished lanes
#endif
}
// run `body(arg)` as the 32 lanes tidBase .. tidBase + 31 of one warp; blockIdx / blockDim / gridDim are the caller's
static inline void emuRunLanes(unsigned tidBase, int nLanes, bool block, void (*body)(void*), void* arg);
static inline void emuRunWarp(unsigned tidBase, void (*body)(void*), void* arg) { emuRunLanes(tidBase, 32, false, body, arg); }

static inline int __any_sync(unsigned, int p) {
    if (!emuWarp) return p;
    const unsigned long long* v = emuExchange(p ? 1 : 0);
    int r = 0;
    for (int i = 0; i < 32; i++) r |= (int)v[i];
    return r;
}
static inline const bool* emuHas(const unsigned long long* v) { return emuWarp->has[v == emuWarp->val[0] ? 0 : 1]; }
static inline int __all_sync(unsigned, int p) {
    if (!emuWarp) return p;
    const unsigned long long* v = emuExchange(p ? 1 : 0);
    const bool* has = emuHas(v);
    int r = 1;
    for (int i = 0; i < 32; i++) if (has[i]) r &= (int)v[i];
    return r;
}
static inline unsigned __ballot_sync(unsigned, int p) {
    if (!emuWarp) return p ? 1u : 0u;
    const unsigned long long* v = emuExchange(p ? 1 : 0);
    unsigned m = 0;
    for (int i = 0; i < 32; i++) m |= (unsigned)v[i] << i;
    return m;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) {
    if (!emuWarp) return v;
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    unsigned long long bits = 0;
    memcpy(&bits, &v, sizeof(T));
    const unsigned long long* all = emuExchange(bits);
    T r;
    memcpy(&r, &all[src & 31], sizeof(T));
    return r;
}
static inline unsigned __reduce_min_sync(unsigned, unsigned x) {
    if (!emuWarp) return x;
    const unsigned long long* v = emuExchange(x);
    const bool* has = emuHas(v);
    unsigned r = 0xffffffffu;
    for (int i = 0; i < 32; i++) if (has[i] && (unsigned)v[i] < r) r = (unsigned)v[i];
    return r;
}
static inline unsigned __reduce_max_sync(unsigned, unsigned x) {
    if (!emuWarp) return x;
    const unsigned long long* v = emuExchange(x);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) if ((unsigned)v[i] > r) r = (unsigned)v[i];
    return r;
}
static inline unsigned __reduce_or_sync(unsigned, unsigned x) {
    if (!emuWarp) return x;
    const unsigned long long* v = emuExchange(x);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= (unsigned)v[i];
    return r;
}
static inline unsigned __reduce_add_sync(unsigned, unsigned x) {
    if (!emuWarp) return x;
    const unsigned long long* v = emuExchange(x);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r += (unsigned)v[i];
    return r;
}
static inline void __syncwarp(unsigned = 0xffffffffu) { if (emuWarp) emuExchange(0); }
static inline void __syncthreads() { if (emuWarp && emuWarp->block) emuExchange(0); }   // warps of a grid launched with emuLaunch share nothing
static inline unsigned __match_any_sync(unsigned, unsigned) { return 1u; }          // one-lane form only (not used under emuRunWarp)
static inline unsigned __activemask() { return 1u; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
// "shared-space addresses" (the kernels keep some as 32-bit cursors): offsets from a per-thread anchor; __shared__ arrays are thread_local
// statics of the same module, so the offsets fit 32 bits
static thread_local char emuSharedAnchor;
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)(unsigned)(int)((const char*)p - &emuSharedAnchor); }
static inline void* emuSharedPtr(unsigned off) { return (void*)((uintptr_t)&emuSharedAnchor + (uintptr_t)(intptr_t)(int)off); }   // integer arithmetic: not an offset INTO the anchor object


#if defined(__SANITIZE_ADDRESS__)
extern "C" void __asan_unpoison_memory_region(void const volatile* addr, size_t size);
#endif
static inline void emuRunLanes(unsigned tidBase, int nLanes, bool block, void (*body)(void*), void* arg) {
    const size_t stackSize = ((size_t)32 * EMU_FIBER_STACK / nLanes) & ~(size_t)4095;      // 512 KB per lane of a warp, 64 KB per lane of a 256-thread block
    if (!emuStacks) emuStacks = (char*)malloc((size_t)32 * EMU_FIBER_STACK);
#if defined(__SANITIZE_ADDRESS__)
    // the previous warp's lanes never returned from their trampolines: the redzones of their frames are still poisoned
    __asan_unpoison_memory_region(emuStacks, (size_t)32 * EMU_FIBER_STACK);
#endif
    EmuWarp w;
    memset(w.done, 0, sizeof w.done); memset(w.sync, 0, sizeof w.sync); memset(w.val, 0, sizeof w.val); memset(w.has, 0, sizeof w.has);
    w.body = body; w.arg = arg; w.lane = 0; w.nLanes = nLanes; w.block = block;
    for (int l = 0; l < nLanes; l++) {
#if defined(__x86_64__)
        emuMakeCtx(&w.ctx[l], emuStacks + (size_t)l * stackSize, stackSize, emuTrampoline);
#else
        getcontext(&w.ctx[l]);
        w.ctx[l].uc_stack.ss_sp = emuStacks + (size_t)l * stackSize;
        w.ctx[l].uc_stack.ss_size = stackSize;
        w.ctx[l].uc_link = &w.sched;
        makecontext(&w.ctx[l], emuTrampoline, 0);
#endif
    }
#if defined(__SANITIZE_THREAD__)
    w.tsanSched = __tsan_get_current_fiber();
    for (int l = 0; l < nLanes; l++) w.tsanLane[l] = __tsan_create_fiber(0);
#endif
    emuWarp = &w;
    for (;;) {
        bool any = false;
        for (int l = 0; l < nLanes; l++)
            if (!w.done[l]) {
                any = true;
                w.lane = l;
                threadIdx.x = tidBase + l;
                EMU_TSAN_SWITCH(w.tsanLane[l]);
                emuSwitch(&w.sched, w.ctx[l]);
            }
        if (!any) break;
        for (int l = 0; l < nLanes; l++)
            if (w.done[l]) { w.val[0][l] = w.val[1][l] = 0; w.has[0][l] = w.has[1][l] = false; }
    }
    emuWarp = nullptr;
#if defined(__SANITIZE_THREAD__)
    for (int l = 0; l < nLanes; l++) __tsan_destroy_fiber(w.tsanLane[l]);
#endif
}
// a grid of blocks of 32 * warpsPerBlock threads (128 unless said otherwise); the warps of the grid run on the OpenMP threads, each warp
// on one of them
template <typename F> static inline void emuLaunch(unsigned gridX, unsigned gridY, F kernelCall, unsigned warpsPerBlock = 4) {
    struct Call { static void run(void* a) { (*(F*)a)(); } };
#ifdef EMU_STD_THREADS
    // the TSan build: thread creation / join are what TSan understands as the launch's ordering with the host code around it
    {
        const unsigned total = gridX * gridY * warpsPerBlock;
        std::atomic<unsigned> next{0};
        auto worker = [&] {
            for (unsigned i = next.fetch_add(1, std::memory_order_relaxed); i < total; i = next.fetch_add(1, std::memory_order_relaxed)) {   // relaxed: no ordering between warps that TSan could take for synchronisation
                const unsigned wp = i % warpsPerBlock, b = i / warpsPerBlock;
                blockIdx.x = b % gridX; blockIdx.y = b / gridX; gridDim.x = gridX; gridDim.y = gridY; blockDim.x = 32 * warpsPerBlock;
                emuRunWarp(wp * 32, &Call::run, &kernelCall);
            }
        };
        std::thread pool[4];
        for (auto& t : pool) t = std::thread(worker);
        for (auto& t : pool) t.join();
        return;
    }
#endif
#pragma omp parallel for collapse(3) schedule(dynamic, 1)
    for (unsigned by = 0; by < gridY; by++)
        for (unsigned bx = 0; bx < gridX; bx++)
            for (unsigned wp = 0; wp < warpsPerBlock; wp++) {
                blockIdx.x = bx; blockIdx.y = by; gridDim.x = gridX; gridDim.y = gridY; blockDim.x = 32 * warpsPerBlock;
                emuRunWarp(wp * 32, &Call::run, &kernelCall);
            }
}
// a grid of blocks whose threads synchronise with __syncthreads (and use no warp intrinsic): the whole block is one set of fibers on one OS
// thread, so its __shared__ arrays are shared by exactly its threads
template <typename F> static inline void emuLaunchBlock(unsigned gridX, F kernelCall, unsigned threadsPerBlock = 256) {
    struct Call { static void run(void* a) { (*(F*)a)(); } };
#ifdef EMU_STD_THREADS
    {
        std::atomic<unsigned> next{0};
        auto worker = [&] {
            for (unsigned b = next.fetch_add(1, std::memory_order_relaxed); b < gridX; b = next.fetch_add(1, std::memory_order_relaxed)) {
                blockIdx.x = b; blockIdx.y = 0; gridDim.x = gridX; gridDim.y = 1; blockDim.x = threadsPerBlock;
                emuRunLanes(0, (int)threadsPerBlock, true, &Call::run, &kernelCall);
            }
        };
        std::thread pool[4];
        for (auto& t : pool) t = std::thread(worker);
        for (auto& t : pool) t.join();
        return;
    }
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (unsigned b = 0; b < gridX; b++) {
        blockIdx.x = b; blockIdx.y = 0; gridDim.x = gridX; gridDim.y = 1; blockDim.x = threadsPerBlock;
        emuRunLanes(0, (int)threadsPerBlock, true, &Call::run, &kernelCall);
    }
}
