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
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicMax(unsigned* p, unsigned v) {
    unsigned o = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (o < v && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return o;
}
static inline long long clock64() { return 0; }
static inline void __syncthreads() {}
static inline void __trap() { abort(); }
// a "warp" of one lane: enough for the warp-level code to compile; the emulation never calls the packet walk
static inline int __any_sync(unsigned, int p) { return p; }
static inline int __all_sync(unsigned, int p) { return p; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
template <typename T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
static inline unsigned __reduce_min_sync(unsigned, unsigned v) { return v; }
static inline unsigned __reduce_max_sync(unsigned, unsigned v) { return v; }
static inline unsigned __reduce_or_sync(unsigned, unsigned v) { return v; }
static inline unsigned __reduce_add_sync(unsigned, unsigned v) { return v; }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline unsigned __match_any_sync(unsigned, unsigned) { return 1u; }
static inline unsigned __activemask() { return 1u; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
