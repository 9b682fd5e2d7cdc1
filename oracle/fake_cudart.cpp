/*
 * fake_cudart.cpp -- host-memory stand-ins for the few CUDA runtime entry points that the
 * reference's HOST code calls (cudaUtil.h:13-62, scene.cpp:435-509, sampler.h:157-170).
 * TEST INFRASTRUCTURE ONLY: linked into oracle/_ref/libref_harness.so instead of libcudart so that
 * Scene::buildDevData() / DevScene::create() run unmodified on a GPU-less machine; "device"
 * pointers are then ordinary heap pointers and DevScene's __device__ methods (plain functions
 * under g++) can be called on them.
 */
#include <cstdlib>
#include <cstring>

extern "C" {
typedef int cudaError_t;
int cudaMalloc(void** p, size_t n) { *p = n ? calloc(1, n) : nullptr; return 0; }
int cudaFree(void* p) { free(p); return 0; }
int cudaMemcpy(void* d, const void* s, size_t n, int) { if (n) memcpy(d, s, n); return 0; }
int cudaMemset(void* d, int v, size_t n) { if (n) memset(d, v, n); return 0; }
int cudaDeviceSynchronize(void) { return 0; }
int cudaGetLastError(void) { return 0; }
const char* cudaGetErrorString(int) { return "fake cudart"; }
}
