// vecmath.h -- fp32 vector helpers shared by host (g++) and device (nvcc) code.
//
// Every function fixes the ORDER of its fp32 operations to the one the reference's glm 0.9.6.3
// expressions have (cited per function), and the library is compiled with -fmad=false /
// -ffp-contract=off, so host and device produce the same bits for the same inputs.
#pragma once

#include <math.h>
#include <stdint.h>
#include <float.h>

#if defined(__CUDACC__)
#define RS_HD __host__ __device__ __forceinline__
#define RS_D __device__ __forceinline__
#else
#define RS_HD inline
#define RS_D inline
#endif

namespace rs {

struct f3 {
    float x, y, z;
};

RS_HD f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
RS_HD f3 mk3(float a) { return mk3(a, a, a); }
RS_HD f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
RS_HD f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
RS_HD f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
RS_HD f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
RS_HD f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
RS_HD f3 operator/(f3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }     // glm type_vec3.inl:707: true division
RS_HD f3 operator/(f3 a, f3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
// glm::min / glm::max are "x < y ? x : y" / "x > y ? x : y" (func_common.inl:409,430): NaN in y wins
RS_HD float gmin(float x, float y) { return x < y ? x : y; }
RS_HD float gmax(float x, float y) { return x > y ? x : y; }
RS_HD f3 gmin(f3 a, f3 b) { return mk3(gmin(a.x, b.x), gmin(a.y, b.y), gmin(a.z, b.z)); }
RS_HD f3 gmax(f3 a, f3 b) { return mk3(gmax(a.x, b.x), gmax(a.y, b.y), gmax(a.z, b.z)); }
RS_HD float dot(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }      // func_geometric.inl:64-72
RS_HD f3 cross(f3 x, f3 y) {                                                     // func_geometric.inl:134
    return mk3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
RS_HD float length(f3 v) { return sqrtf(dot(v, v)); }                            // :95
RS_HD f3 normalize(f3 v) { return v * (1.f / sqrtf(dot(v, v))); }                // :154, func_exponential.inl:150
RS_HD float mixf(float x, float y, float a) { return x + a * (y - x); }          // func_common.inl:97
RS_HD f3 mix(f3 x, f3 y, float a) { return x + (y - x) * a; }
RS_HD f3 mix(f3 x, f3 y, f3 a) { return x + a * (y - x); }
RS_HD float comp(f3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

#define RS_PI 3.1415926535897932384626422832795028841971f                        /* mathUtil.h:10 */
#define RS_GLM_PI ((float)3.14159265358979323846264338327950288)                 /* glm::pi<float>() */

RS_HD float radians(float deg) { return deg * (float)0.01745329251994329576923690768489; }  // func_trigonometric.inl:45
RS_HD float luminance(f3 c) { return dot(c, mk3(.2126f, .7152f, .0722f)); }      // mathUtil.h:119
RS_HD bool isNanOrInf(float x) { return isnan(x) || isinf(x); }                  // mathUtil.h:56
RS_HD bool hasNanOrInf(f3 v) { return isNanOrInf(v.x) || isNanOrInf(v.y) || isNanOrInf(v.z); }
RS_HD float satDot(f3 a, f3 b) { return gmax(dot(a, b), 0.f); }                  // mathUtil.h:64
RS_HD float absDot(f3 a, f3 b) { return fabsf(dot(a, b)); }                      // mathUtil.h:68
RS_HD float pow5(float x) { float x2 = x * x; return x2 * x2 * x; }              // mathUtil.h:72
RS_HD float triangleArea(f3 v0, f3 v1, f3 v2) { return length(cross(v1 - v0, v2 - v0)) * .5f; }   // mathUtil.h:86
RS_HD f3 triangleNormal(f3 v0, f3 v1, f3 v2) { return normalize(cross(v1 - v0, v2 - v0)); }       // mathUtil.h:90

// mathUtil.h:190-198
RS_HD uint32_t utilhash(uint32_t a) {
    a = (a + 0x7ed55d16) + (a << 12);
    a = (a ^ 0xc761c23c) ^ (a >> 19);
    a = (a + 0x165667b1) + (a << 5);
    a = (a + 0xd3a2646c) ^ (a << 9);
    a = (a + 0xfd7046c5) + (a << 3);
    a = (a ^ 0xb55a4f09) ^ (a >> 16);
    return a;
}

// thrust::minstd_rand seeded as in sampler.h:39-44, drawn through uniform_real_distribution<float>(0,1)
// (thrust/random/detail/linear_congruential_engine.inl:45-63, uniform_real_distribution.inl:63-74).
struct Rng {
    uint32_t x;
    RS_HD void seed(int looper, int index) {
        uint32_t h = utilhash((1u << 31) | (uint32_t)looper) ^ utilhash((uint32_t)index);
        x = h % 2147483647u;
        if (x == 0) x = 1;
    }
    RS_HD float next() {
        // x * 48271 mod (2^31 - 1) via the Mersenne fold (exactly Schrage's result)
        uint64_t p = (uint64_t)x * 48271ull;
        uint32_t lo = (uint32_t)(p & 0x7fffffffu), hi = (uint32_t)(p >> 31);
        uint32_t s = lo + hi;
        if (s >= 2147483647u) s -= 2147483647u;
        x = s;
        return (float)(x - 1u) * 4.656612873077392578125e-10f;   // / 2^31 (exact power of two)
    }
};

}  // namespace rs
