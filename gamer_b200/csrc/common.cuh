// Shared device/host helpers for the gamer_b200 sm_100a kernels.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "gamer_b200.h"  // the C-ABI: definitions below must match these declarations

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

// ---------------------------------------------------------------------------------------------------------
// error plumbing for the C-ABI (thread-local last error string)
// ---------------------------------------------------------------------------------------------------------
void gamer_set_error(const char* fmt, ...);

#define GAMER_CHECK_CUDA(expr)                                                                   \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            gamer_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return -2;                                                                           \
        }                                                                                        \
    } while (0)

#define GAMER_REQUIRE(cond, ...)                 \
    do {                                         \
        if (!(cond)) {                           \
            gamer_set_error(__VA_ARGS__);        \
            return -1;                           \
        }                                        \
    } while (0)

#define GAMER_LAUNCH_CHECK() GAMER_CHECK_CUDA(cudaGetLastError())

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// mask kinds (include/gamer_b200.h: GAMER_MASK_*)
enum { MASK_CAUSAL = 0, MASK_MULTI_CROSS = 1, MASK_SESSION = 2, MASK_SESSION_CROSS = 3 };

// ---------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    bf162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
    bf162 t = *reinterpret_cast<bf162*>(&u);
    return __bfloat1622float2(t);
}

// 16-byte vector of 8 bf16
struct __align__(16) bf16x8 {
    uint32_t u[4];
};
__device__ __forceinline__ void bf16x8_to_float(const bf16x8& v, float* f) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 t = unpack_bf16(v.u[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ bf16x8 float_to_bf16x8(const float* f) {
    bf16x8 v;
#pragma unroll
    for (int i = 0; i < 4; ++i) v.u[i] = pack_bf16(f[2 * i], f[2 * i + 1]);
    return v;
}

// Philox4x32-10 counter RNG (dropout masks): deterministic in (seed, offset, index)
__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                            uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
