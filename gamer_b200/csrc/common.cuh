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

// cudaFuncSetAttribute is per device: a call site keeps one flag per device ordinal (a process may drive several GPUs)
struct PerDeviceOnce {
    unsigned long long value[64] = {};
    // true when `want` exceeds what this call site has configured on the current device (and records it)
    bool need(unsigned long long want = 1) {
        int dev = 0;
        cudaGetDevice(&dev);
        unsigned long long& v = value[dev & 63];
        if (want <= v) return false;
        v = want;
        return true;
    }
};

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

// ---------------------------------------------------------------------------------------------------------
// Dropout (nn.Dropout on the residual branches / inside the experts, SDPA dropout_p on the attention probabilities:
// Qwen3Multi/model.py:139,217,235,241, Qwen3Moe/FFN.py:26).  Masks come from the counter-based Philox4x32-7 generator
// keyed on (seed, step offset) and indexed by (site, row, column), so the backward regenerates them instead of storing
// them.  Hidden-state sites use 16 random bits per element, the attention probabilities 8 (one Philox call covers 16
// keys); `scale` is the reciprocal of the realised keep probability, so E[dropout(x)] = x exactly.
// ---------------------------------------------------------------------------------------------------------
struct DropParams {
    uint32_t k0, k1;   // Philox key: seed low word, seed high word ^ step offset
    uint32_t site;     // which dropout call of the step (layer * 8 + kind)
    uint32_t thresh;   // drop when the element's random value < thresh (0 = dropout off)
    float scale;       // 1 / keep probability
    const uint32_t* off_dev;  // optional device word XOR-ed into k1 when the kernel starts (CUDA-graph replays)
};

static inline DropParams make_drop(const gamer_dropout_t* d, int bits) {
    DropParams r{0u, 0u, 0u, 0u, 1.0f, nullptr};
    if (d == nullptr || !(d->p > 0.0f)) return r;
    const uint32_t full = 1u << bits;
    uint32_t t = (uint32_t)((double)d->p * (double)full);
    if (t >= full) t = full - 1;
    r.k0 = (uint32_t)(d->seed & 0xffffffffull);
    r.k1 = (uint32_t)(d->seed >> 32) ^ d->offset;
    r.site = d->site;
    r.thresh = t;
    r.scale = (float)((double)full / (double)(full - t));
    r.off_dev = d->offset_dev;
    return r;
}

// kernels call this once: folds the device-side offset word into the key
__device__ __forceinline__ DropParams drop_resolve(DropParams d) {
    if (d.thresh != 0u && d.off_dev != nullptr) d.k1 ^= __ldg(d.off_dev);
    d.off_dev = nullptr;
    return d;
}

__device__ __forceinline__ uint4 philox4x32_7(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 7; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

constexpr uint32_t DROP_HIDDEN_TAG = 0x64726f70u;  // counter word 3 of the hidden-state sites (attention uses the site id)

// keep bits of the 8 columns [8*col8, 8*col8+8) of hidden row `row`: bit e set = element e is kept
__device__ __forceinline__ uint32_t drop_keep8(const DropParams& d, uint32_t row, uint32_t col8) {
    const uint4 r = philox4x32_7(col8, row, d.site, DROP_HIDDEN_TAG, d.k0, d.k1);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    uint32_t keep = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        keep |= ((w[q] & 0xffffu) >= d.thresh ? 1u : 0u) << (2 * q);
        keep |= ((w[q] >> 16) >= d.thresh ? 1u : 0u) << (2 * q + 1);
    }
    return keep;
}
// in-place dropout of 8 consecutive fp32 values
__device__ __forceinline__ void drop_apply8(const DropParams& d, uint32_t row, uint32_t col8, float* v) {
    const uint32_t keep = drop_keep8(d, row, col8);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = ((keep >> e) & 1u) ? v[e] * d.scale : 0.f;
}
// attention probabilities: random bytes of the 16 keys [16*jb, 16*jb+16) of query row i of (sequence, head) bh; key
// 16*jb+e <-> byte (e & 3) of word (e >> 2); kept iff byte >= thresh
__device__ __forceinline__ uint4 drop_attn16(const DropParams& d, uint32_t bh, uint32_t i, uint32_t jb) {
    return philox4x32_7(jb, i, bh, d.site, d.k0, d.k1);
}
