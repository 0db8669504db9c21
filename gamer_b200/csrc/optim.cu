// Fused AdamW over the flat fp32 master-weight buffer (torch `adamw_torch` as configured at
// SeqRec/tasks/train_SMB_decoder.py:396-428, plus HF Trainer's clip_grad_norm_(max_grad_norm)).  One pass reads p, g, m, v and
// writes p, m, v and the bf16 operand copy the GEMMs consume, so no separate cast pass is needed.
#include "common.cuh"

namespace {

__global__ void sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
    float acc = 0.f;
    const long long n4 = n / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(g)[i];
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (blockIdx.x == 0)
        for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) acc += g[i] * g[i];
    acc = warp_sum(acc);
    __shared__ float part[32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(out, v);
    }
}

// hp (device): [0] lr  [1] 1-beta1^t  [2] 1-beta2^t
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, const unsigned char* __restrict__ decay_mask, bf16* __restrict__ p_bf16,
                             long long n, const float* __restrict__ hp, float beta1, float beta2, float eps,
                             float weight_decay, const float* __restrict__ gnorm_sq, float max_grad_norm,
                             float grad_scale) {
    const float lr = hp[0], bc1 = hp[1], bc2 = hp[2];
    float gs = grad_scale;
    if (gnorm_sq != nullptr && max_grad_norm > 0.f) {
        const float norm = sqrtf(*gnorm_sq) * grad_scale;
        gs *= fminf(1.0f, max_grad_norm / (norm + 1e-6f));
    }
    const float step = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i] * gs;
        float pi = p[i];
        if (decay_mask[i]) pi *= (1.0f - lr * weight_decay);
        const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
        pi -= step * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
        p[i] = pi;
        m[i] = mi;
        v[i] = vi;
        if (p_bf16 != nullptr) p_bf16[i] = __float2bfloat16(pi);
    }
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = __float2bfloat16(src[i]);
}

// Batched out-of-place transpose of bf16 matrices (the dgrad copies of every weight after an optimizer step): block =
// one 32x32 tile of one matrix, found by binary search over the matrices' tile offsets.
// desc[i] = {src, dst, rows, cols, ld_src, ld_dst} (int64 each): dst[c * ld_dst + r] = src[r * ld_src + c].
__global__ void __launch_bounds__(256) transpose_batch_kernel(const long long* __restrict__ desc,
                                                              const int* __restrict__ tile_start, int n_mats) {
    __shared__ unsigned short tile[32][33];
    int lo = 0, hi = n_mats;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (tile_start[mid] <= (int)blockIdx.x) lo = mid; else hi = mid;
    }
    const long long* d = desc + 6 * lo;
    const unsigned short* src = reinterpret_cast<const unsigned short*>(d[0]);
    unsigned short* dst = reinterpret_cast<unsigned short*>(d[1]);
    const int rows = (int)d[2], cols = (int)d[3];
    const long long ld_src = d[4], ld_dst = d[5];
    const int t = (int)blockIdx.x - tile_start[lo];
    const int tiles_x = (cols + 31) >> 5;
    const int r0 = (t / tiles_x) << 5, c0 = (t % tiles_x) << 5;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty + 8 * i, c = c0 + tx;
        if (r < rows && c < cols) tile[ty + 8 * i][tx] = src[(long long)r * ld_src + c];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i, r = r0 + tx;
        if (r < rows && c < cols) dst[(long long)c * ld_dst + r] = tile[tx][ty + 8 * i];
    }
}

}  // namespace

extern "C" int gamer_transpose_bf16_batch(const long long* desc, const int* tile_start, int n_mats, int total_tiles,
                                          cudaStream_t stream) {
    if (n_mats == 0 || total_tiles == 0) return 0;
    GAMER_REQUIRE(n_mats > 0 && total_tiles > 0, "bad transpose batch (%d matrices, %d tiles)", n_mats, total_tiles);
    transpose_batch_kernel<<<total_tiles, 256, 0, stream>>>(desc, tile_start, n_mats);
    GAMER_LAUNCH_CHECK();
    return 0;
}

extern "C" int gamer_sumsq_accumulate(const float* g, long long n, float* out, cudaStream_t stream) {
    if (n == 0) return 0;
    GAMER_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0, "gradient buffer must be 16-byte aligned");
    sumsq_kernel<<<148 * 4, 256, 0, stream>>>(g, n, out);
    GAMER_LAUNCH_CHECK();
    return 0;
}

extern "C" int gamer_adamw_step(float* p, const float* g, float* m, float* v, const unsigned char* decay_mask, void* p_bf16,
                                long long n, const float* hp, float beta1, float beta2, float eps, float weight_decay,
                                const float* gnorm_sq, float max_grad_norm, float grad_scale, cudaStream_t stream) {
    if (n == 0) return 0;
    adamw_kernel<<<148 * 8, 256, 0, stream>>>(p, g, m, v, decay_mask, reinterpret_cast<bf16*>(p_bf16), n, hp, beta1, beta2,
                                              eps, weight_decay, gnorm_sq, max_grad_norm, grad_scale);
    GAMER_LAUNCH_CHECK();
    return 0;
}

extern "C" int gamer_cast_f32_bf16(const float* src, void* dst, long long n, cudaStream_t stream) {
    if (n == 0) return 0;
    cast_bf16_kernel<<<148 * 8, 256, 0, stream>>>(src, reinterpret_cast<bf16*>(dst), n);
    GAMER_LAUNCH_CHECK();
    return 0;
}
