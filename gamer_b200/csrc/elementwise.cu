// Small fused elementwise kernels around the GEMMs: SwiGLU, the cross-attention output gate, row gather into the
// expert-permuted space, and the fused softmax-cross-entropy (loss + dlogits in one pass over the logits).
//
// Reference semantics: MyQwen3MoeMLP.forward  down(silu(gate x) * up x)  (SeqRec/models/generative/Qwen3Moe/FFN.py:25-27);
// o_proj(a) * silu(gating(h)) (Qwen3Multi/model.py:146-147); ForCausalLMLoss (transformers loss_utils.py:28-68).
#include "common.cuh"

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// flat vector index -> (row, first column): a shift and a mask when the row holds a power-of-two number of 16-byte
// vectors (every shipped width), instead of a 64-bit division per element
__device__ __forceinline__ int pow2_shift(int n) { return (n & (n - 1)) == 0 ? 31 - __clz(n) : -1; }
__device__ __forceinline__ void split_index(long long i, int vec_per_row, int vshift, long long& r, int& c) {
    if (vshift >= 0) {
        r = i >> vshift;
        c = (int)(i & (vec_per_row - 1)) * 8;
    } else {
        r = i / vec_per_row;
        c = (int)(i % vec_per_row) * 8;
    }
}

// gu: [R, 2*I] (gate | up), act: [R, I] = dropout(silu(gate) * up); the mask row is row_ids[r] (token row) when given
__global__ void swiglu_fwd_kernel(const bf16* __restrict__ gu, long long ld_gu, bf16* __restrict__ act, long long ld_act,
                                  long long R, int I, const int* __restrict__ row_ids, DropParams dp) {
    const int vec_per_row = I / 8;
    const int vshift = pow2_shift(vec_per_row);
    dp = drop_resolve(dp);
    const long long total = R * vec_per_row;
    const long long stride = (long long)gridDim.x * blockDim.x;
    // two vectors per iteration: their four 16-byte loads are issued before any arithmetic
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 2 * stride) {
        long long r[2];
        int c[2];
        bf16x8 gv[2], uv[2];
        int mrow[2];
        bool ok[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const long long i = i0 + t * stride;
            ok[t] = i < total;
            split_index(ok[t] ? i : 0, vec_per_row, vshift, r[t], c[t]);
            gv[t] = *reinterpret_cast<const bf16x8*>(gu + r[t] * ld_gu + c[t]);
            uv[t] = *reinterpret_cast<const bf16x8*>(gu + r[t] * ld_gu + I + c[t]);
            mrow[t] = (dp.thresh && row_ids) ? row_ids[r[t]] : (int)r[t];
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            if (!ok[t]) continue;
            float g[8], u[8], o[8];
            bf16x8_to_float(gv[t], g);
            bf16x8_to_float(uv[t], u);
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = g[k] * sigmoidf_(g[k]) * u[k];
            if (dp.thresh && mrow[t] >= 0) drop_apply8(dp, (uint32_t)mrow[t], (uint32_t)(c[t] >> 3), o);
            *reinterpret_cast<bf16x8*>(act + r[t] * ld_act + c[t]) = float_to_bf16x8(o);
        }
    }
}

__global__ void swiglu_bwd_kernel(const bf16* __restrict__ gu, long long ld_gu, const bf16* __restrict__ dact,
                                  long long ld_dact, bf16* __restrict__ dgu, long long ld_dgu, long long R, int I,
                                  const int* __restrict__ row_ids, DropParams dp) {
    const int vec_per_row = I / 8;
    const int vshift = pow2_shift(vec_per_row);
    dp = drop_resolve(dp);
    const long long total = R * vec_per_row;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r;
        int c;
        split_index(i, vec_per_row, vshift, r, c);
        float g[8], u[8], d[8], dg[8], du[8];
        bf16x8_to_float(*reinterpret_cast<const bf16x8*>(gu + r * ld_gu + c), g);
        bf16x8_to_float(*reinterpret_cast<const bf16x8*>(gu + r * ld_gu + I + c), u);
        bf16x8_to_float(*reinterpret_cast<const bf16x8*>(dact + r * ld_dact + c), d);
        if (dp.thresh) {
            const long long mr = row_ids ? (long long)row_ids[r] : r;
            if (mr >= 0) drop_apply8(dp, (uint32_t)mr, (uint32_t)(c >> 3), d);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float s = sigmoidf_(g[k]);
            du[k] = d[k] * g[k] * s;
            dg[k] = d[k] * u[k] * s * (1.0f + g[k] * (1.0f - s));
        }
        *reinterpret_cast<bf16x8*>(dgu + r * ld_dgu + c) = float_to_bf16x8(dg);
        *reinterpret_cast<bf16x8*>(dgu + r * ld_dgu + I + c) = float_to_bf16x8(du);
    }
}

// out = x + dropout(y * silu(g))   (all [R, W]; g has its own row stride: it lives inside the fused projection buffer)
__global__ void gate_residual_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ y,
                                         const bf16* __restrict__ g, long long ld_g, bf16* __restrict__ out, long long R,
                                         int W, DropParams dp) {
    const int vec_per_row = W / 8;
    const int vshift = pow2_shift(vec_per_row);
    dp = drop_resolve(dp);
    const long long total = R * vec_per_row;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r;
        int c;
        split_index(i, vec_per_row, vshift, r, c);
        float xf[8], yf[8], gf[8], o[8];
        bf16x8_to_float(*reinterpret_cast<const bf16x8*>(x + r * W + c), xf);
        bf16x8_to_float(*reinterpret_cast<const bf16x8*>(y + r * W + c), yf);
        bf16x8_to_float(*reinterpret_cast<const bf16x8*>(g + r * ld_g + c), gf);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = yf[k] * gf[k] * sigmoidf_(gf[k]);
        if (dp.thresh) drop_apply8(dp, (uint32_t)r, (uint32_t)(c >> 3), o);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] += xf[k];
        *reinterpret_cast<bf16x8*>(out + r * W + c) = float_to_bf16x8(o);
    }
}

// dz = dropout_mask * dout;  dy = dz * silu(g);  dg = dz * y * silu'(g)
__global__ void gate_residual_bwd_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ y,
                                         const bf16* __restrict__ g, long long ld_g, bf16* __restrict__ dy,
                                         bf16* __restrict__ dg, long long ld_dg, long long R, int W, DropParams dp) {
    const int vec_per_row = W / 8;
    const int vshift = pow2_shift(vec_per_row);
    dp = drop_resolve(dp);
    const long long total = R * vec_per_row;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r;
        int c;
        split_index(i, vec_per_row, vshift, r, c);
        float d[8], yf[8], gf[8], o1[8], o2[8];
        bf16x8_to_float(*reinterpret_cast<const bf16x8*>(dout + r * W + c), d);
        bf16x8_to_float(*reinterpret_cast<const bf16x8*>(y + r * W + c), yf);
        bf16x8_to_float(*reinterpret_cast<const bf16x8*>(g + r * ld_g + c), gf);
        if (dp.thresh) drop_apply8(dp, (uint32_t)r, (uint32_t)(c >> 3), d);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float s = sigmoidf_(gf[k]);
            o1[k] = d[k] * gf[k] * s;
            o2[k] = d[k] * yf[k] * s * (1.0f + gf[k] * (1.0f - s));
        }
        *reinterpret_cast<bf16x8*>(dy + r * W + c) = float_to_bf16x8(o1);
        *reinterpret_cast<bf16x8*>(dg + r * ld_dg + c) = float_to_bf16x8(o2);
    }
}

// dst[r, :] = rows[r] >= 0 ? dropout_mask[rows[r], :] * src[rows[r], :] : 0     (rows == nullptr: identity)
__global__ void gather_rows_kernel(const bf16* __restrict__ src, long long ld_src, const int* __restrict__ rows,
                                   const int* __restrict__ n_rows_dev, long long n_rows_max, bf16* __restrict__ dst,
                                   long long ld_dst, int W, DropParams dp) {
    const int vec_per_row = W / 8;
    const int vshift = pow2_shift(vec_per_row);
    dp = drop_resolve(dp);
    const long long n_rows = n_rows_dev ? min((long long)*n_rows_dev, n_rows_max) : n_rows_max;
    const long long total = n_rows * vec_per_row;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r;
        int c;
        split_index(i, vec_per_row, vshift, r, c);
        const long long sr = rows ? (long long)rows[r] : r;
        bf16x8 v;
        v.u[0] = v.u[1] = v.u[2] = v.u[3] = 0u;
        if (sr >= 0) {
            v = *reinterpret_cast<const bf16x8*>(src + sr * ld_src + c);
            if (dp.thresh) {
                float f[8];
                bf16x8_to_float(v, f);
                drop_apply8(dp, (uint32_t)sr, (uint32_t)(c >> 3), f);
                v = float_to_bf16x8(f);
            }
        }
        *reinterpret_cast<bf16x8*>(dst + r * ld_dst + c) = v;
    }
}

// fused softmax cross-entropy over fp32 logits rows: one warp per row.
//   loss_row[r] = lse - logit[label] (0 when label == ignore);  dlogits[r, :] = (softmax - onehot) * scale (bf16; columns
//   [V, ld_d) zero-filled so the buffer can feed TMA-tiled GEMMs).  scale = grad_scale * (*inv_norm).
// REG_COLS > 0: the row (V <= 32 * REG_COLS logits) is read from HBM once and kept in registers for the three steps (max,
// sum of exponentials, gradient); REG_COLS == 0: generic three-pass version for larger vocabularies.
template <int REG_COLS>
__global__ void ce_fwd_bwd_kernel(const float* __restrict__ logits, long long ld_l, const long long* __restrict__ labels,
                                  long long R, int V, int ignore_index, const float* __restrict__ inv_norm,
                                  float grad_scale, float* __restrict__ loss_row, bf16* __restrict__ dlogits,
                                  long long ld_d) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const float scale = grad_scale * (inv_norm ? *inv_norm : 1.0f);
    for (long long r = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); r < R; r += (long long)gridDim.x * wpb) {
        const float* lp = logits + r * ld_l;
        const long long lab = labels[r];
        const bool valid = lab != ignore_index;
        if constexpr (REG_COLS > 0) {
            float v[REG_COLS];
            float mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < REG_COLS; ++k) {
                const int c = lane + 32 * k;
                v[k] = (c < V) ? lp[c] : -INFINITY;
                mx = fmaxf(mx, v[k]);
            }
            mx = warp_max(mx);
            float se = 0.f, at_label = 0.f;
#pragma unroll
            for (int k = 0; k < REG_COLS; ++k) {
                const int c = lane + 32 * k;
                if (c == lab) at_label = v[k];
                v[k] = expf(v[k] - mx);          // exp(-inf) = 0 on the padding columns
                se += v[k];
            }
            se = warp_sum(se);
            at_label = warp_sum(at_label);
            if (lane == 0) loss_row[r] = valid ? (mx + logf(se) - at_label) : 0.f;
            if (dlogits != nullptr) {
                bf16* dp = dlogits + r * ld_d;
                const float inv = valid ? scale / se : 0.f;
#pragma unroll
                for (int k = 0; k < REG_COLS; ++k) {
                    const int c = lane + 32 * k;
                    if (c < ld_d) {
                        float g = 0.f;
                        if (c < V && valid) g = v[k] * inv - ((c == lab) ? scale : 0.f);
                        dp[c] = __float2bfloat16(g);
                    }
                }
                for (int c = lane + 32 * REG_COLS; c < ld_d; c += 32) dp[c] = __float2bfloat16(0.f);
            }
        } else {
            float mx = -INFINITY;
            for (int c = lane; c < V; c += 32) mx = fmaxf(mx, lp[c]);
            mx = warp_max(mx);
            float se = 0.f;
            for (int c = lane; c < V; c += 32) se += expf(lp[c] - mx);
            se = warp_sum(se);
            const float lse = mx + logf(se);
            if (lane == 0) loss_row[r] = valid ? (lse - lp[lab]) : 0.f;
            if (dlogits != nullptr) {
                bf16* dp = dlogits + r * ld_d;
                const float inv = valid ? scale / se : 0.f;
                for (int c = lane; c < ld_d; c += 32) {
                    float v = 0.f;
                    if (c < V && valid) v = expf(lp[c] - mx) * inv - ((c == lab) ? scale : 0.f);
                    dp[c] = __float2bfloat16(v);
                }
            }
        }
    }
}

inline int grid_for(long long total, int threads) {
    long long b = (total + threads - 1) / threads;
    return (int)(b < 148 * 16 ? (b > 0 ? b : 1) : 148 * 16);
}

// dst[r, 0..W) = 0 for every row r with rows[r] < 0 (the padding rows of the expert-permuted token space): a warp
// looks at 32 map entries per iteration (one coalesced load) and clears the unmapped ones, 16 bytes per lane.
__global__ void zero_unmapped_rows_kernel(bf16* __restrict__ dst, long long ld, const int* __restrict__ rows,
                                          long long n_rows, int W) {
    const int lane = threadIdx.x & 31;
    const long long w0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    for (long long r0 = w0 * 32; r0 < n_rows; r0 += (long long)gridDim.x * (blockDim.x >> 5) * 32) {
        const bool unmapped = (r0 + lane < n_rows) && rows[r0 + lane] < 0;
        for (unsigned m = __ballot_sync(0xffffffffu, unmapped); m; m &= m - 1) {
            const long long r = r0 + (__ffs(m) - 1);
            for (int c = lane * 8; c < W; c += 256) *reinterpret_cast<uint4*>(dst + r * ld + c) = make_uint4(0, 0, 0, 0);
        }
    }
}

}  // namespace

extern "C" int gamer_swiglu_fwd(const void* gu, long long ld_gu, void* act, long long ld_act, long long R, int I,
                                const int* row_ids, const gamer_dropout_t* drop, cudaStream_t stream) {
    GAMER_REQUIRE(I % 8 == 0, "intermediate size must be a multiple of 8");
    if (R == 0) return 0;
    swiglu_fwd_kernel<<<grid_for(R * (I / 8), 256), 256, 0, stream>>>(reinterpret_cast<const bf16*>(gu), ld_gu,
                                                                      reinterpret_cast<bf16*>(act), ld_act, R, I,
                                                                      row_ids, make_drop(drop, 16));
    GAMER_LAUNCH_CHECK();
    return 0;
}

extern "C" int gamer_swiglu_bwd(const void* gu, long long ld_gu, const void* dact, long long ld_dact, void* dgu,
                                long long ld_dgu, long long R, int I, const int* row_ids,
                                const gamer_dropout_t* drop, cudaStream_t stream) {
    GAMER_REQUIRE(I % 8 == 0, "intermediate size must be a multiple of 8");
    if (R == 0) return 0;
    swiglu_bwd_kernel<<<grid_for(R * (I / 8), 256), 256, 0, stream>>>(
        reinterpret_cast<const bf16*>(gu), ld_gu, reinterpret_cast<const bf16*>(dact), ld_dact,
        reinterpret_cast<bf16*>(dgu), ld_dgu, R, I, row_ids, make_drop(drop, 16));
    GAMER_LAUNCH_CHECK();
    return 0;
}

extern "C" int gamer_gate_residual_fwd(const void* x, const void* y, const void* g, long long ld_g, void* out,
                                       long long R, int W, const gamer_dropout_t* drop, cudaStream_t stream) {
    GAMER_REQUIRE(W % 8 == 0, "width must be a multiple of 8");
    if (R == 0) return 0;
    gate_residual_fwd_kernel<<<grid_for(R * (W / 8), 256), 256, 0, stream>>>(
        reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(y), reinterpret_cast<const bf16*>(g), ld_g,
        reinterpret_cast<bf16*>(out), R, W, make_drop(drop, 16));
    GAMER_LAUNCH_CHECK();
    return 0;
}

extern "C" int gamer_gate_residual_bwd(const void* dout, const void* y, const void* g, long long ld_g, void* dy,
                                       void* dg, long long ld_dg, long long R, int W, const gamer_dropout_t* drop,
                                       cudaStream_t stream) {
    GAMER_REQUIRE(W % 8 == 0, "width must be a multiple of 8");
    if (R == 0) return 0;
    gate_residual_bwd_kernel<<<grid_for(R * (W / 8), 256), 256, 0, stream>>>(
        reinterpret_cast<const bf16*>(dout), reinterpret_cast<const bf16*>(y), reinterpret_cast<const bf16*>(g), ld_g,
        reinterpret_cast<bf16*>(dy), reinterpret_cast<bf16*>(dg), ld_dg, R, W, make_drop(drop, 16));
    GAMER_LAUNCH_CHECK();
    return 0;
}

extern "C" int gamer_gather_rows(const void* src, long long ld_src, const int* rows, const int* n_rows_dev,
                                 long long n_rows_max, void* dst, long long ld_dst, int W, const gamer_dropout_t* drop,
                                 cudaStream_t stream) {
    GAMER_REQUIRE(W % 8 == 0, "width must be a multiple of 8");
    if (n_rows_max == 0) return 0;
    gather_rows_kernel<<<grid_for(n_rows_max * (W / 8), 256), 256, 0, stream>>>(
        reinterpret_cast<const bf16*>(src), ld_src, rows, n_rows_dev, n_rows_max, reinterpret_cast<bf16*>(dst), ld_dst, W,
        make_drop(drop, 16));
    GAMER_LAUNCH_CHECK();
    return 0;
}

extern "C" int gamer_zero_unmapped_rows(void* dst, long long ld_dst, const int* rows, long long n_rows, int W,
                                        cudaStream_t stream) {
    GAMER_REQUIRE(W % 8 == 0 && ld_dst % 8 == 0, "width and row stride must be multiples of 8");
    if (n_rows == 0) return 0;
    const long long blocks = (n_rows + 255) / 256;       // 8 warps x 32 map entries per block and iteration
    zero_unmapped_rows_kernel<<<(int)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, stream>>>(
        reinterpret_cast<bf16*>(dst), ld_dst, rows, n_rows, W);
    GAMER_LAUNCH_CHECK();
    return 0;
}

extern "C" int gamer_dropout_apply(const void* in, void* out, long long R, int W, const gamer_dropout_t* drop,
                                   cudaStream_t stream) {
    GAMER_REQUIRE(W % 8 == 0, "width must be a multiple of 8");
    if (R == 0) return 0;
    gather_rows_kernel<<<grid_for(R * (W / 8), 256), 256, 0, stream>>>(reinterpret_cast<const bf16*>(in), W, nullptr,
                                                                       nullptr, R, reinterpret_cast<bf16*>(out), W, W,
                                                                       make_drop(drop, 16));
    GAMER_LAUNCH_CHECK();
    return 0;
}

extern "C" int gamer_ce_fwd_bwd(const float* logits, long long ld_l, const long long* labels, long long R, int V,
                                int ignore_index, const float* inv_norm, float grad_scale, float* loss_row,
                                void* dlogits, long long ld_d, cudaStream_t stream) {
    if (R == 0) return 0;
    const int wpb = 8;
    const int grid = (int)((R + wpb - 1) / wpb < 148 * 8 ? (R + wpb - 1) / wpb : 148 * 8);
    if (V <= 32 * 36)
        ce_fwd_bwd_kernel<36><<<grid, wpb * 32, 0, stream>>>(logits, ld_l, labels, R, V, ignore_index, inv_norm, grad_scale,
                                                             loss_row, reinterpret_cast<bf16*>(dlogits), ld_d);
    else
        ce_fwd_bwd_kernel<0><<<grid, wpb * 32, 0, stream>>>(logits, ld_l, labels, R, V, ignore_index, inv_norm, grad_scale,
                                                            loss_row, reinterpret_cast<bf16*>(dlogits), ld_d);
    GAMER_LAUNCH_CHECK();
    return 0;
}
