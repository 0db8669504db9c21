// K2/K3: RMSNorm (hidden rows), and the per-head q/k RMSNorm + behaviour-embedding add + RoPE block.
//
// Reference semantics: Qwen3RMSNorm  w * (x * rsqrt(mean(x^2) + eps))  with fp32 statistics
// (transformers modeling_qwen3_moe.py:290-308, used at SeqRec/models/generative/Qwen3Multi/model.py:50-51,165-176,284);
// q/k/v behaviour embeddings added before the head norm (Qwen3Multi/model.py:88-95); half-split RoPE
// (modeling_qwen3_moe.py:56-86) with cos/sin tables computed by the host exactly as Qwen3RotaryEmbedding does.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------------------
// rmsnorm over H = 256 columns: one warp per row, 8 columns (16 B) per lane.
//   out row = row_map ? row_map[m] : m, row stride ld_out; optional 64-wide behaviour-embedding concat after col H.
// ------------------------------------------------------------------------------------------------------------
constexpr int H256 = 256;

constexpr int NR = 2;  // rows per warp iteration: their loads are issued back to back (twice the bytes in flight)

__global__ void rmsnorm_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ w, float eps, long long M,
                                   bf16* __restrict__ out, long long ld_out, const int* __restrict__ row_map,
                                   const bf16* __restrict__ cat_table, const int* __restrict__ cat_idx, int cat_dim,
                                   float* __restrict__ rstd_out) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    float wv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) wv[i] = w[lane * 8 + i];
    const long long stride = (long long)gridDim.x * wpb;
    for (long long m0 = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); m0 < M; m0 += stride * NR) {
        bf16x8 raw[NR];
        long long orow[NR];
        int cidx[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const long long m = m0 + r * stride;     // warp-uniform
            if (m < M) {
                raw[r] = *reinterpret_cast<const bf16x8*>(x + m * H256 + lane * 8);
                orow[r] = row_map ? (long long)row_map[m] : m;
                cidx[r] = (cat_table != nullptr) ? cat_idx[m] : 0;
            }
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const long long m = m0 + r * stride;
            if (m >= M) continue;
            float f[8];
            bf16x8_to_float(raw[r], f);
            float ss = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) ss += f[i] * f[i];
            ss = warp_sum(ss);
            const float rstd = rsqrtf(ss * (1.0f / H256) + eps);
            if (lane == 0 && rstd_out != nullptr) rstd_out[m] = rstd;
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = wv[i] * (f[i] * rstd);
            bf16* op = out + orow[r] * ld_out;
            *reinterpret_cast<bf16x8*>(op + lane * 8) = float_to_bf16x8(f);
            if (cat_table != nullptr && lane < cat_dim / 8) {
                const bf16x8 e = *reinterpret_cast<const bf16x8*>(cat_table + (long long)cidx[r] * cat_dim + lane * 8);
                *reinterpret_cast<bf16x8*>(op + H256 + lane * 8) = e;
            }
        }
    }
}

// backward.  dh rows come from (row_map ? row_map[m] : m) with stride ld_dh.
//   dx[m] = (dres ? dres[m] : 0) + rstd*g - x*rstd^3*mean(g.x),  g = w*dh
//   dw   += sum_m dh*x*rstd        (block partials -> fp32 atomics)
//   dcat[cat_idx[m]] += dh[m, H:H+cat_dim]
template <bool HAS_CAT>
__global__ void __launch_bounds__(256, HAS_CAT ? 2 : 3)
rmsnorm_bwd_kernel(const bf16* __restrict__ x, const float* __restrict__ w,
                                   const float* __restrict__ rstd_in, float eps, long long M,
                                   const bf16* __restrict__ dh, long long ld_dh, const int* __restrict__ row_map,
                                   const bf16* __restrict__ dres, bf16* __restrict__ dx, float* __restrict__ dw,
                                   const int* __restrict__ cat_idx, int cat_dim, int cat_rows,
                                   float* __restrict__ dcat) {
    extern __shared__ float sh[];  // [H256] dw partials, then [cat_rows*cat_dim]
    float* sdw = sh;
    float* scat = sh + H256;
    for (int i = threadIdx.x; i < H256 + cat_rows * cat_dim; i += blockDim.x) sh[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    float wv[8], dwacc[8], catacc[HAS_CAT ? 4 : 1][8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        wv[i] = w[lane * 8 + i];
        dwacc[i] = 0.f;
#pragma unroll
        for (int q = 0; q < (HAS_CAT ? 4 : 1); ++q) catacc[q][i] = 0.f;
    }
    const long long stride = (long long)gridDim.x * wpb;
    for (long long m0 = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); m0 < M; m0 += stride * NR) {
        bf16x8 xr[NR], dr[NR], rr[NR], cr[NR];
        float rs[NR];
        int ci[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) {                 // all loads of the NR rows first
            const long long m = m0 + r * stride;        // warp-uniform
            if (m < M) {
                xr[r] = *reinterpret_cast<const bf16x8*>(x + m * H256 + lane * 8);
                const long long srow = row_map ? (long long)row_map[m] : m;
                const bf16* dp = dh + srow * ld_dh;
                dr[r] = *reinterpret_cast<const bf16x8*>(dp + lane * 8);
                if (dres != nullptr) rr[r] = *reinterpret_cast<const bf16x8*>(dres + m * H256 + lane * 8);
                rs[r] = (rstd_in != nullptr) ? rstd_in[m] : 0.f;
                if (HAS_CAT && lane < cat_dim / 8) {
                    cr[r] = *reinterpret_cast<const bf16x8*>(dp + H256 + lane * 8);
                    ci[r] = cat_idx[m];
                }
            }
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const long long m = m0 + r * stride;
            if (m >= M) continue;
            float xf[8], df[8];
            bf16x8_to_float(xr[r], xf);
            bf16x8_to_float(dr[r], df);
            float rstd = rs[r];
            if (rstd_in == nullptr) {
                float ss = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) ss += xf[i] * xf[i];
                rstd = rsqrtf(warp_sum(ss) * (1.0f / H256) + eps);
            }
            float dot = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                dwacc[i] += df[i] * xf[i] * rstd;
                df[i] *= wv[i];
                dot += df[i] * xf[i];
            }
            dot = warp_sum(dot) * (1.0f / H256) * rstd * rstd * rstd;
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = rstd * df[i] - xf[i] * dot;
            if (dres != nullptr) {
                float rf[8];
                bf16x8_to_float(rr[r], rf);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] += rf[i];
            }
            *reinterpret_cast<bf16x8*>(dx + m * H256 + lane * 8) = float_to_bf16x8(o);
            if (HAS_CAT && lane < cat_dim / 8) {
                float cf[8];
                bf16x8_to_float(cr[r], cf);
                const int rw = ci[r];
                if (rw < 4) {                      // register partials for the (<= 4) behaviour rows, flushed once below
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (q == rw) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) catacc[q][i] += cf[i];
                        }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) atomicAdd(&scat[rw * cat_dim + lane * 8 + i], cf[i]);
                }
            }
        }
    }
    if (HAS_CAT && lane < cat_dim / 8) {
#pragma unroll
        for (int q = 0; q < (HAS_CAT ? 4 : 1); ++q)
            if (q < cat_rows) {
#pragma unroll
                for (int i = 0; i < 8; ++i) atomicAdd(&scat[q * cat_dim + lane * 8 + i], catacc[q][i]);
            }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(&sdw[lane * 8 + i], dwacc[i]);
    __syncthreads();
    for (int i = threadIdx.x; i < H256; i += blockDim.x) atomicAdd(&dw[i], sdw[i]);
    if (HAS_CAT)
        for (int i = threadIdx.x; i < cat_rows * cat_dim; i += blockDim.x)
            if (scat[i] != 0.f) atomicAdd(&dcat[i], scat[i]);
}

// ------------------------------------------------------------------------------------------------------------
// per-head block on the fused projection buffer [M, ld]: columns [0,384) q (6 heads), [384,576) k (3), [576,768) v (3).
// 8 lanes per 64-wide head (16 B each): lanes j<4 hold the first half, j>=4 the second (RoPE partner = lane ^ 4).
//   q,k: u = raw (+ emb[act]);  n = wn * u * rsqrt(mean(u^2)+eps);  out = n*cos + rotate_half(n)*sin
//   v  : out = raw (+ emb[act])
// ------------------------------------------------------------------------------------------------------------
constexpr int HD = 64;
constexpr int PF = 4;     // rows in flight per thread in the head backward kernel

struct HeadArgs {
    long long M;
    int L;                 // tokens per sequence (position = m % L when pos_ids == nullptr)
    int n_q, n_kv;
    const int* pos_ids;    // [M] RoPE positions or nullptr
    int pos0;              // added to m % L when pos_ids == nullptr (decode steps)
    int n_pos;             // rows of the cos / sin tables: positions are clamped to [0, n_pos)
    const float* cos_tab;  // [n_pos, 32]
    const float* sin_tab;
    const float* qn_w;     // [64]
    const float* kn_w;
    const bf16* q_emb;     // [n_beh+1, n_q*64] or nullptr (self attention)
    const bf16* k_emb;
    const bf16* v_emb;
    const int* act_idx;    // [M]
    float eps;
};

// Thread block = (8 lanes, n_heads, Z tokens): a thread keeps one head for the whole kernel and walks tokens with stride
// gridDim.x * Z.  Lane j of a head holds columns [4j, 4j+4) of the first half and [32+4j, 32+4j+4) of the second half, so
// the RoPE partner of every element is in the same thread (no shuffles) and cos/sin are one float4 each.
struct HeadRow {
    float f[8];  // [0,4): first-half columns, [4,8): second-half columns
};
struct HeadRaw {
    uint2 lo, hi;
};
__device__ __forceinline__ HeadRaw load_head_raw(const bf16* p, int sub) {
    HeadRaw r;
    r.lo = *reinterpret_cast<const uint2*>(p + 4 * sub);
    r.hi = *reinterpret_cast<const uint2*>(p + 32 + 4 * sub);
    return r;
}
__device__ __forceinline__ HeadRow head_row_from_raw(const HeadRaw& w) {
    const uint2 lo = w.lo, hi = w.hi;
    HeadRow r;
    float2 t;
    t = unpack_bf16(lo.x); r.f[0] = t.x; r.f[1] = t.y;
    t = unpack_bf16(lo.y); r.f[2] = t.x; r.f[3] = t.y;
    t = unpack_bf16(hi.x); r.f[4] = t.x; r.f[5] = t.y;
    t = unpack_bf16(hi.y); r.f[6] = t.x; r.f[7] = t.y;
    return r;
}
__device__ __forceinline__ HeadRow load_head_row(const bf16* p, int sub) {
    const uint2 lo = *reinterpret_cast<const uint2*>(p + 4 * sub);
    const uint2 hi = *reinterpret_cast<const uint2*>(p + 32 + 4 * sub);
    HeadRow r;
    float2 t;
    t = unpack_bf16(lo.x); r.f[0] = t.x; r.f[1] = t.y;
    t = unpack_bf16(lo.y); r.f[2] = t.x; r.f[3] = t.y;
    t = unpack_bf16(hi.x); r.f[4] = t.x; r.f[5] = t.y;
    t = unpack_bf16(hi.y); r.f[6] = t.x; r.f[7] = t.y;
    return r;
}
__device__ __forceinline__ void store_head_row(bf16* p, int sub, const float* f) {
    *reinterpret_cast<uint2*>(p + 4 * sub) = make_uint2(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]));
    *reinterpret_cast<uint2*>(p + 32 + 4 * sub) = make_uint2(pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}
__device__ __forceinline__ float group8_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}

// Token ranges: block b owns a contiguous chunk of tokens and each of its Z slices a contiguous sub-chunk, which it
// walks token by token — row pointers advance by one row stride, the RoPE position by one (mod L), and the behaviour
// (action) index of a thread changes only every few tokens (items are 5 tokens long), so no 64-bit multiply, division or
// per-token select survives in the loop.  Every slice runs the same trip count (`live` masks the tail) because a warp may
// straddle slices when the head count is not a multiple of four.
struct TokenRange {
    long long begin, end;  // this slice's tokens
    int iters;             // uniform trip count
};
__device__ __forceinline__ TokenRange slice_range(long long M, int Z, int z) {
    const long long per_block = (M + gridDim.x - 1) / gridDim.x;
    const long long per_z = (per_block + Z - 1) / Z;
    const long long b0 = (long long)blockIdx.x * per_block;
    TokenRange r;
    r.begin = b0 + (long long)z * per_z;
    long long e = r.begin + per_z;
    if (e > b0 + per_block) e = b0 + per_block;
    if (e > M) e = M;
    r.end = e;
    r.iters = (int)per_z;
    return r;
}

template <bool HAS_EMB>
__global__ void __launch_bounds__(384)
qk_norm_rope_fwd_kernel(HeadArgs a, const bf16* __restrict__ raw, long long ld_raw, bf16* __restrict__ out,
                        long long ld_out) {
    const int sub = threadIdx.x, h = threadIdx.y, Z = blockDim.z;
    const bool is_q = h < a.n_q, is_k = !is_q && h < a.n_q + a.n_kv;
    const bf16* emb = is_q ? a.q_emb : (is_k ? a.k_emb : a.v_emb);
    const int width = is_q ? a.n_q * HD : a.n_kv * HD;
    const int hc = (is_q ? h : (is_k ? h - a.n_q : h - a.n_q - a.n_kv)) * HD;
    const float* wn = is_q ? a.qn_w : a.kn_w;
    const float4 w_lo = *reinterpret_cast<const float4*>(wn + 4 * sub);
    const float4 w_hi = *reinterpret_cast<const float4*>(wn + 32 + 4 * sub);
    const float wv[8] = {w_lo.x, w_lo.y, w_lo.z, w_lo.w, w_hi.x, w_hi.y, w_hi.z, w_hi.w};
    const TokenRange tr = slice_range(a.M, Z, threadIdx.z);
    const long long mlast = a.M - 1;
    long long m = tr.begin;
    const long long mc0 = m < a.M ? m : mlast;
    const bf16* rp = raw + mc0 * ld_raw + h * HD;      // row of the NEXT load (clamped to the last token)
    bf16* op = out + mc0 * ld_out + h * HD;
    int pos = (int)(mc0 % a.L);
    HeadRaw nxt = load_head_raw(rp, sub);
    for (int it = 0; it < tr.iters; ++it, ++m) {
        const bool live = m < tr.end;
        HeadRow u = head_row_from_raw(nxt);
        if (m + 1 <= mlast) rp += ld_raw;              // next token's row: in flight while this one is processed
        nxt = load_head_raw(rp, sub);
        const long long mi = m <= mlast ? m : mlast;
        if (HAS_EMB) {
            const HeadRow e = load_head_row(emb + (long long)a.act_idx[mi] * width + hc, sub);
#pragma unroll
            for (int i = 0; i < 8; ++i) u.f[i] += e.f[i];
        }
        const int p = min(max(a.pos_ids ? a.pos_ids[mi] : pos + a.pos0, 0), a.n_pos - 1);
        const float4 c4 = *reinterpret_cast<const float4*>(a.cos_tab + (long long)p * 32 + 4 * sub);
        const float4 s4 = *reinterpret_cast<const float4*>(a.sin_tab + (long long)p * 32 + 4 * sub);
        const float cs[4] = {c4.x, c4.y, c4.z, c4.w}, sn[4] = {s4.x, s4.y, s4.z, s4.w};
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) ss += u.f[i] * u.f[i];
        const float rstd = rsqrtf(group8_sum(ss) * (1.0f / HD) + a.eps);
        float o[8];
        if (is_q || is_k) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float n1 = wv[i] * (u.f[i] * rstd), n2 = wv[4 + i] * (u.f[4 + i] * rstd);
                o[i] = n1 * cs[i] - n2 * sn[i];
                o[4 + i] = n2 * cs[i] + n1 * sn[i];
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = u.f[i];
        }
        if (live) store_head_row(op, sub, o);
        op += ld_out;
        if (++pos == a.L) pos = 0;
    }
}

// backward: d_out (grad wrt rotated q,k and v) -> d_raw; accumulates d qn_w / d kn_w and the behaviour-embedding grads.
// A thread keeps ONE head for the whole kernel: the norm-weight partial sums live in registers and are flushed once;
// the behaviour-embedding partial sum is ONE register row for the current action index, flushed to shared memory when
// the index changes (every few tokens along the thread's contiguous range) — no per-token atomics or selects.
constexpr int MAX_EMB_ROWS = 4;

template <bool HAS_EMB>
__global__ void __launch_bounds__(384)
qk_norm_rope_bwd_kernel(HeadArgs a, const bf16* __restrict__ raw, long long ld_raw, const bf16* __restrict__ dout,
                        long long ld_dout, bf16* __restrict__ draw, long long ld_draw, float* __restrict__ d_qn_w,
                        float* __restrict__ d_kn_w, float* __restrict__ d_q_emb, float* __restrict__ d_k_emb,
                        float* __restrict__ d_v_emb, int emb_rows) {
    extern __shared__ float sh[];  // [2*64] norm-weight grads | [emb_rows * (n_q + 2 n_kv) * 64] embedding grads
    const int n_heads = a.n_q + 2 * a.n_kv;
    const int emb_w = n_heads * HD;
    float* sdw = sh;
    float* semb = sh + 2 * HD;
    const int tid = threadIdx.x + 8 * (threadIdx.y + n_heads * threadIdx.z);
    const int nthr = 8 * n_heads * blockDim.z;
    for (int i = tid; i < 2 * HD + (HAS_EMB ? emb_rows * emb_w : 0); i += nthr) sh[i] = 0.f;
    __syncthreads();
    const int sub = threadIdx.x, h = threadIdx.y, Z = blockDim.z;
    const bool is_q = h < a.n_q, is_k = !is_q && h < a.n_q + a.n_kv;
    const bf16* emb = is_q ? a.q_emb : (is_k ? a.k_emb : a.v_emb);
    const int width = is_q ? a.n_q * HD : a.n_kv * HD;
    const int hc = (is_q ? h : (is_k ? h - a.n_q : h - a.n_q - a.n_kv)) * HD;
    const float* wn = is_q ? a.qn_w : a.kn_w;
    const float4 w_lo = *reinterpret_cast<const float4*>(wn + 4 * sub);
    const float4 w_hi = *reinterpret_cast<const float4*>(wn + 32 + 4 * sub);
    const float wv[8] = {w_lo.x, w_lo.y, w_lo.z, w_lo.w, w_hi.x, w_hi.y, w_hi.z, w_hi.w};
    float dwn[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float demb[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int cur_act = -1;
    // columns of this thread within the head: 4*sub + i (i < 4), 32 + 4*sub + (i - 4)
    auto flush_emb = [&]() {
        if (cur_act >= 0 && cur_act < emb_rows) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                atomicAdd(&semb[cur_act * emb_w + h * HD + (i < 4 ? 4 * sub + i : 28 + 4 * sub + i)], demb[i]);
        }
    };

    const TokenRange tr = slice_range(a.M, Z, threadIdx.z);
    const long long mlast = a.M - 1;
    long long m = tr.begin;
    const long long mc0 = m < a.M ? m : mlast;
    const bf16* rp = raw + mc0 * ld_raw + h * HD;
    const bf16* dp = dout + mc0 * ld_dout + h * HD;
    bf16* wp = draw + mc0 * ld_draw + h * HD;
    int pos = (int)(mc0 % a.L);
    // PF rows of raw and of d_out are in flight per thread (a ring of raw loads refilled as it is consumed): measured
    // 8.8 -> 8.1 ms per step; the same ring made the forward slower (4.3 -> 5.1 ms), which keeps its one-row look-ahead
    HeadRaw ring_u[PF], ring_d[PF];
#pragma unroll
    for (int j = 0; j < PF; ++j) {
        ring_u[j] = load_head_raw(rp, sub);
        ring_d[j] = load_head_raw(dp, sub);
        if (m + j + 1 <= mlast) {
            rp += ld_raw;
            dp += ld_dout;
        }
    }
    int act_next = HAS_EMB ? a.act_idx[mc0] : 0;
    int p_next = a.pos_ids ? a.pos_ids[mc0] : 0;
    for (int it = 0; it < tr.iters; it += PF) {
#pragma unroll
      for (int j = 0; j < PF; ++j, ++m) {
        const bool live = m < tr.end;
        HeadRow u = head_row_from_raw(ring_u[j]);
        const HeadRow d = head_row_from_raw(ring_d[j]);
        ring_u[j] = load_head_raw(rp, sub);            // token m + PF
        ring_d[j] = load_head_raw(dp, sub);
        if (m + PF + 1 <= mlast) {
            rp += ld_raw;
            dp += ld_dout;
        }
        // the action index and RoPE position of a token are fetched one token ahead: the embedding row and the cos / sin
        // rows they select are dependent loads, which would otherwise sit behind a second memory latency every token
        const long long mn = m + 1 <= mlast ? m + 1 : mlast;
        const int act = act_next;
        if (HAS_EMB) {
            act_next = a.act_idx[mn];
            const HeadRow e = load_head_row(emb + (long long)act * width + hc, sub);
#pragma unroll
            for (int i = 0; i < 8; ++i) u.f[i] += e.f[i];
        }
        const int p = min(max(a.pos_ids ? p_next : pos + a.pos0, 0), a.n_pos - 1);
        if (a.pos_ids) p_next = a.pos_ids[mn];
        const float4 c4 = *reinterpret_cast<const float4*>(a.cos_tab + (long long)p * 32 + 4 * sub);
        const float4 s4 = *reinterpret_cast<const float4*>(a.sin_tab + (long long)p * 32 + 4 * sub);
        const float cs[4] = {c4.x, c4.y, c4.z, c4.w}, sn[4] = {s4.x, s4.y, s4.z, s4.w};
        // every lane runs the group reductions (a warp mixes q/k and v heads); v heads discard the result
        float du[8];
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) ss += u.f[i] * u.f[i];
        const float rstd = rsqrtf(group8_sum(ss) * (1.0f / HD) + a.eps);
        float dn[8], dot = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            // transpose of the rotation
            dn[i] = d.f[i] * cs[i] + d.f[4 + i] * sn[i];
            dn[4 + i] = d.f[4 + i] * cs[i] - d.f[i] * sn[i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            du[i] = wv[i] * dn[i];
            dot += du[i] * u.f[i];
        }
        dot = group8_sum(dot) * (1.0f / HD) * rstd * rstd * rstd;
        if (is_q || is_k) {
            const float lr = live ? rstd : 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                dwn[i] += dn[i] * u.f[i] * lr;
                du[i] = rstd * du[i] - u.f[i] * dot;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) du[i] = d.f[i];
        }
        if (live) {
            store_head_row(wp, sub, du);
            if (HAS_EMB) {
                if (act != cur_act) {
                    flush_emb();
                    cur_act = act;
#pragma unroll
                    for (int i = 0; i < 8; ++i) demb[i] = 0.f;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) demb[i] += du[i];
            }
        }
        wp += ld_draw;
        if (++pos == a.L) pos = 0;
      }
    }
    if (HAS_EMB) flush_emb();
    if (is_q || is_k) {
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(&sdw[(is_q ? 0 : HD) + (i < 4 ? 4 * sub + i : 28 + 4 * sub + i)], dwn[i]);
    }
    __syncthreads();
    for (int i = tid; i < HD; i += nthr) {
        if (sdw[i] != 0.f) atomicAdd(&d_qn_w[i], sdw[i]);
        if (sdw[HD + i] != 0.f) atomicAdd(&d_kn_w[i], sdw[HD + i]);
    }
    if (HAS_EMB) {
        const int qw = a.n_q * HD, kw = a.n_kv * HD;
        for (int i = tid; i < emb_rows * emb_w; i += nthr) {
            const float v = semb[i];
            if (v == 0.f) continue;
            const int r = i / emb_w, c = i % emb_w;
            if (c < qw) atomicAdd(&d_q_emb[r * qw + c], v);
            else if (c < qw + kw) atomicAdd(&d_k_emb[r * kw + (c - qw)], v);
            else atomicAdd(&d_v_emb[r * kw + (c - qw - kw)], v);
        }
    }
}

}  // namespace

extern "C" int gamer_rmsnorm_fwd(const void* x, const float* w, float eps, long long M, int H, void* out,
                                 long long ld_out, const int* row_map, const void* cat_table, const int* cat_idx,
                                 int cat_dim, float* rstd, cudaStream_t stream) {
    GAMER_REQUIRE(H == H256, "rmsnorm kernels are specialised for hidden size 256 (got %d)", H);
    GAMER_REQUIRE(cat_dim % 8 == 0 && cat_dim <= 256, "cat_dim must be a multiple of 8 and <= 256");
    if (M == 0) return 0;
    const int wpb = 8;
    const int grid = (int)((M + wpb - 1) / wpb < 148 * 8 ? (M + wpb - 1) / wpb : 148 * 8);
    rmsnorm_fwd_kernel<<<grid, wpb * 32, 0, stream>>>(reinterpret_cast<const bf16*>(x), w, eps, M,
                                                      reinterpret_cast<bf16*>(out), ld_out, row_map,
                                                      reinterpret_cast<const bf16*>(cat_table), cat_idx, cat_dim, rstd);
    GAMER_LAUNCH_CHECK();
    return 0;
}

extern "C" int gamer_rmsnorm_bwd(const void* x, const float* w, const float* rstd, float eps, long long M, int H,
                                 const void* dh, long long ld_dh, const int* row_map, const void* dres, void* dx,
                                 float* dw, const int* cat_idx, int cat_dim, int cat_rows, float* dcat,
                                 cudaStream_t stream) {
    GAMER_REQUIRE(H == H256, "rmsnorm kernels are specialised for hidden size 256 (got %d)", H);
    if (M == 0) return 0;
    const int wpb = 8;
    const int grid = (int)((M + wpb - 1) / wpb < 148 * 8 ? (M + wpb - 1) / wpb : 148 * 8);
    const int cr = dcat ? cat_rows : 0;
    const size_t smem = (H256 + cr * cat_dim) * sizeof(float);
    if (dcat != nullptr)
        rmsnorm_bwd_kernel<true><<<grid, wpb * 32, smem, stream>>>(
            reinterpret_cast<const bf16*>(x), w, rstd, eps, M, reinterpret_cast<const bf16*>(dh), ld_dh, row_map,
            reinterpret_cast<const bf16*>(dres), reinterpret_cast<bf16*>(dx), dw, cat_idx, cat_dim, cr, dcat);
    else
        rmsnorm_bwd_kernel<false><<<grid, wpb * 32, smem, stream>>>(
            reinterpret_cast<const bf16*>(x), w, rstd, eps, M, reinterpret_cast<const bf16*>(dh), ld_dh, row_map,
            reinterpret_cast<const bf16*>(dres), reinterpret_cast<bf16*>(dx), dw, cat_idx, cat_dim, cr, dcat);
    GAMER_LAUNCH_CHECK();
    return 0;
}

static HeadArgs make_head_args(long long M, int L, int n_q, int n_kv, const int* pos_ids, int pos0, int n_pos,
                               const float* cos_tab, const float* sin_tab, const float* qn_w, const float* kn_w,
                               const void* q_emb, const void* k_emb, const void* v_emb, const int* act_idx, float eps) {
    HeadArgs a;
    a.M = M; a.L = L; a.n_q = n_q; a.n_kv = n_kv; a.pos_ids = pos_ids; a.pos0 = pos0; a.n_pos = n_pos;
    a.cos_tab = cos_tab; a.sin_tab = sin_tab; a.qn_w = qn_w; a.kn_w = kn_w;
    a.q_emb = reinterpret_cast<const bf16*>(q_emb); a.k_emb = reinterpret_cast<const bf16*>(k_emb);
    a.v_emb = reinterpret_cast<const bf16*>(v_emb); a.act_idx = act_idx; a.eps = eps;
    return a;
}

extern "C" int gamer_qk_norm_rope_fwd(const void* raw, long long ld_raw, void* out, long long ld_out, long long M, int L,
                                      int n_q, int n_kv, int head_dim, const int* pos_ids, int pos0, int n_pos,
                                      const float* cos_tab, const float* sin_tab, const float* qn_w, const float* kn_w,
                                      const void* q_emb, const void* k_emb, const void* v_emb, const int* act_idx,
                                      float eps, cudaStream_t stream) {
    GAMER_REQUIRE(head_dim == HD, "head kernels are specialised for head_dim 64 (got %d)", head_dim);
    GAMER_REQUIRE(n_pos > 0, "the RoPE tables need at least one row (n_pos = %d)", n_pos);
    if (M == 0) return 0;
    HeadArgs a = make_head_args(M, L, n_q, n_kv, pos_ids, pos0, n_pos, cos_tab, sin_tab, qn_w, kn_w, q_emb, k_emb, v_emb,
                                act_idx, eps);
    const int n_heads = n_q + 2 * n_kv;
    GAMER_REQUIRE(n_heads <= 48, "too many heads for the (8, heads, Z) block layout");
    const int Z = n_heads <= 12 ? 4 : (n_heads <= 24 ? 2 : 1);
    const dim3 block(8, n_heads, Z);
    const long long want = (M + Z - 1) / Z;
    const int grid = (int)(want < 148 * 8 ? want : 148 * 8);
    if (q_emb != nullptr)
        qk_norm_rope_fwd_kernel<true><<<grid, block, 0, stream>>>(a, reinterpret_cast<const bf16*>(raw), ld_raw,
                                                                  reinterpret_cast<bf16*>(out), ld_out);
    else
        qk_norm_rope_fwd_kernel<false><<<grid, block, 0, stream>>>(a, reinterpret_cast<const bf16*>(raw), ld_raw,
                                                                   reinterpret_cast<bf16*>(out), ld_out);
    GAMER_LAUNCH_CHECK();
    return 0;
}

extern "C" int gamer_qk_norm_rope_bwd(const void* raw, long long ld_raw, const void* dout, long long ld_dout, void* draw,
                                      long long ld_draw, long long M, int L, int n_q, int n_kv, int head_dim,
                                      const int* pos_ids, int pos0, int n_pos, const float* cos_tab, const float* sin_tab,
                                      const float* qn_w, const float* kn_w, const void* q_emb, const void* k_emb,
                                      const void* v_emb, const int* act_idx, int emb_rows, float eps, float* d_qn_w,
                                      float* d_kn_w, float* d_q_emb, float* d_k_emb, float* d_v_emb,
                                      cudaStream_t stream) {
    GAMER_REQUIRE(head_dim == HD, "head kernels are specialised for head_dim 64 (got %d)", head_dim);
    GAMER_REQUIRE(n_pos > 0, "the RoPE tables need at least one row (n_pos = %d)", n_pos);
    if (M == 0) return 0;
    HeadArgs a = make_head_args(M, L, n_q, n_kv, pos_ids, pos0, n_pos, cos_tab, sin_tab, qn_w, kn_w, q_emb, k_emb, v_emb,
                                act_idx, eps);
    const bool has_emb = q_emb != nullptr;
    GAMER_REQUIRE(!has_emb || (d_q_emb && d_k_emb && d_v_emb), "behaviour-embedding grads missing");
    const size_t smem = (2 * HD + (has_emb ? emb_rows * (n_q + 2 * n_kv) * HD : 0)) * sizeof(float);
    GAMER_REQUIRE(smem <= 48 * 1024, "too many behaviour rows for the shared-memory accumulator");
    GAMER_REQUIRE(!has_emb || emb_rows <= MAX_EMB_ROWS, "at most %d behaviour-embedding rows (num_behavior + 1)", MAX_EMB_ROWS);
    const int n_heads = n_q + 2 * n_kv;
    GAMER_REQUIRE(n_heads <= 48, "too many heads for the (8, heads, Z) block layout");
    const int Z = n_heads <= 12 ? 4 : (n_heads <= 24 ? 2 : 1);
    const dim3 block(8, n_heads, Z);
    const long long want = (M + Z - 1) / Z;
    const int grid = (int)(want < 148 * 3 ? want : 148 * 3);
    if (has_emb)
        qk_norm_rope_bwd_kernel<true><<<grid, block, smem, stream>>>(
            a, reinterpret_cast<const bf16*>(raw), ld_raw, reinterpret_cast<const bf16*>(dout), ld_dout,
            reinterpret_cast<bf16*>(draw), ld_draw, d_qn_w, d_kn_w, d_q_emb, d_k_emb, d_v_emb, emb_rows);
    else
        qk_norm_rope_bwd_kernel<false><<<grid, block, smem, stream>>>(
            a, reinterpret_cast<const bf16*>(raw), ld_raw, reinterpret_cast<const bf16*>(dout), ld_dout,
            reinterpret_cast<bf16*>(draw), ld_draw, d_qn_w, d_kn_w, nullptr, nullptr, nullptr, emb_rows);
    GAMER_LAUNCH_CHECK();
    return 0;
}
