// K6 on the 5th-gen tensor cores: session / behaviour-masked attention, forward and fused backward, GQA 2:1, head_dim 64.
//
// Every matrix product (QK^T, PV, dO V^T, P^T dO, dS^T Q, dS K) is a tcgen05.mma over 128x128 (or 128x64) tiles with
// TMA-staged SWIZZLE_128B operands and TMEM accumulators; the softmax / mask / dS arithmetic runs in dedicated warps that
// read the score tiles from TMEM (one query row per thread) and hand P / dS back through shared memory.
//
//   forward  : CTA = (sequence, kv head, 128-query tile).  The two query heads of the GQA group ping-pong: while the
//              softmax warpgroup of head A works on tile t, the tensor core computes S_B(t) / O_B += P_B V, so the MUFU and
//              the tensor pipe overlap.  Online softmax with a lazy rescale (the running max is only raised — and O only
//              rescaled, on a cold path — when it would grow by more than 2^64; P stays representable in bf16).
//   backward : CTA = (sequence, kv head, 128-key tile); loops over (head of the group, query tile).  dK and dV accumulate
//              in TMEM over the whole loop; dQ tiles leave through a TMA fp32 reduce-add into a tile-major accumulator.
//
// Mask predicate (reference: SeqRec/models/generative/Qwen3Multi/model.py:573-741, Qwen3SessionMoe/model.py:416-468) is
// evaluated per (query, key) from per-key codes that fold the key padding mask in (attn_meta_kernel).  Rows with no
// allowed key are "uniform rows" (quirk Q1): forward output = column mean of V over all L keys, lse = +inf; backward uses
// P = 1/L over all keys (the analytic gradient of that uniform softmax).
#include <stdlib.h>

#include "attention_tc.cuh"
#include "sm100_ptx.cuh"

using namespace sm100;

namespace {

constexpr int D = 64;
constexpr int BT = 128;               // tile edge: queries and keys
constexpr int TILE_BYTES = BT * D * 2;  // 16 KB: one [128 x 64] bf16 SWIZZLE_128B tile
constexpr int KC_MAX = 0x7fffffff;

// Optional in-kernel timeline (debug hook gamer_attn_set_trace): CTA 0 appends (tag, clock64) pairs.
struct Trace {
    long long* buf;
    int cap;
};
__device__ __forceinline__ void trace_pt(const Trace& tr, int role, int& n, int tag) {
    if (tr.buf != nullptr && blockIdx.x == 0 && n < tr.cap) {
        tr.buf[(role * tr.cap + n) * 2] = tag;
        tr.buf[(role * tr.cap + n) * 2 + 1] = clock64();
        ++n;
    }
}
Trace g_trace = {nullptr, 0};

template <int KIND>
__host__ __device__ constexpr bool kind_causal() { return KIND == MASK_CAUSAL || KIND == MASK_MULTI_CROSS; }
template <int KIND>
__host__ __device__ constexpr bool kind_uses_act() { return KIND == MASK_MULTI_CROSS || KIND == MASK_SESSION_CROSS; }
template <int KIND>
__host__ __device__ constexpr bool kind_uses_sess() { return KIND == MASK_SESSION || KIND == MASK_SESSION_CROSS; }

// ---------------------------------------------------------------------------------------------------------------
// per-key codes: ka = valid ? act : MAX, ks = valid ? sess : MAX (valid = j < L and attention_mask[j]); tile flag = all
// 128 keys of the tile valid.  One block per (sequence, key tile).
// ---------------------------------------------------------------------------------------------------------------
__global__ void attn_meta_kernel(const int* __restrict__ am, const int* __restrict__ act, const int* __restrict__ sess,
                                 int L, int Lp, int k_tiles, int* __restrict__ ka, int* __restrict__ ks,
                                 int* __restrict__ qa, int* __restrict__ qs, int* __restrict__ tflag) {
    const int b = blockIdx.x / k_tiles, t = blockIdx.x % k_tiles;
    const int j = t * BT + threadIdx.x;
    bool valid = false;
    int a = KC_MAX, s = KC_MAX, a_raw = 0, s_raw = 0;
    if (j < L) {
        const long long idx = (long long)b * L + j;
        valid = am[idx] != 0;
        a_raw = act ? act[idx] : 0;
        s_raw = sess ? sess[idx] : 0;
        if (valid) {
            a = a_raw;
            s = s_raw;
        }
    }
    ka[(long long)b * Lp + j] = a;
    ks[(long long)b * Lp + j] = s;
    qa[(long long)b * Lp + j] = a_raw;  // query side: the predicate does not look at the query's own padding bit
    qs[(long long)b * Lp + j] = s_raw;
    const int all = __syncthreads_and(valid ? 1 : 0);
    if (threadIdx.x == 0) tflag[blockIdx.x] = all;
}

template <int KIND, bool DIAG>
__device__ __forceinline__ bool allow_tc(int ka, int ks, int act_i, int sess_i, int c, int row, int j, int i, int istart) {
    if constexpr (KIND == MASK_CAUSAL) {
        bool ok = (ka == 0);
        if (DIAG) ok = ok && (c <= row);
        return ok;
    } else if constexpr (KIND == MASK_MULTI_CROSS) {
        bool ok = (ka < act_i);
        if (DIAG) ok = ok && (c <= row);
        return ok;
    } else if constexpr (KIND == MASK_SESSION) {
        return (ks < sess_i) || (j >= istart && j <= i && ks != KC_MAX);
    } else {
        return (ks < sess_i) && (ka < act_i);
    }
}

__device__ __forceinline__ uint64_t desc_k(uint32_t saddr) { return umma_desc_sw128(saddr, 16, 1024); }
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo) { return umma_desc_sw128(saddr, lbo, 1024); }

// D[128 x 128] = A[128 x 64] B[128 x 64]^T, both K-major tiles (k = head dim)
__device__ __forceinline__ void issue_nt_128x128x64(uint32_t d_tmem, uint32_t sa, uint32_t sb) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, desc_k(sa + k * 32), desc_k(sb + k * 32), idesc, k != 0);
}
// D[128 x 64] (+)= A[128 x 128] B[128 x 64]: A K-major in two 64-wide halves (16 KB apart), B MN-major [128 k-rows x 64]
__device__ __forceinline__ void issue_nn_128x64x128(uint32_t d_tmem, uint32_t sa, uint32_t sb, bool acc) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);
#pragma unroll
    for (int k = 0; k < 8; ++k)
        umma_bf16(d_tmem, desc_k(sa + (k >> 2) * TILE_BYTES + (k & 3) * 32), desc_mn(sb + k * 2048, TILE_BYTES), idesc,
                  (acc || k != 0) ? 1u : 0u);
}
// D[128 x 64] (+)= A^T B with A stored [128 k-rows x 128 m] (two 64-wide halves, MN-major) and B [128 k-rows x 64] MN-major
__device__ __forceinline__ void issue_tn_128x64x128(uint32_t d_tmem, uint32_t sa, uint32_t sb, bool acc) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
#pragma unroll
    for (int k = 0; k < 8; ++k)
        umma_bf16(d_tmem, desc_mn(sa + k * 2048, TILE_BYTES), desc_mn(sb + k * 2048, TILE_BYTES), idesc,
                  (acc || k != 0) ? 1u : 0u);
}

// cold path of the online softmax: multiply this thread's 32 fp32 TMEM columns of an O row by alpha
__device__ __noinline__ void rescale_o_row(uint32_t t_o, float alpha) {
    uint32_t o[32];
    tmem_ld_32x32(t_o, o);
    tmem_ld_wait();
#pragma unroll
    for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
    tmem_st_32x32(t_o, o);
    tmem_st_wait();
}

// =================================================================================================================
// forward
// =================================================================================================================
constexpr int F_STAGES = 3;
constexpr int F_META_BYTES = 1024;                             // ka[128] | ks[128]
constexpr int F_STAGE_BYTES = 2 * TILE_BYTES + F_META_BYTES;   // K | V | codes
constexpr int F_OFF_Q = 0;                                     // 2 heads
constexpr int F_OFF_KV = 2 * TILE_BYTES;
constexpr int F_OFF_P = F_OFF_KV + F_STAGES * F_STAGE_BYTES;   // 2 heads x 2 halves
constexpr int F_OFF_BAR = F_OFF_P + 4 * TILE_BYTES;
constexpr int F_OFF_X = F_OFF_BAR + 256;                      // row max / row sum exchange: [2 heads][2][2][128] fp32
constexpr int F_SMEM = F_OFF_X + 4096 + 1024;
constexpr int F_THREADS = 640;

struct FwdParams {
    int B, L, Lp, n_q, n_kv, P, q_tiles, k_tiles, total;
    const int* ka;
    const int* ks;
    const int* tflag;
    const int* act;
    const int* sess;
    float scale_log2;
    const float* vmean;
    float* lse;
    DropParams drop;  // attention-probability dropout (8-bit thresholds); thresh == 0: off
    Trace tr;
};

template <int KIND>
__device__ __forceinline__ void fwd_decode(int w, const FwdParams& p, int& b, int& g, int& qt, int& T) {
    const int per = p.B * p.n_kv;
    qt = p.q_tiles - 1 - w / per;  // heaviest (most key tiles) first
    const int rem = w % per;
    b = rem / p.n_kv;
    g = rem % p.n_kv;
    T = kind_causal<KIND>() ? min(qt + 1, p.k_tiles) : p.k_tiles;
}

// mask + partial row max over this thread's 64 scores of a tile (4 independent max chains).  `cc0` = first key column of
// the thread's half within the 128-key tile.  CODES = false: the tile's keys are all valid and the kind is CAUSAL, so
// only the diagonal test remains.
template <int KIND, bool DIAG, bool CODES>
__device__ __forceinline__ float mask_max(uint32_t (&s)[64], const int4* mk4, int act_i, int sess_i, int cc0, int row, int j0,
                                          int i, int istart) {
    float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int c4 = 0; c4 < 16; ++c4) {
        int4 a4 = make_int4(0, 0, 0, 0), s4 = make_int4(0, 0, 0, 0);
        if (CODES && KIND != MASK_SESSION) a4 = mk4[c4];
        if (CODES && kind_uses_sess<KIND>()) s4 = mk4[32 + c4];
        const int av[4] = {a4.x, a4.y, a4.z, a4.w};
        const int sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = c4 * 4 + e;
            const bool ok = allow_tc<KIND, DIAG>(av[e], sv[e], act_i, sess_i, cc0 + c, row, j0 + c, i, istart);
            const float v = ok ? __uint_as_float(s[c]) : -INFINITY;
            s[c] = __float_as_uint(v);
            mx[e] = fmaxf(mx[e], v);
        }
    }
    return fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
}

template <int KIND>
__global__ void __launch_bounds__(F_THREADS, 1)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, FwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F_OFF_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* q_empty = bars + 1;
    uint64_t* kv_full = bars + 2;                 // [F_STAGES]
    uint64_t* kv_empty = kv_full + F_STAGES;      // [F_STAGES]
    uint64_t* s_full = kv_empty + F_STAGES;       // [2]
    uint64_t* p_full = s_full + 2;                // [2]
    uint64_t* o_full = p_full + 2;                // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        prefetch_tmap(&tmQ);
        prefetch_tmap(&tmK);
        prefetch_tmap(&tmV);
        prefetch_tmap(&tmO);
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int s = 0; s < F_STAGES; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        for (int h = 0; h < 2; ++h) {
            mbar_init(&s_full[h], 1);
            mbar_init(&p_full[h], 256);
            mbar_init(&o_full[h], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // TMEM columns: S_A [0,128)  S_B [128,256)  O_A [256,320)  O_B [320,384)

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int st = 0;
            uint32_t ph = 0, item_par = 0;
            for (int w = blockIdx.x; w < p.total; w += gridDim.x) {
                int b, g, qt, T;
                fwd_decode<KIND>(w, p, b, g, qt, T);
                mbar_wait(q_empty, item_par ^ 1);
                mbar_expect_tx(q_full, 2 * TILE_BYTES);
                tma_load_3d(smem + F_OFF_Q, &tmQ, q_full, (2 * g) * D, qt * BT, b);
                tma_load_3d(smem + F_OFF_Q + TILE_BYTES, &tmQ, q_full, (2 * g + 1) * D, qt * BT, b);
                for (int t = 0; t < T; ++t) {
                    mbar_wait(&kv_empty[st], ph ^ 1);
                    uint8_t* sk = smem + F_OFF_KV + st * F_STAGE_BYTES;
                    mbar_expect_tx(&kv_full[st], F_STAGE_BYTES);
                    tma_load_3d(sk, &tmK, &kv_full[st], g * D, t * BT, b);
                    tma_load_3d(sk + TILE_BYTES, &tmV, &kv_full[st], g * D, t * BT, b);
                    bulk_load_1d(sk + 2 * TILE_BYTES, p.ka + (long long)b * p.Lp + t * BT, 512, &kv_full[st]);
                    bulk_load_1d(sk + 2 * TILE_BYTES + 512, p.ks + (long long)b * p.Lp + t * BT, 512, &kv_full[st]);
                    if (++st == F_STAGES) {
                        st = 0;
                        ph ^= 1;
                    }
                }
                item_par ^= 1;
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int st = 0, tn = 0;
            uint32_t ph = 0, item_par = 0, tile_n = 0;
            const uint32_t sq = smem_u32(smem + F_OFF_Q);
            const uint32_t sp = smem_u32(smem + F_OFF_P);
            for (int w = blockIdx.x; w < p.total; w += gridDim.x) {
                int b, g, qt, T;
                fwd_decode<KIND>(w, p, b, g, qt, T);
                mbar_wait(q_full, item_par);
                trace_pt(p.tr, 0, tn, 1);
                mbar_wait(&kv_full[st], ph);
                tc_fence_after();
                trace_pt(p.tr, 0, tn, 2);
                {
                    const uint32_t sk = smem_u32(smem + F_OFF_KV + st * F_STAGE_BYTES);
                    issue_nt_128x128x64(tmem_base + 0, sq, sk);
                    umma_commit(&s_full[0]);
                    issue_nt_128x128x64(tmem_base + 128, sq + TILE_BYTES, sk);
                    umma_commit(&s_full[1]);
                    if (T == 1) umma_commit(q_empty);  // Q is dead once the item's last S tiles are done
                }
                for (int t = 0; t < T; ++t) {
                    int nst = st + 1;
                    uint32_t nph = ph;
                    if (nst == F_STAGES) {
                        nst = 0;
                        nph ^= 1;
                    }
                    const uint32_t sv = smem_u32(smem + F_OFF_KV + st * F_STAGE_BYTES + TILE_BYTES);
                    const uint32_t skn = smem_u32(smem + F_OFF_KV + nst * F_STAGE_BYTES);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        mbar_wait(&p_full[h], tile_n & 1);
                        tc_fence_after();
                        trace_pt(p.tr, 0, tn, 10 + h);
                        if (t + 1 < T) {
                            if (h == 0) {
                                mbar_wait(&kv_full[nst], nph);
                                tc_fence_after();
                            }
                            issue_nt_128x128x64(tmem_base + h * 128, sq + h * TILE_BYTES, skn);
                            umma_commit(&s_full[h]);
                            if (h == 1 && t + 2 == T) umma_commit(q_empty);
                        }
                        issue_nn_128x64x128(tmem_base + 256 + h * 64, sp + h * 2 * TILE_BYTES, sv, t > 0);
                        umma_commit(&o_full[h]);
                    }
                    umma_commit(&kv_empty[st]);
                    st = nst;
                    ph = nph;
                    ++tile_n;
                }
                item_par ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ===================== softmax: 16 warps.  warps 4-11 = head A, 12-19 = head B; within a head the first four
        // warps take key columns 0-63 of the tile, the other four 64-127 (one query row per thread, row max exchanged
        // through shared memory) =====================
        const int hd = (warp - 4) >> 3;
        const int half = ((warp - 4) >> 2) & 1;
        const int wq = warp & 3;  // TMEM lane quarter of this warp
        const int row = wq * 32 + lane;
        const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
        const uint32_t t_s = tmem_base + lane_off + hd * 128 + half * 64;
        const uint32_t t_o = tmem_base + lane_off + 256 + hd * 64 + half * 32;
        uint8_t* sStage = smem + F_OFF_P + hd * 2 * TILE_BYTES;          // O staging = this head's first P half buffer
        const uint32_t sPh = smem_u32(sStage + half * TILE_BYTES);        // the P half this thread writes
        float* xchg = reinterpret_cast<float*>(smem + F_OFF_X) + hd * 512;  // [2 buffers][2 halves][128 rows]
        const int bar_id = 1 + hd;
        int st = 0, tn = 0;
        uint32_t ph = 0, tile_n = 0;
        Trace tr = p.tr;
        if (row != 0 || half != 0) tr.buf = nullptr;
        for (int w = blockIdx.x; w < p.total; w += gridDim.x) {
            int b, g, qt, T;
            fwd_decode<KIND>(w, p, b, g, qt, T);
            const int h = 2 * g + hd;
            const int i = qt * BT + row;
            int act_i = 0, sess_i = 0;
            if (i < p.L) {
                if (kind_uses_act<KIND>()) act_i = p.act[(long long)b * p.L + i];
                if (kind_uses_sess<KIND>()) sess_i = p.sess[(long long)b * p.L + i];
            }
            const int istart = (i / p.P) * p.P;
            unsigned allvalid = 0;  // bit t: every key of tile t is valid (CAUSAL tiles below the diagonal then need no mask)
            if (KIND == MASK_CAUSAL) {
                for (int t = 0; t < T; ++t) allvalid |= (p.tflag[b * p.k_tiles + t] != 0 ? 1u : 0u) << t;
            }
            float m = -INFINITY, l = 0.f;
            for (int t = 0; t < T; ++t) {
                const int4* mk4 =
                    reinterpret_cast<const int4*>(smem + F_OFF_KV + st * F_STAGE_BYTES + 2 * TILE_BYTES) + half * 16;
                const bool diag = kind_causal<KIND>() && (t == qt);
                const bool codes = (KIND != MASK_CAUSAL) || (((allvalid >> t) & 1u) == 0);
                mbar_wait(&kv_full[st], ph);  // key codes of this stage (already complete: the S MMA waited on it)
                trace_pt(tr, 1 + hd, tn, 20);
                mbar_wait(&s_full[hd], tile_n & 1);
                tc_fence_after();
                trace_pt(tr, 1 + hd, tn, 21);
                uint32_t s[64];
                tmem_ld_32x32(t_s, s);
                tmem_ld_32x32(t_s + 32, s + 32);
                tmem_ld_wait();
                trace_pt(tr, 1 + hd, tn, 22);
                float mx;
                const int j0 = t * BT + half * 64;
                if (codes) {
                    if (diag) mx = mask_max<KIND, true, true>(s, mk4, act_i, sess_i, half * 64, row, j0, i, istart);
                    else mx = mask_max<KIND, false, true>(s, mk4, act_i, sess_i, half * 64, row, j0, i, istart);
                } else if (diag) {
                    mx = mask_max<KIND, true, false>(s, mk4, act_i, sess_i, half * 64, row, j0, i, istart);
                } else {
                    float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                    for (int c = 0; c < 64; ++c) mx4[c & 3] = fmaxf(mx4[c & 3], __uint_as_float(s[c]));
                    mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
                }
                // the other half of the row
                float* xb = xchg + (tile_n & 1) * 256;
                xb[half * 128 + row] = mx;
                named_bar_sync(bar_id, 256);
                mx = fmaxf(mx, xb[(half ^ 1) * 128 + row]);
                // lazy running max (log2 domain): raise it only from -inf or by more than 2^64
                const float m_tile = mx * p.scale_log2;
                float m_new = m;
                if (m == -INFINITY) m_new = m_tile;
                else if (m_tile > m + 64.f) m_new = m_tile;
                const bool rescale = (m != -INFINITY) && (m_new != m);
                trace_pt(tr, 1 + hd, tn, 23);
                if (t > 0) {
                    mbar_wait(&o_full[hd], (tile_n - 1) & 1);  // PV(t-1) done: O stable, P buffer free
                    tc_fence_after();
                    if (__any_sync(0xffffffffu, rescale)) rescale_o_row(t_o, rescale ? ex2_approx(m - m_new) : 1.f);
                }
                if (rescale) l *= ex2_approx(m - m_new);
                trace_pt(tr, 1 + hd, tn, 24);
                m = m_new;
                const float neg_m = (m == -INFINITY) ? 0.f : -m;
                float sum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int c = 0; c < 64; c += 2) {
                    const float p0 = ex2_approx(fmaf(__uint_as_float(s[c]), p.scale_log2, neg_m));
                    const float p1 = ex2_approx(fmaf(__uint_as_float(s[c + 1]), p.scale_log2, neg_m));
                    sum[(c >> 1) & 1] += p0;
                    sum[2 + ((c >> 1) & 1)] += p1;
                    s[c >> 1] = pack_bf16(p0, p1);
                }
                l += (sum[0] + sum[1]) + (sum[2] + sum[3]);
                trace_pt(tr, 1 + hd, tn, 25);
#pragma unroll
                for (int ch = 0; ch < 8; ++ch)
                    sts128(sPh + row * 128 + ((ch ^ (row & 7)) << 4), s[4 * ch], s[4 * ch + 1], s[4 * ch + 2], s[4 * ch + 3]);
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(&p_full[hd]);
                trace_pt(tr, 1 + hd, tn, 26);
                if (++st == F_STAGES) {
                    st = 0;
                    ph ^= 1;
                }
                ++tile_n;
            }
            // ---- epilogue: O / l (or the V column mean on uniform rows) -> bf16 -> smem -> TMA store
            {
                float* xb = xchg + (tile_n & 1) * 256;
                xb[half * 128 + row] = l;
                named_bar_sync(bar_id, 256);
                l += xb[(half ^ 1) * 128 + row];
            }
            mbar_wait(&o_full[hd], (tile_n - 1) & 1);
            tc_fence_after();
            trace_pt(tr, 1 + hd, tn, 30);
            const bool uniform = !(l > 0.f);
            const float inv = uniform ? 0.f : 1.f / l;
            const float* vm = p.vmean + ((long long)b * p.n_kv + g) * D + half * 32;
            {
                uint32_t o[32];
                tmem_ld_32x32(t_o, o);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(o[8 * q + e]) * inv;
                    if (uniform) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = vm[8 * q + e];
                    }
                    const int ch = half * 4 + q;
                    *reinterpret_cast<bf16x8*>(sStage + row * 128 + ((ch ^ (row & 7)) << 4)) = float_to_bf16x8(v);
                }
            }
            tc_fence_before();
            fence_proxy_async();
            named_bar_sync(bar_id, 256);
            if (row == 0 && half == 0) {
                tma_store_3d(&tmO, sStage, h * D, qt * BT, b);
                bulk_commit();
                bulk_wait_read0();
            }
            named_bar_sync(bar_id, 256);
            trace_pt(tr, 1 + hd, tn, 31);
            if (half == 0 && i < p.L) p.lse[((long long)b * p.n_q + h) * p.L + i] = uniform ? INFINITY : (m + log2f(l));
        }
        if (row == 0 && half == 0) bulk_wait0();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}


// =================================================================================================================
// forward, "small CTA" variant: one 128-thread CTA per (sequence, query head, 128-query tile), 64-key tiles, three CTAs
// resident per SM (TMEM: 128 columns each, smem ~66 KB each).  Inside a CTA everything is sequential — thread 0 issues
// the TMA loads and the MMAs, all four warps run the softmax (one query row per thread, the whole 64-key row of the
// tile in registers, so the row max needs no exchange) — and the latencies of one CTA are hidden by the other two.
// The exp2 throughput (16 / clk / SM) is what bounds it.
// =================================================================================================================
constexpr int S_KT = 64;                                   // keys per tile
constexpr int S_HALF = S_KT * D * 2;                       // 8 KB: one [64 x 64] bf16 tile
constexpr int S_OFF_Q = 0;                                 // [128 x 64] Q, later the O staging tile
constexpr int S_OFF_K = TILE_BYTES;                        // 2 stages
constexpr int S_OFF_V = S_OFF_K + 2 * S_HALF;              // 2 stages
constexpr int S_OFF_P = S_OFF_V + 2 * S_HALF;              // [128 q x 64 keys] bf16
constexpr int S_OFF_C = S_OFF_P + TILE_BYTES;              // key codes ring: 4 x (ka[64] | ks[64])
constexpr int S_OFF_BAR = S_OFF_C + 4 * 512;
constexpr int S_SMEM = S_OFF_BAR + 128 + 1024;
constexpr int S_THREADS = 160;                             // 4 softmax warps + 1 issuer warp

template <int KIND, bool DIAG, bool CODES>
__device__ __forceinline__ float mask_max64(uint32_t (&s)[64], const int4* mk4, int act_i, int sess_i, int cc0, int row, int j0,
                                            int i, int istart) {
    float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int c4 = 0; c4 < 16; ++c4) {
        int4 a4 = make_int4(0, 0, 0, 0), s4 = make_int4(0, 0, 0, 0);
        if (CODES && KIND != MASK_SESSION) a4 = mk4[c4];
        if (CODES && kind_uses_sess<KIND>()) s4 = mk4[16 + c4];
        const int av[4] = {a4.x, a4.y, a4.z, a4.w};
        const int sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = c4 * 4 + e;
            const bool ok = allow_tc<KIND, DIAG>(av[e], sv[e], act_i, sess_i, cc0 + c, row, j0 + c, i, istart);
            const float v = ok ? __uint_as_float(s[c]) : -INFINITY;
            s[c] = __float_as_uint(v);
            mx[e] = fmaxf(mx[e], v);
        }
    }
    return fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
}

// cold path: multiply this thread's O row (64 fp32 TMEM columns) by alpha
__device__ __noinline__ void rescale_o_row64(uint32_t t_o, float alpha) {
#pragma unroll 1
    for (int hh = 0; hh < 2; ++hh) {
        uint32_t o[32];
        tmem_ld_32x32(t_o + hh * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
        tmem_st_32x32(t_o + hh * 32, o);
    }
    tmem_st_wait();
}

// Dropout of the probabilities (SDPA dropout_p, Qwen3Multi/model.py:139): P is zeroed where the Philox byte of (sequence,
// head, query, key) is below the threshold; the row sum l keeps the undropped P (softmax normalisation happens before
// dropout) and the 1/keep scale is folded into the final O / l.  Uniform rows (quirk Q1) take their expectation: the
// column mean of V, no dropout.
template <int KIND, bool DROP>
__global__ void __launch_bounds__(S_THREADS, 3)
attn_fwd_small_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, FwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S_OFF_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;   // [2]
    uint64_t* v_full = bars + 3;   // [2]
    uint64_t* c_full = bars + 5;   // [4] key codes
    uint64_t* s_full = bars + 9;
    uint64_t* s_free = bars + 10;  // all softmax threads have read S out of TMEM
    uint64_t* p_full = bars + 11;  // P tile written (and O rescaled)
    uint64_t* o_full = bars + 12;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

    const int warp = threadIdx.x >> 5;
    // work item: heaviest query tiles first
    const int per = p.B * p.n_q;
    const int qt = p.q_tiles - 1 - (int)blockIdx.x / per;
    const int rem = (int)blockIdx.x % per;
    const int b = rem / p.n_q, h = rem % p.n_q;
    const int g = h / (p.n_q / p.n_kv);
    const int kt_all = (p.L + S_KT - 1) / S_KT;
    const int nkt = kind_causal<KIND>() ? min(2 * (qt + 1), kt_all) : kt_all;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmQ);
        prefetch_tmap(&tmK);
        prefetch_tmap(&tmV);
        prefetch_tmap(&tmO);
        for (int i = 0; i < 13; ++i) mbar_init(&bars[i], (i == 10 || i == 11) ? 128 : 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc<128>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t sq = smem_u32(smem + S_OFF_Q), sp = smem_u32(smem + S_OFF_P);

    if (warp == 4) {
        // ===================== issuer: TMA loads + MMAs =====================
        if ((threadIdx.x & 31) == 0) {
            auto load_k = [&](int j) {  // K tile j and its key codes
                const int st = j & 1, cs = j & 3;
                mbar_expect_tx(&k_full[st], S_HALF);
                tma_load_3d(smem + S_OFF_K + st * S_HALF, &tmK, &k_full[st], g * D, j * S_KT, b);
                mbar_expect_tx(&c_full[cs], 512);
                bulk_load_1d(smem + S_OFF_C + cs * 512, p.ka + (long long)b * p.Lp + j * S_KT, 256, &c_full[cs]);
                bulk_load_1d(smem + S_OFF_C + cs * 512 + 256, p.ks + (long long)b * p.Lp + j * S_KT, 256, &c_full[cs]);
            };
            auto load_v = [&](int j) {
                const int st = j & 1;
                mbar_expect_tx(&v_full[st], S_HALF);
                tma_load_3d(smem + S_OFF_V + st * S_HALF, &tmV, &v_full[st], g * D, j * S_KT, b);
            };
            auto issue_s = [&](int j) {  // S[128 x 64] = Q K_j^T
                constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
                const uint32_t sk = smem_u32(smem + S_OFF_K + (j & 1) * S_HALF);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, desc_k(sq + k * 32), desc_k(sk + k * 32), idesc, k != 0);
                umma_commit(s_full);
            };
            auto issue_pv = [&](int j) {  // O[128 x 64] (+)= P[128 x 64 keys] V_j[64 keys x 64]
                constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);
                const uint32_t sv = smem_u32(smem + S_OFF_V + (j & 1) * S_HALF);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base + 64, desc_k(sp + k * 32), desc_mn(sv + k * 2048, S_HALF), idesc,
                              (j > 0 || k != 0) ? 1u : 0u);
                umma_commit(o_full);
            };
            mbar_expect_tx(q_full, TILE_BYTES);
            tma_load_3d(smem + S_OFF_Q, &tmQ, q_full, h * D, qt * BT, b);
            load_k(0);
            load_v(0);
            if (nkt > 1) {
                load_k(1);
                load_v(1);
            }
            mbar_wait(q_full, 0);
            mbar_wait(&k_full[0], 0);
            tc_fence_after();
            issue_s(0);
            for (int j = 0; j < nkt; ++j) {
                if (j + 1 < nkt) {  // S(j+1) as soon as S(j) has been read out of TMEM
                    mbar_wait(s_free, j & 1);
                    mbar_wait(&k_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
                    tc_fence_after();
                    issue_s(j + 1);
                    // K_j is dead (S(j) completed before it was read out): prefetch two tiles ahead.  The codes slot
                    // (j+2)&3 was last read during tile j-2, which p_full(j-2) (waited below, last iteration) closes.
                    if (j + 2 < nkt) load_k(j + 2);
                }
                mbar_wait(p_full, j & 1);
                mbar_wait(&v_full[j & 1], (j >> 1) & 1);
                // wait for PV(j-1) BEFORE issuing PV(j): a parity wait must never fall two phases behind its barrier
                if (j >= 1) mbar_wait(o_full, (j - 1) & 1);
                tc_fence_after();
                issue_pv(j);
                if (j >= 1 && j + 1 < nkt) load_v(j + 1);  // the V stage of tile j-1 is free
            }
        }
    } else {
        // ===================== softmax: one query row per thread =====================
        const int row = threadIdx.x;
        const uint32_t t_s = tmem_base + ((uint32_t)(warp * 32) << 16);
        const uint32_t t_o = t_s + 64;
        const int i = qt * BT + row;
        int act_i = 0, sess_i = 0;
        if (i < p.L) {
            if (kind_uses_act<KIND>()) act_i = p.act[(long long)b * p.L + i];
            if (kind_uses_sess<KIND>()) sess_i = p.sess[(long long)b * p.L + i];
        }
        const int istart = (i / p.P) * p.P;
        const DropParams drop = drop_resolve(p.drop);
        unsigned allvalid = 0;  // bit t: every key of 128-key tile t is valid (CAUSAL tiles off the diagonal need no mask)
        if (KIND == MASK_CAUSAL) {
            for (int t = 0; t < p.k_tiles; ++t) allvalid |= (p.tflag[b * p.k_tiles + t] != 0 ? 1u : 0u) << t;
        }
        float m = -INFINITY, l = 0.f;
        for (int j = 0; j < nkt; ++j) {
            mbar_wait(&c_full[j & 3], (j >> 2) & 1);
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            uint32_t s[64];
            tmem_ld_32x32(t_s, s);
            tmem_ld_32x32(t_s + 32, s + 32);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(s_free);
            const int4* mk4 = reinterpret_cast<const int4*>(smem + S_OFF_C + (j & 3) * 512);
            // keys of this tile relative to the query tile's first row (only the diagonal band needs the j <= i test)
            const int cc0 = j * S_KT - qt * BT;
            const bool diag = kind_causal<KIND>() && (cc0 + S_KT - 1 > 0);
            const bool codes = (KIND != MASK_CAUSAL) || (((allvalid >> (j >> 1)) & 1u) == 0);
            float mx;
            if (codes) {
                if (diag) mx = mask_max64<KIND, true, true>(s, mk4, act_i, sess_i, cc0, row, j * S_KT, i, istart);
                else mx = mask_max64<KIND, false, true>(s, mk4, act_i, sess_i, cc0, row, j * S_KT, i, istart);
            } else if (diag) {
                mx = mask_max64<KIND, true, false>(s, mk4, act_i, sess_i, cc0, row, j * S_KT, i, istart);
            } else {
                float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int c = 0; c < 64; ++c) mx4[c & 3] = fmaxf(mx4[c & 3], __uint_as_float(s[c]));
                mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            }
            // lazy running max (log2 domain): raise it only from -inf or by more than 2^64
            const float m_tile = mx * p.scale_log2;
            float m_new = m;
            if (m == -INFINITY) m_new = m_tile;
            else if (m_tile > m + 64.f) m_new = m_tile;
            const bool rescale = (m != -INFINITY) && (m_new != m);
            if (rescale) l *= ex2_approx(m - m_new);
            const float m_old = m;
            m = m_new;
            const float neg_m = (m == -INFINITY) ? 0.f : -m;
            float sum[4] = {0.f, 0.f, 0.f, 0.f};
            if constexpr (DROP) {
#pragma unroll
                for (int cb = 0; cb < 4; ++cb) {
                    const uint4 rnd = drop_attn16(drop, (uint32_t)(b * p.n_q + h), (uint32_t)i, (uint32_t)(j * 4 + cb));
                    const uint32_t rw[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
                    for (int e = 0; e < 16; e += 2) {
                        const int c = cb * 16 + e;
                        const float p0 = ex2_approx(fmaf(__uint_as_float(s[c]), p.scale_log2, neg_m));
                        const float p1 = ex2_approx(fmaf(__uint_as_float(s[c + 1]), p.scale_log2, neg_m));
                        sum[(c >> 1) & 1] += p0;
                        sum[2 + ((c >> 1) & 1)] += p1;
                        const uint32_t w = rw[e >> 2];
                        const bool k0 = ((w >> (8 * (e & 3))) & 0xffu) >= drop.thresh;
                        const bool k1 = ((w >> (8 * (e & 3) + 8)) & 0xffu) >= drop.thresh;
                        s[c >> 1] = pack_bf16(k0 ? p0 : 0.f, k1 ? p1 : 0.f);
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < 64; c += 2) {
                    const float p0 = ex2_approx(fmaf(__uint_as_float(s[c]), p.scale_log2, neg_m));
                    const float p1 = ex2_approx(fmaf(__uint_as_float(s[c + 1]), p.scale_log2, neg_m));
                    sum[(c >> 1) & 1] += p0;
                    sum[2 + ((c >> 1) & 1)] += p1;
                    s[c >> 1] = pack_bf16(p0, p1);
                }
            }
            l += (sum[0] + sum[1]) + (sum[2] + sum[3]);
            if (j > 0) {
                mbar_wait(o_full, (j - 1) & 1);  // PV(j-1) done: O stable, P buffer free
                tc_fence_after();
                if (__any_sync(0xffffffffu, rescale)) rescale_o_row64(t_o, rescale ? ex2_approx(m_old - m_new) : 1.f);
            }
#pragma unroll
            for (int ch = 0; ch < 8; ++ch)
                sts128(sp + row * 128 + ((ch ^ (row & 7)) << 4), s[4 * ch], s[4 * ch + 1], s[4 * ch + 2], s[4 * ch + 3]);
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(p_full);
        }
        // ---- epilogue: O / l (or the V column mean on uniform rows) -> bf16 -> smem (the dead Q tile) -> TMA store
        mbar_wait(o_full, (nkt - 1) & 1);
        tc_fence_after();
        const bool uniform = !(l > 0.f);
        const float inv = uniform ? 0.f : (DROP ? drop.scale : 1.f) / l;
        const float* vm = p.vmean + ((long long)b * p.n_kv + g) * D;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            uint32_t o[32];
            tmem_ld_32x32(t_o + hh * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(o[8 * q + e]) * inv;
                if (uniform) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = vm[hh * 32 + 8 * q + e];
                }
                const int ch = hh * 4 + q;
                const bf16x8 ov = float_to_bf16x8(v);
                sts128(sq + row * 128 + ((ch ^ (row & 7)) << 4), ov.u[0], ov.u[1], ov.u[2], ov.u[3]);
            }
        }
        if (i < p.L) p.lse[((long long)b * p.n_q + h) * p.L + i] = uniform ? INFINITY : (m + log2f(l));
        tc_fence_before();
        fence_proxy_async();
        named_bar_sync(1, 128);
        if (threadIdx.x == 0) {
            tma_store_3d(&tmO, smem + S_OFF_Q, h * D, qt * BT, b);
            bulk_commit();
            bulk_wait0();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<128>(tmem_base);
}

// =================================================================================================================
// backward
// =================================================================================================================
constexpr int B_KV_STAGE = 2 * TILE_BYTES + 1024;         // K | V | ka[128] | ks[128]
constexpr int B_OFF_KV = 0;                               // 2 stages (items)
constexpr int B_QDO_STAGE = 2 * TILE_BYTES + 2048;        // Q | dO | lse[128] | dsum[128] | level[128] | session[128]
constexpr int B_OFF_QDO = 2 * B_KV_STAGE;                 // 2 stages
constexpr int B_OFF_P = B_OFF_QDO + 2 * B_QDO_STAGE;      // [128 q x 128 k] bf16 in two 64-key halves
constexpr int B_OFF_DS = B_OFF_P + 2 * TILE_BYTES;
constexpr int B_OFF_STG = B_OFF_DS + 2 * TILE_BYTES;      // 16 KB staging: dQ half tiles (fp32), dK / dV tiles (bf16)
constexpr int B_OFF_BAR = B_OFF_STG + TILE_BYTES;
constexpr int B_SMEM = B_OFF_BAR + 256 + 1024;
constexpr int B_THREADS = 768;
constexpr int DQ_TILE_FLOATS = BT * D;

struct BwdParams {
    int B, L, Lp, n_q, n_kv, P, q_tiles, k_tiles, total;
    const int* ka;
    const int* ks;
    const int* qa;
    const int* qs;       // [B, Lp] query-side behaviour level / session (zero past L)
    const float* lse_p;  // [B, n_q, Lp] log2 domain, +inf = uniform row or i >= L
    const float* dsum_p; // [B, n_q, Lp]
    const unsigned* uni_bits;  // [B]: bit qt = query tile qt holds a uniform row
    float scale, scale_log2, inv_L;
    float* dq_acc;       // [B, n_q, q_tiles][2 halves][128 rows][32 floats], 16-byte chunks XOR-swizzled by (row & 7)
    DropParams drop;     // the forward's attention-probability dropout, regenerated here
    Trace tr;
};

template <int KIND>
__device__ __forceinline__ void bwd_decode(int w, const BwdParams& p, int& b, int& g, int& kt, unsigned& qmask) {
    // Work order.  A "group" = (sequence, kv head); its k_tiles items share Q / dO tiles and dQ accumulator tiles.
    //  * full rounds: groups are dealt one per CTA and a CTA walks ITS group's key tiles 0..k_tiles-1 back to back (w
    //    advances by gridDim.x between its items): every CTA carries the same load, and the group's tiles stay in L2;
    //  * the last (n_groups mod gridDim.x) groups are dealt key-tile-major over all CTAs, heaviest key tile first (tile 0
    //    meets the most query tiles), so the tail of the launch is made of the light items.
    const int grid = (int)gridDim.x;
    const int n_groups = p.B * p.n_kv;
    const int full = (n_groups / grid) * grid;
    int grp;
    if (w < full * p.k_tiles) {
        const int r = w % (grid * p.k_tiles);
        kt = r / grid;
        grp = (w / (grid * p.k_tiles)) * grid + r % grid;
    } else {
        const int wr = w - full * p.k_tiles, rem = n_groups - full;
        kt = wr / rem;
        grp = full + wr % rem;
    }
    b = grp / p.n_kv;
    g = grp % p.n_kv;
    const unsigned all = (p.q_tiles >= 32) ? 0xffffffffu : ((1u << p.q_tiles) - 1u);
    if (kind_causal<KIND>()) qmask = ((all >> kt) << kt) | (p.uni_bits[b] & all);
    else qmask = all;
}
// step n of an item -> (head within the group, query tile): head-major, query tiles ascending
__device__ __forceinline__ void bwd_step(unsigned qmask, int n_per_head, int n, int& hh, int& qt) {
    hh = n >= n_per_head ? 1 : 0;
    const int k = n - hh * n_per_head;
    unsigned m = qmask;
    for (int x = 0; x < k; ++x) m &= m - 1;
    qt = __ffs(m) - 1;
}

// P = exp2(S * scale_log2 - lse) on allowed pairs, 0 elsewhere; uniform rows (lim_u > 0) get 1/L on every key column
// c < lim_u.  Branch-free per element (the tile variant is a template parameter): MODE 0 = below the diagonal / non-causal,
// 1 = diagonal tile, 2 = above the diagonal (only uniform rows are non-zero there).  `cc0` = first key column of this
// thread's half within the 128-key tile.
template <int KIND, int MODE>
__device__ __forceinline__ void bwd_p_tile(uint32_t (&s)[32], const int4* mk4, int act_i, int sess_i, int cc0, int row,
                                           int jbase, int i, int istart, float scale_log2, float lse_i, int lim_u,
                                           float inv_L) {
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
        int4 a4 = make_int4(0, 0, 0, 0), s4 = make_int4(0, 0, 0, 0);
        if (MODE != 2 && KIND != MASK_SESSION) a4 = mk4[c4];
        if (MODE != 2 && kind_uses_sess<KIND>()) s4 = mk4[32 + c4];
        const int av[4] = {a4.x, a4.y, a4.z, a4.w};
        const int sv4[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = c4 * 4 + e;
            float pv = 0.f;
            if (MODE != 2) {
                const bool ok = allow_tc<KIND, MODE == 1>(av[e], sv4[e], act_i, sess_i, cc0 + c, row, jbase + c, i, istart);
                const float x = fmaf(__uint_as_float(s[c]), scale_log2, -lse_i);
                pv = ex2_approx(ok ? x : -INFINITY);
            }
            pv = (c < lim_u) ? inv_L : pv;
            s[c] = __float_as_uint(pv);
        }
    }
}

// byte-wise x >= t over the four bytes of a word (SWAR): bit 7 of every result byte is the comparison
__device__ __forceinline__ uint32_t bytes_ge(uint32_t x, uint32_t t) {
    const uint32_t sum = (x & 0x7f7f7f7fu) + (0x80808080u - (t & 0x7fu) * 0x01010101u);   // bit 7: low 7 bits >= those of t
    return (t & 0x80u) ? (x & sum) : (x | sum);
}
// 0xffffffff when bit 7 of byte K of w is set, else 0 (one PRMT: sign-replicating byte select)
template <int K>
__device__ __forceinline__ uint32_t byte_sign_mask(uint32_t w) {
    return prmt(w, 0u, 0x8888u | (K * 0x1111u));
}

// With dropout (DROP): O = (P o Z) V, Z = keep / keep_prob.  dV += (P o Z)^T dO; dS = P o (Z o dP - dsum) with
// dsum = rowsum(dO o O) unchanged.  Z = zs * keep: the keep bits are byte sign flags (SWAR compare of the Philox bytes)
// expanded to AND masks by one PRMT each, zs multiplies dV once in its epilogue and enters dS through an FMA.
// Uniform rows carry no dropout (keep = all, their P is stored divided by zs), as in the forward.
template <int KIND, bool DROP>
__global__ void __launch_bounds__(B_THREADS, 1)
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                   const __grid_constant__ CUtensorMap tmdK, const __grid_constant__ CUtensorMap tmdV, BwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B_OFF_BAR);
    uint64_t* kv_full = bars + 0;    // [2]
    uint64_t* kv_free = bars + 2;    // [2]
    uint64_t* qdo_full = bars + 4;   // [2]
    uint64_t* qdo_free = bars + 6;   // [2]
    uint64_t* s_full = bars + 8;
    uint64_t* s_free = bars + 9;
    uint64_t* dp_full = bars + 10;
    uint64_t* pds_full = bars + 11;
    uint64_t* p_free = bars + 12;
    uint64_t* ds_free = bars + 13;
    uint64_t* dq_full = bars + 14;   // [2]: dQ lives in two TMEM buffers, so dQ(n) does not wait for the drain of dQ(n-1)
    uint64_t* dq_free = bars + 16;   // [2]
    uint64_t* dkv_full = bars + 18;
    uint64_t* dkv_free = bars + 19;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        prefetch_tmap(&tmQ);
        prefetch_tmap(&tmK);
        prefetch_tmap(&tmV);
        prefetch_tmap(&tmdO);
        prefetch_tmap(&tmdK);
        prefetch_tmap(&tmdV);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_free[s], 1);
            mbar_init(&qdo_full[s], 1);
            mbar_init(&qdo_free[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_free, 512);
        mbar_init(dp_full, 1);
        mbar_init(pds_full, 512);
        mbar_init(p_free, 1);
        mbar_init(ds_free, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&dq_full[s], 1);
            mbar_init(&dq_free[s], 128);
        }
        mbar_init(dkv_full, 1);
        mbar_init(dkv_free, 128);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // TMEM columns: S [0,128)  dP [128,256)  dV [256,320)  dK [320,384)  dQ [384,448) and [448,512)
    constexpr uint32_t T_S = 0, T_DP = 128, T_DV = 256, T_DK = 320, T_DQ = 384;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t item_n = 0, step_n = 0;
            for (int w = blockIdx.x; w < p.total; w += gridDim.x, ++item_n) {
                int b, g, kt;
                unsigned qmask;
                bwd_decode<KIND>(w, p, b, g, kt, qmask);
                const int nph = __popc(qmask), N = 2 * nph;
                const int ks = item_n & 1;
                uint8_t* skv = smem + B_OFF_KV + ks * B_KV_STAGE;
                mbar_wait(&kv_free[ks], ((item_n >> 1) & 1) ^ 1);
                mbar_expect_tx(&kv_full[ks], B_KV_STAGE);
                tma_load_3d(skv, &tmK, &kv_full[ks], g * D, kt * BT, b);
                tma_load_3d(skv + TILE_BYTES, &tmV, &kv_full[ks], g * D, kt * BT, b);
                bulk_load_1d(skv + 2 * TILE_BYTES, p.ka + (long long)b * p.Lp + kt * BT, 512, &kv_full[ks]);
                bulk_load_1d(skv + 2 * TILE_BYTES + 512, p.ks + (long long)b * p.Lp + kt * BT, 512, &kv_full[ks]);
                for (int n = 0; n < N; ++n, ++step_n) {
                    int hh, qt;
                    bwd_step(qmask, nph, n, hh, qt);
                    const int st = step_n & 1;
                    mbar_wait(&qdo_free[st], ((step_n >> 1) & 1) ^ 1);
                    uint8_t* sq = smem + B_OFF_QDO + st * B_QDO_STAGE;
                    mbar_expect_tx(&qdo_full[st], B_QDO_STAGE);
                    tma_load_3d(sq, &tmQ, &qdo_full[st], (2 * g + hh) * D, qt * BT, b);
                    tma_load_3d(sq + TILE_BYTES, &tmdO, &qdo_full[st], (2 * g + hh) * D, qt * BT, b);
                    const long long ro = ((long long)b * p.n_q + 2 * g + hh) * p.Lp + qt * BT;
                    bulk_load_1d(sq + 2 * TILE_BYTES, p.lse_p + ro, 512, &qdo_full[st]);
                    bulk_load_1d(sq + 2 * TILE_BYTES + 512, p.dsum_p + ro, 512, &qdo_full[st]);
                    bulk_load_1d(sq + 2 * TILE_BYTES + 1024, p.qa + (long long)b * p.Lp + qt * BT, 512, &qdo_full[st]);
                    bulk_load_1d(sq + 2 * TILE_BYTES + 1536, p.qs + (long long)b * p.Lp + qt * BT, 512, &qdo_full[st]);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t item_n = 0, step_n = 0;
            int tn = 0;
            const uint32_t sp = smem_u32(smem + B_OFF_P), sds = smem_u32(smem + B_OFF_DS);
            for (int w = blockIdx.x; w < p.total; w += gridDim.x, ++item_n) {
                int b, g, kt;
                unsigned qmask;
                bwd_decode<KIND>(w, p, b, g, kt, qmask);
                const int N = 2 * __popc(qmask);
                const int ks = item_n & 1;
                const uint32_t sk = smem_u32(smem + B_OFF_KV + ks * B_KV_STAGE), sv = sk + TILE_BYTES;
                mbar_wait(&kv_full[ks], (item_n >> 1) & 1);
                trace_pt(p.tr, 0, tn, 1);
                {
                    const int st = step_n & 1;
                    mbar_wait(&qdo_full[st], (step_n >> 1) & 1);
                    trace_pt(p.tr, 0, tn, 2);
                    if (step_n > 0) mbar_wait(s_free, (step_n - 1) & 1);  // S read by the previous item's last step
                    tc_fence_after();
                    const uint32_t sq = smem_u32(smem + B_OFF_QDO + st * B_QDO_STAGE);
                    issue_nt_128x128x64(tmem_base + T_S, sq, sk);
                    umma_commit(s_full);
                    // dP is free: the previous step's pds_full (waited on below) follows its last dP read
                    issue_nt_128x128x64(tmem_base + T_DP, sq + TILE_BYTES, sv);
                    umma_commit(dp_full);
                }
                for (int n = 0; n < N; ++n, ++step_n) {
                    const int st = step_n & 1;
                    const uint32_t sq = smem_u32(smem + B_OFF_QDO + st * B_QDO_STAGE);
                    const uint32_t sqn = smem_u32(smem + B_OFF_QDO + (st ^ 1) * B_QDO_STAGE);
                    if (n + 1 < N) {
                        mbar_wait(&qdo_full[st ^ 1], ((step_n + 1) >> 1) & 1);
                        mbar_wait(s_free, step_n & 1);
                        tc_fence_after();
                        issue_nt_128x128x64(tmem_base + T_S, sqn, sk);
                        umma_commit(s_full);
                    }
                    trace_pt(p.tr, 0, tn, 3);
                    mbar_wait(pds_full, step_n & 1);
                    trace_pt(p.tr, 0, tn, 4);
                    if (n == 0 && item_n > 0) mbar_wait(dkv_free, (item_n - 1) & 1);  // dK/dV of the previous item drained
                    tc_fence_after();
                    if (n + 1 < N) {  // dP of the next step first: its buffer is free (pds_full follows the dP read)
                        issue_nt_128x128x64(tmem_base + T_DP, sqn + TILE_BYTES, sv);     // dP(n+1) = dO V^T
                        umma_commit(dp_full);
                    }
                    issue_tn_128x64x128(tmem_base + T_DV, sp, sq + TILE_BYTES, n > 0);   // dV += P^T dO
                    umma_commit(p_free);
                    issue_tn_128x64x128(tmem_base + T_DK, sds, sq, n > 0);               // dK += dS^T Q
                    const uint32_t db = step_n & 1, du = step_n >> 1;
                    if (du > 0) {
                        mbar_wait(&dq_free[db], (du - 1) & 1);
                        tc_fence_after();
                    }
                    trace_pt(p.tr, 0, tn, 5);
                    umma_commit(&qdo_free[st]);   // Q / dO of this step are dead once dK is done: dQ reads dS and K only
                    issue_nn_128x64x128(tmem_base + T_DQ + db * 64, sds, sk, false);     // dQ = dS K
                    umma_commit(&dq_full[db]);
                    umma_commit(ds_free);
                    if (n + 1 == N) {
                        umma_commit(dkv_full);
                        umma_commit(&kv_free[ks]);
                    }
                }
            }
        }
    } else if (warp >= 4 && warp < 20) {
        // ===================== softmax / dS warps: 16 warps, warpgroup k (warps 4+4k..7+4k) owns key columns [32k, 32k+32)
        // of the tile; one query row per thread =====================
        const int wgi = (warp - 4) >> 2;
        const int wq = warp & 3;
        const int row = wq * 32 + lane;
        const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
        const uint32_t t_s = tmem_base + lane_off + T_S + wgi * 32;
        const uint32_t t_dp = tmem_base + lane_off + T_DP + wgi * 32;
        // P / dS live as two 64-key halves; this thread writes 16-byte chunks [4 (wgi & 1), 4 (wgi & 1) + 4) of its row
        const uint32_t sP = smem_u32(smem + B_OFF_P + (wgi >> 1) * TILE_BYTES) + row * 128;
        const uint32_t sDS = smem_u32(smem + B_OFF_DS + (wgi >> 1) * TILE_BYTES) + row * 128;
        const int ch0 = (wgi & 1) * 4;
        const DropParams drop = drop_resolve(p.drop);
        uint32_t item_n = 0, step_n = 0;
        int tn = 0;
        Trace tr = p.tr;
        if (row != 0 || wgi != 0) tr.buf = nullptr;
        for (int w = blockIdx.x; w < p.total; w += gridDim.x, ++item_n) {
            int b, g, kt;
            unsigned qmask;
            bwd_decode<KIND>(w, p, b, g, kt, qmask);
            const int nph = __popc(qmask), N = 2 * nph;
            const int ks = item_n & 1;
            mbar_wait(&kv_full[ks], (item_n >> 1) & 1);  // key codes
            const int4* mk4 = reinterpret_cast<const int4*>(smem + B_OFF_KV + ks * B_KV_STAGE + 2 * TILE_BYTES) + wgi * 8;
            const int jbase = kt * BT + wgi * 32;
            for (int n = 0; n < N; ++n, ++step_n) {
                int hh, qt;
                bwd_step(qmask, nph, n, hh, qt);
                const int i = qt * BT + row;
                // per-row scalars travel with the Q / dO stage (padded copies: lse = +inf, dsum = 0 past L)
                const int st = step_n & 1;
                mbar_wait(&qdo_full[st], (step_n >> 1) & 1);
                const float* rowf = reinterpret_cast<const float*>(smem + B_OFF_QDO + st * B_QDO_STAGE + 2 * TILE_BYTES);
                const float lse_i = rowf[row], dsum_i = rowf[128 + row];
                const int act_i = reinterpret_cast<const int*>(rowf)[256 + row];
                const int sess_i = reinterpret_cast<const int*>(rowf)[384 + row];
                const bool uni = (i < p.L) && (lse_i == INFINITY);
                const int istart = (i / p.P) * p.P;
                const bool diag = kind_causal<KIND>() && (qt == kt);
                const bool above = kind_causal<KIND>() && (qt < kt);  // only uniform rows reach keys above the diagonal
                trace_pt(tr, 1, tn, 20);
                mbar_wait(s_full, step_n & 1);
                tc_fence_after();
                trace_pt(tr, 1, tn, 21);
                uint32_t s[32];
                tmem_ld_32x32(t_s, s);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(s_free);
                trace_pt(tr, 1, tn, 22);
                // ---- P
                {
                    const int lim_u = uni ? (p.L - jbase) : 0;
                    if (above)
                        bwd_p_tile<KIND, 2>(s, mk4, act_i, sess_i, wgi * 32, row, jbase, i, istart, p.scale_log2, lse_i, lim_u, p.inv_L);
                    else if (diag)
                        bwd_p_tile<KIND, 1>(s, mk4, act_i, sess_i, wgi * 32, row, jbase, i, istart, p.scale_log2, lse_i, lim_u, p.inv_L);
                    else
                        bwd_p_tile<KIND, 0>(s, mk4, act_i, sess_i, wgi * 32, row, jbase, i, istart, p.scale_log2, lse_i, lim_u, p.inv_L);
                }
                trace_pt(tr, 1, tn, 23);
                // keep flags: bit 7 of byte (c & 3) of kw[c >> 2] <-> (query, key column c) is kept
                uint32_t kw[8];
                const float zs = DROP ? drop.scale : 1.f;
                const float zrow = (DROP && !uni) ? zs : 1.f;        // factor of dP in dS
                const float pmul = (DROP && uni) ? 1.f / zs : 1.f;   // dV is multiplied by zs in its epilogue
                if constexpr (DROP) {
#pragma unroll
                    for (int cb = 0; cb < 2; ++cb) {
                        const uint4 rnd = drop_attn16(drop, (uint32_t)(b * p.n_q + 2 * g + hh), (uint32_t)i,
                                                      (uint32_t)((jbase >> 4) + cb));
                        kw[4 * cb + 0] = uni ? 0x80808080u : bytes_ge(rnd.x, drop.thresh);
                        kw[4 * cb + 1] = uni ? 0x80808080u : bytes_ge(rnd.y, drop.thresh);
                        kw[4 * cb + 2] = uni ? 0x80808080u : bytes_ge(rnd.z, drop.thresh);
                        kw[4 * cb + 3] = uni ? 0x80808080u : bytes_ge(rnd.w, drop.thresh);
                    }
                }
                if (step_n > 0) mbar_wait(p_free, (step_n - 1) & 1);
                trace_pt(tr, 1, tn, 24);
                // P leaves as bf16 pairs: the masked (and, on uniform rows, pre-divided) copy goes to smem for dV, the
                // plain copy stays in 16 registers for dS — the fp32 tile dies here, which is what keeps this loop
                // inside the 80-register budget of a 768-thread CTA
                uint32_t pp[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t pk[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int c = q * 8 + 2 * e;
                        uint32_t p0 = s[c], p1 = s[c + 1];
                        pp[c >> 1] = pack_bf16(__uint_as_float(p0), __uint_as_float(p1));
                        if constexpr (DROP) {
                            p0 &= (c & 2) ? byte_sign_mask<2>(kw[c >> 2]) : byte_sign_mask<0>(kw[c >> 2]);
                            p1 &= (c & 2) ? byte_sign_mask<3>(kw[c >> 2]) : byte_sign_mask<1>(kw[c >> 2]);
                            pk[e] = pack_bf16(__uint_as_float(p0) * pmul, __uint_as_float(p1) * pmul);
                        } else {
                            pk[e] = pp[c >> 1];
                        }
                    }
                    sts128(sP + (((ch0 + q) ^ (row & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
                }
                // ---- dS = P o (zs * keep o dP - dsum), dP read in two 16-column halves
                trace_pt(tr, 1, tn, 25);
                mbar_wait(dp_full, step_n & 1);
                tc_fence_after();
                trace_pt(tr, 1, tn, 26);
                if (step_n > 0) mbar_wait(ds_free, (step_n - 1) & 1);
                trace_pt(tr, 1, tn, 27);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t dp[16];
                    tmem_ld_32x16(t_dp + half * 16, dp);
                    tmem_ld_wait();
#pragma unroll
                    for (int q2 = 0; q2 < 2; ++q2) {
                        uint32_t pk[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int cl = q2 * 8 + 2 * e;          // column within the half
                            const int c = half * 16 + cl;
                            uint32_t g0 = dp[cl], g1 = dp[cl + 1];
                            if constexpr (DROP) {
                                g0 &= (c & 2) ? byte_sign_mask<2>(kw[c >> 2]) : byte_sign_mask<0>(kw[c >> 2]);
                                g1 &= (c & 2) ? byte_sign_mask<3>(kw[c >> 2]) : byte_sign_mask<1>(kw[c >> 2]);
                            }
                            const float2 pv = unpack_bf16(pp[c >> 1]);
                            const float d0 = pv.x * fmaf(__uint_as_float(g0), zrow, -dsum_i);
                            const float d1 = pv.y * fmaf(__uint_as_float(g1), zrow, -dsum_i);
                            pk[e] = pack_bf16(d0, d1);
                        }
                        sts128(sDS + (((ch0 + half * 2 + q2) ^ (row & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
                    }
                }
                trace_pt(tr, 1, tn, 28);
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(pds_full);
                trace_pt(tr, 1, tn, 29);
            }
        }
    } else if (warp >= 20) {
        // ===================== drain warps: dQ tiles -> TMA fp32 reduce-add; dK / dV -> bf16 TMA store ================
        const int wq = warp & 3;
        const int row = wq * 32 + lane;
        const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
        const uint32_t t_dq = tmem_base + lane_off + T_DQ;
        uint8_t* stg = smem + B_OFF_STG;
        const uint32_t stg_row = smem_u32(stg) + row * 128;
        uint32_t item_n = 0, step_n = 0;
        int tn = 0;
        Trace tr = p.tr;
        if (row != 0) tr.buf = nullptr;
        for (int w = blockIdx.x; w < p.total; w += gridDim.x, ++item_n) {
            int b, g, kt;
            unsigned qmask;
            bwd_decode<KIND>(w, p, b, g, kt, qmask);
            const int nph = __popc(qmask), N = 2 * nph;
            for (int n = 0; n < N; ++n, ++step_n) {
                int hh, qt;
                bwd_step(qmask, nph, n, hh, qt);
                trace_pt(tr, 2, tn, 40);
                const uint32_t db = step_n & 1, du = step_n >> 1;
                mbar_wait(&dq_full[db], du & 1);
                tc_fence_after();
                trace_pt(tr, 2, tn, 41);
                float* dst = p.dq_acc + (((long long)b * p.n_q + 2 * g + hh) * p.q_tiles + qt) * DQ_TILE_FLOATS;
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    uint32_t o[32];
                    tmem_ld_32x32(t_dq + db * 64 + half * 32, o);
                    tmem_ld_wait();
                    if (half == 1) {
                        tc_fence_before();
                        mbar_arrive(&dq_free[db]);
                    }
                    if (row == 0) bulk_wait_read0();  // the previous bulk op has finished reading the staging buffer
                    trace_pt(tr, 2, tn, 42 + half);
                    named_bar_sync(3, 128);
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch)
                        sts128(stg_row + ((ch ^ (row & 7)) << 4), __float_as_uint(__uint_as_float(o[4 * ch]) * p.scale),
                               __float_as_uint(__uint_as_float(o[4 * ch + 1]) * p.scale),
                               __float_as_uint(__uint_as_float(o[4 * ch + 2]) * p.scale),
                               __float_as_uint(__uint_as_float(o[4 * ch + 3]) * p.scale));
                    fence_proxy_async();
                    named_bar_sync(3, 128);
                    if (row == 0) {
                        bulk_reduce_add_f32(dst + half * (DQ_TILE_FLOATS / 2), stg, DQ_TILE_FLOATS * 2);
                        bulk_commit();
                    }
                }
            }
            // ---- item epilogue: dK then dV -> bf16 -> staging -> TMA store (rows past L are clipped)
            mbar_wait(dkv_full, item_n & 1);
            tc_fence_after();
#pragma unroll 1
            for (int which = 0; which < 2; ++which) {
                const uint32_t t_acc = tmem_base + lane_off + (which == 0 ? T_DK : T_DV);
                const float mul = (which == 0) ? p.scale : (DROP ? p.drop.scale : 1.f);   // dV = zs * (P o keep)^T dO
                uint32_t pk[32];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t o[32];
                    tmem_ld_32x32(t_acc + half * 32, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        pk[half * 16 + k] = pack_bf16(__uint_as_float(o[2 * k]) * mul, __uint_as_float(o[2 * k + 1]) * mul);
                }
                if (which == 1) {
                    tc_fence_before();
                    mbar_arrive(dkv_free);
                }
                if (row == 0) bulk_wait_read0();
                named_bar_sync(3, 128);
#pragma unroll
                for (int ch = 0; ch < 8; ++ch)
                    sts128(stg_row + ((ch ^ (row & 7)) << 4), pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
                fence_proxy_async();
                named_bar_sync(3, 128);
                if (row == 0) {
                    tma_store_3d(which == 0 ? &tmdK : &tmdV, stg, g * D, kt * BT, b);
                    bulk_commit();
                }
            }
        }
        if (row == 0) bulk_wait0();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// Padded per-row inputs of the backward: dsum_p[b,h,i] = sum_d dO*O and lse_p[b,h,i] = lse (rows i >= L: 0 / +inf);
// uni_bits[b] |= 1 << (i / 128) for uniform rows (lse = +inf; head 0 decides: the mask is head-independent)
__global__ void attn_tc_bwd_prep_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o, long long ld_o, int B,
                                        int L, int Lp, int n_q, const float* __restrict__ lse, float* __restrict__ dsum_p,
                                        float* __restrict__ lse_p, unsigned* __restrict__ uni_bits) {
    // grid = (groups of one sequence / 32, B); 8 threads per (padded row, head) group, 32-bit index math only
    const int b = blockIdx.y;
    const int sub = threadIdx.x & 7;
    const int grp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 3);   // (row i, head h) of sequence b
    const bool live = grp < Lp * n_q;
    const int i = live ? grp / n_q : 0;
    const int h = live ? grp - i * n_q : 0;
    const bool real = live && i < L;
    const long long row = real ? (long long)b * L + i : 0;
    float a[8], d[8];
    bf16x8_to_float(*reinterpret_cast<const bf16x8*>(o + row * ld_o + h * D + sub * 8), a);
    bf16x8_to_float(*reinterpret_cast<const bf16x8*>(d_o + row * ld_o + h * D + sub * 8), d);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k] * d[k];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (live && sub == 0) {
        const long long pi = ((long long)b * n_q + h) * Lp + i;
        float ls = INFINITY;
        if (real) {
            ls = lse[((long long)b * n_q + h) * L + i];
            if (h == 0 && ls == INFINITY) atomicOr(&uni_bits[b], 1u << (i / BT));
        }
        dsum_p[pi] = real ? s : 0.f;
        lse_p[pi] = ls;
    }
}

// dq_acc (tile-major, swizzled fp32) -> dq bf16 [B*L, ld_d] + h*64.  One thread per 8 consecutive head-dim elements.
__global__ void attn_tc_dq_convert_kernel(const float* __restrict__ acc, int B, int L, int n_q, int q_tiles,
                                          bf16* __restrict__ dq, long long ld_d) {
    const long long total = (long long)B * n_q * L * 8;
    for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(x & 7);
        long long r = x >> 3;
        const int h = (int)(r % n_q);
        r /= n_q;
        const int i = (int)(r % L), b = (int)(r / L);
        const int qt = i >> 7, row = i & 127;
        const float* tile = acc + (((long long)b * n_q + h) * q_tiles + qt) * DQ_TILE_FLOATS + (c8 >> 2) * (DQ_TILE_FLOATS / 2) +
                            row * 32;
        const int ch = (c8 & 3) * 2;
        const float4 v0 = *reinterpret_cast<const float4*>(tile + ((ch ^ (row & 7)) << 2));
        const float4 v1 = *reinterpret_cast<const float4*>(tile + (((ch + 1) ^ (row & 7)) << 2));
        const float f[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        *reinterpret_cast<bf16x8*>(dq + ((long long)b * L + i) * ld_d + h * D + c8 * 8) = float_to_bf16x8(f);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// bf16 [B][L][cols] view (row stride ld elements, sequence stride L*ld); box = [1][128 rows][64 cols], SWIZZLE_128B
int make_tmap_seq(CUtensorMap* m, const void* base, int B, int L, int cols, long long ld, int box_rows = 128) {
    EncodeTiledFn fn = encode_fn();
    GAMER_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
    GAMER_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 2) % 16 == 0,
                  "attention operands must be 16-byte aligned (ld=%lld)", ld);
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)L, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)L * ld * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GAMER_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (attention) failed with %d (B=%d L=%d cols=%d ld=%lld)", (int)r, B,
                  L, cols, ld);
    return 0;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    }
    return n;
}

inline long long align256(long long x) { return (x + 255) / 256 * 256; }

struct MetaLayout {
    int Lp, k_tiles;
    long long off_ka, off_ks, off_qa, off_qs, off_flag, bytes;
};
MetaLayout meta_layout(int B, int L) {
    MetaLayout m;
    m.k_tiles = (L + BT - 1) / BT;
    m.Lp = m.k_tiles * BT;
    m.off_ka = 0;
    const long long arr = align256((long long)B * m.Lp * 4);
    m.off_ks = arr;
    m.off_qa = 2 * arr;
    m.off_qs = 3 * arr;
    m.off_flag = 4 * arr;
    m.bytes = m.off_flag + align256((long long)B * m.k_tiles * 4);
    return m;
}

int build_meta(int kind, const int* am, const int* act, const int* sess, int B, int L, uint8_t* ws, const MetaLayout& ml,
               cudaStream_t stream) {
    const int* act_in = (kind == MASK_MULTI_CROSS || kind == MASK_SESSION_CROSS) ? act : nullptr;
    const int* sess_in = (kind == MASK_SESSION || kind == MASK_SESSION_CROSS) ? sess : nullptr;
    attn_meta_kernel<<<B * ml.k_tiles, BT, 0, stream>>>(am, act_in, sess_in, L, ml.Lp, ml.k_tiles,
                                                        reinterpret_cast<int*>(ws + ml.off_ka),
                                                        reinterpret_cast<int*>(ws + ml.off_ks),
                                                        reinterpret_cast<int*>(ws + ml.off_qa),
                                                        reinterpret_cast<int*>(ws + ml.off_qs),
                                                        reinterpret_cast<int*>(ws + ml.off_flag));
    GAMER_LAUNCH_CHECK();
    return 0;
}

template <int KIND>
int launch_fwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to,
               const FwdParams& p, cudaStream_t stream) {
    static bool cfg = false;
    if (!cfg) {
        GAMER_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM));
        cfg = true;
    }
    const int grid = p.total < sm_count() ? p.total : sm_count();
    attn_tc_fwd_kernel<KIND><<<grid, F_THREADS, F_SMEM, stream>>>(tq, tk, tv, to, p);
    GAMER_LAUNCH_CHECK();
    return 0;
}

template <int KIND, bool DROP>
int launch_fwd_small_t(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to,
                       const FwdParams& p, cudaStream_t stream) {
    static bool cfg = false;
    if (!cfg) {
        GAMER_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_small_kernel<KIND, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, S_SMEM));
        cfg = true;
    }
    attn_fwd_small_kernel<KIND, DROP><<<p.B * p.n_q * p.q_tiles, S_THREADS, S_SMEM, stream>>>(tq, tk, tv, to, p);
    GAMER_LAUNCH_CHECK();
    return 0;
}
template <int KIND>
int launch_fwd_small(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to,
                     const FwdParams& p, cudaStream_t stream) {
    return p.drop.thresh ? launch_fwd_small_t<KIND, true>(tq, tk, tv, to, p, stream)
                         : launch_fwd_small_t<KIND, false>(tq, tk, tv, to, p, stream);
}

template <int KIND, bool DROP>
int launch_bwd_t(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
                 const CUtensorMap& tdk, const CUtensorMap& tdv, const BwdParams& p, cudaStream_t stream) {
    static bool cfg = false;
    if (!cfg) {
        GAMER_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel<KIND, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM));
        cfg = true;
    }
    const int grid = p.total < sm_count() ? p.total : sm_count();
    attn_tc_bwd_kernel<KIND, DROP><<<grid, B_THREADS, B_SMEM, stream>>>(tq, tk, tv, tdo, tdk, tdv, p);
    GAMER_LAUNCH_CHECK();
    return 0;
}
template <int KIND>
int launch_bwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
               const CUtensorMap& tdk, const CUtensorMap& tdv, const BwdParams& p, cudaStream_t stream) {
    return p.drop.thresh ? launch_bwd_t<KIND, true>(tq, tk, tv, tdo, tdk, tdv, p, stream)
                         : launch_bwd_t<KIND, false>(tq, tk, tv, tdo, tdk, tdv, p, stream);
}

}  // namespace

// debug hook: subsequent forward launches record a timeline of CTA 0 into buf (3 roles x cap x (tag, clock) int64 pairs)
extern "C" int gamer_attn_set_trace(void* buf, int cap) {
    g_trace.buf = reinterpret_cast<long long*>(buf);
    g_trace.cap = cap;
    return 0;
}

bool attn_tc_supported(int L, int n_q, int n_kv, int head_dim) {
    return head_dim == D && n_kv > 0 && n_q == 2 * n_kv && L >= 1 && (L + BT - 1) / BT <= 32;
}

long long attn_tc_fwd_ws_bytes(int B, int L) { return meta_layout(B, L).bytes; }

int attn_tc_fwd(const void* q, const void* k, const void* v, long long ld, int B, int L, int n_q, int n_kv, int kind, int P,
                const int* am, const int* act, const int* sess, float scale, const float* vmean, void* ws, void* o,
                long long ld_o, float* lse, const gamer_dropout_t* drop, cudaStream_t stream) {
    const MetaLayout ml = meta_layout(B, L);
    uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
    if (int e = build_meta(kind, am, act, sess, B, L, w8, ml, stream)) return e;
    static int use_ws = -1;  // GAMER_ATTN_FWD_WS=1: the warp-specialised GQA ping-pong kernel instead of the small-CTA one
    if (use_ws < 0) {
        const char* e = getenv("GAMER_ATTN_FWD_WS");
        use_ws = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    const int kv_box = use_ws ? 128 : S_KT;
    CUtensorMap tq, tk, tv, to;
    if (int e = make_tmap_seq(&tq, q, B, L, n_q * D, ld)) return e;
    if (int e = make_tmap_seq(&tk, k, B, L, n_kv * D, ld, kv_box)) return e;
    if (int e = make_tmap_seq(&tv, v, B, L, n_kv * D, ld, kv_box)) return e;
    if (int e = make_tmap_seq(&to, o, B, L, n_q * D, ld_o)) return e;
    FwdParams p{};
    p.B = B; p.L = L; p.Lp = ml.Lp; p.n_q = n_q; p.n_kv = n_kv; p.P = P;
    p.q_tiles = ml.k_tiles; p.k_tiles = ml.k_tiles; p.total = B * n_kv * ml.k_tiles;
    p.ka = reinterpret_cast<const int*>(w8 + ml.off_ka);
    p.ks = reinterpret_cast<const int*>(w8 + ml.off_ks);
    p.tflag = reinterpret_cast<const int*>(w8 + ml.off_flag);
    p.act = act; p.sess = sess; p.scale_log2 = scale * 1.4426950408889634f; p.vmean = vmean; p.lse = lse;
    p.tr = g_trace;
    p.drop = make_drop(drop, 8);
    GAMER_REQUIRE(!(use_ws && p.drop.thresh), "attention dropout is implemented in the small-CTA forward kernel (unset GAMER_ATTN_FWD_WS)");
    if (!use_ws) {
        switch (kind) {
            case 0: return launch_fwd_small<0>(tq, tk, tv, to, p, stream);
            case 1: return launch_fwd_small<1>(tq, tk, tv, to, p, stream);
            case 2: return launch_fwd_small<2>(tq, tk, tv, to, p, stream);
            default: return launch_fwd_small<3>(tq, tk, tv, to, p, stream);
        }
    }
    switch (kind) {
        case 0: return launch_fwd<0>(tq, tk, tv, to, p, stream);
        case 1: return launch_fwd<1>(tq, tk, tv, to, p, stream);
        case 2: return launch_fwd<2>(tq, tk, tv, to, p, stream);
        default: return launch_fwd<3>(tq, tk, tv, to, p, stream);
    }
}

struct BwdLayout {
    MetaLayout ml;
    long long off_dsum, off_lse, off_uni, off_acc, acc_bytes, bytes;
};
static BwdLayout bwd_layout(int B, int L, int n_q) {
    BwdLayout bl;
    bl.ml = meta_layout(B, L);
    bl.off_dsum = bl.ml.bytes;
    bl.off_lse = bl.off_dsum + align256((long long)B * n_q * bl.ml.Lp * 4);
    bl.off_uni = bl.off_lse + align256((long long)B * n_q * bl.ml.Lp * 4);
    bl.off_acc = bl.off_uni + align256((long long)B * 4);
    bl.acc_bytes = (long long)B * n_q * bl.ml.k_tiles * DQ_TILE_FLOATS * 4;
    bl.bytes = bl.off_acc + align256(bl.acc_bytes);
    return bl;
}

long long attn_tc_bwd_ws_bytes(int B, int L, int n_q) { return bwd_layout(B, L, n_q).bytes; }

int attn_tc_bwd(const void* q, const void* k, const void* v, long long ld, int B, int L, int n_q, int n_kv, int kind, int P,
                const int* am, const int* act, const int* sess, float scale, const void* o, const void* d_o, long long ld_o,
                const float* lse, void* ws, void* dq, void* dk, void* dv, long long ld_d, const gamer_dropout_t* drop,
                cudaStream_t stream) {
    const BwdLayout bl = bwd_layout(B, L, n_q);
    uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
    if (int e = build_meta(kind, am, act, sess, B, L, w8, bl.ml, stream)) return e;
    float* dsum = reinterpret_cast<float*>(w8 + bl.off_dsum);
    float* lse_p = reinterpret_cast<float*>(w8 + bl.off_lse);
    unsigned* uni = reinterpret_cast<unsigned*>(w8 + bl.off_uni);
    float* acc = reinterpret_cast<float*>(w8 + bl.off_acc);
    // uni bits and the dQ accumulator are adjacent: one memset
    GAMER_CHECK_CUDA(cudaMemsetAsync(uni, 0, (size_t)(bl.off_acc - bl.off_uni) + (size_t)bl.acc_bytes, stream));
    {
        const dim3 grid((bl.ml.Lp * n_q * 8 + 255) / 256, B);
        attn_tc_bwd_prep_kernel<<<grid, 256, 0, stream>>>(
            reinterpret_cast<const bf16*>(o), reinterpret_cast<const bf16*>(d_o), ld_o, B, L, bl.ml.Lp, n_q, lse, dsum, lse_p,
            uni);
        GAMER_LAUNCH_CHECK();
    }
    CUtensorMap tq, tk, tv, tdo, tdk, tdv;
    if (int e = make_tmap_seq(&tq, q, B, L, n_q * D, ld)) return e;
    if (int e = make_tmap_seq(&tk, k, B, L, n_kv * D, ld)) return e;
    if (int e = make_tmap_seq(&tv, v, B, L, n_kv * D, ld)) return e;
    if (int e = make_tmap_seq(&tdo, d_o, B, L, n_q * D, ld_o)) return e;
    if (int e = make_tmap_seq(&tdk, dk, B, L, n_kv * D, ld_d)) return e;
    if (int e = make_tmap_seq(&tdv, dv, B, L, n_kv * D, ld_d)) return e;
    BwdParams p{};
    p.B = B; p.L = L; p.Lp = bl.ml.Lp; p.n_q = n_q; p.n_kv = n_kv; p.P = P;
    p.q_tiles = bl.ml.k_tiles; p.k_tiles = bl.ml.k_tiles; p.total = B * n_kv * bl.ml.k_tiles;
    p.ka = reinterpret_cast<const int*>(w8 + bl.ml.off_ka);
    p.ks = reinterpret_cast<const int*>(w8 + bl.ml.off_ks);
    p.qa = reinterpret_cast<const int*>(w8 + bl.ml.off_qa);
    p.qs = reinterpret_cast<const int*>(w8 + bl.ml.off_qs);
    p.lse_p = lse_p; p.dsum_p = dsum; p.uni_bits = uni;
    p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f; p.inv_L = 1.0f / (float)L; p.dq_acc = acc;
    p.tr = g_trace;
    p.drop = make_drop(drop, 8);
    int e;
    switch (kind) {
        case 0: e = launch_bwd<0>(tq, tk, tv, tdo, tdk, tdv, p, stream); break;
        case 1: e = launch_bwd<1>(tq, tk, tv, tdo, tdk, tdv, p, stream); break;
        case 2: e = launch_bwd<2>(tq, tk, tv, tdo, tdk, tdv, p, stream); break;
        default: e = launch_bwd<3>(tq, tk, tv, tdo, tdk, tdv, p, stream); break;
    }
    if (e) return e;
    {
        const long long total = (long long)B * n_q * L * 8;
        const long long blocks = (total + 255) / 256;
        attn_tc_dq_convert_kernel<<<(int)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, stream>>>(
            acc, B, L, n_q, bl.ml.k_tiles, reinterpret_cast<bf16*>(dq), ld_d);
        GAMER_LAUNCH_CHECK();
    }
    return 0;
}
