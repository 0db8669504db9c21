// K6 on the 5th-gen tensor cores: session / behaviour-masked attention, forward and fused backward, GQA 2:1, head_dim 64.
//
// Every matrix product (QK^T, PV, dO V^T, P^T dO, dS^T Q, dS K) is a tcgen05.mma with TMA-staged SWIZZLE_128B operands and
// TMEM accumulators; the softmax / mask / dS arithmetic runs in warps that read the score tiles from TMEM, one query
// row per thread.  Both kernels are bound by the instruction issue of those warps (and, behind that, MUFU.EX2 and the
// shared-memory port), so the per-element work is kept to the minimum:
//
//   * packed fp32 arithmetic (FFMA2 / FADD2 / FMUL2) and three-input max on register pairs;
//   * the mask predicate is classified per (warp, 32-key block) from per-block key summaries and per-warp query ranges:
//     FULL blocks run mask-free, SKIP blocks produce zeros without touching the MUFU, only boundary blocks (the causal
//     diagonal, partially padded blocks, behaviour-level tests) evaluate the predicate per element;
//   * dropout keep flags are one 32-bit word per (query, 32-key block), derived bit-parallel from eight Philox words (a
//     borrow chain of eight LOP3 compares the per-lane random bytes with the threshold).  The forward applies them to the
//     packed bf16 P with byte-sign PRMT masks and STORES the words; the backward reads them back (2 KB per 128x128 tile,
//     travelling with the Q / dO stage) instead of regenerating them;
//   * forward: P never visits shared memory — it is written to TMEM (tcgen05.st, aliasing the S columns) and consumed by
//     the PV MMA as its TMEM A operand.
//
//   forward  : CTA = (sequence, query head, 128-query tile), 64-key tiles, three CTAs per SM; one issuer thread (TMA + MMA)
//              and four softmax warps.  Online softmax in the exp2 domain with a lazy rescale (the running max is only
//              raised from -inf or when it would grow by more than 2^64; P stays representable in bf16).
//   backward : persistent CTA walks items (sequence, kv head, 128-key tile); per item it loops over (head of the group,
//              query tile).  dK and dV accumulate in TMEM over the item; dQ tiles leave through a TMA fp32 reduce-add into
//              a tile-major accumulator.  8 softmax / dS warps (64 key columns each), 4 drain warps.
//
// Mask predicate (reference: SeqRec/models/generative/Qwen3Multi/model.py:573-741, Qwen3SessionMoe/model.py:416-468) is
// evaluated from per-key codes that fold the key padding mask in (attn_meta_kernel).  Rows with no allowed key are
// "uniform rows" (quirk Q1): forward output = column mean of V over all L keys, lse = +inf; backward uses P = 1/L over
// all keys (the analytic gradient of that uniform softmax).  Uniform rows are not dropped.
#include <stdlib.h>

#include "attention_tc.cuh"
#include "sm100_ptx.cuh"

using namespace sm100;

namespace {

constexpr int D = 64;
constexpr int BT = 128;               // tile edge: queries (and keys in the backward)
constexpr int TILE_BYTES = BT * D * 2;  // 16 KB: one [128 x 64] bf16 SWIZZLE_128B tile
constexpr int KC_MAX = 0x7fffffff;

// Optional in-kernel timeline (debug hook gamer_attn_set_trace): one thread per role of CTA 0 appends (tag, clock64) pairs.
struct Trace {
    long long* buf;
    int cap;
};
__device__ __forceinline__ void trace_pt(const Trace& tr, int role, int& n, int tag) {
    if (tr.buf != nullptr && n < tr.cap) {
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
// per-key codes: ka = valid ? act : MAX, ks = valid ? sess : MAX (valid = j < L and attention_mask[j]); query-side raw
// copies qa / qs; per 32-key block the summary {min ka, max ka, min ks, max ks}.  One block per (sequence, 128 keys).
// ---------------------------------------------------------------------------------------------------------------
__global__ void attn_meta_kernel(const int* __restrict__ am, const int* __restrict__ act, const int* __restrict__ sess,
                                 int L, int Lp, int k_tiles, int* __restrict__ ka, int* __restrict__ ks,
                                 int* __restrict__ qa, int* __restrict__ qs, int4* __restrict__ blk) {
    const int b = blockIdx.x / k_tiles, t = blockIdx.x % k_tiles;
    const int j = t * BT + threadIdx.x;
    int a = KC_MAX, s = KC_MAX, a_raw = 0, s_raw = 0;
    if (j < L) {
        const long long idx = (long long)b * L + j;
        a_raw = act ? act[idx] : 0;
        s_raw = sess ? sess[idx] : 0;
        if (am[idx] != 0) {
            a = a_raw;
            s = s_raw;
        }
    }
    ka[(long long)b * Lp + j] = a;
    ks[(long long)b * Lp + j] = s;
    qa[(long long)b * Lp + j] = a_raw;  // query side: the predicate does not look at the query's own padding bit
    qs[(long long)b * Lp + j] = s_raw;
    const int mn_a = __reduce_min_sync(0xffffffffu, a), mx_a = __reduce_max_sync(0xffffffffu, a);
    const int mn_s = __reduce_min_sync(0xffffffffu, s), mx_s = __reduce_max_sync(0xffffffffu, s);
    if ((threadIdx.x & 31) == 0) blk[((long long)b * Lp + j) >> 5] = make_int4(mn_a, mx_a, mn_s, mx_s);
}

template <int KIND, bool DIAG>
__device__ __forceinline__ bool allow_tc(int ka, int ks, int act_i, int sess_i, int j, int i, int istart) {
    if constexpr (KIND == MASK_CAUSAL) {
        bool ok = (ka == 0);
        if (DIAG) ok = ok && (j <= i);
        return ok;
    } else if constexpr (KIND == MASK_MULTI_CROSS) {
        bool ok = (ka < act_i);
        if (DIAG) ok = ok && (j <= i);
        return ok;
    } else if constexpr (KIND == MASK_SESSION) {
        return (ks < sess_i) || (j >= istart && j <= i && ks != KC_MAX);
    } else {
        return (ks < sess_i) && (ka < act_i);
    }
}

// classification of a 32-key block [j0, j0 + 32) against the 32 query rows [i_lo, i_hi] of a warp
enum { BLK_SKIP = 0, BLK_FULL = 1, BLK_MASK = 2, BLK_MASK_DIAG = 3 };
struct WarpRange {
    int act_min, act_max, sess_min, sess_max;   // over the warp's real rows (i < L); empty warp: min = MAX, max = -1
    int i_lo, i_hi, istart_lo;
};
template <int KIND>
__device__ __forceinline__ int classify_block(const int4 bs, const WarpRange& w, int j0) {
    const bool c_skip = kind_causal<KIND>() && (j0 > w.i_hi);
    const bool c_full = !kind_causal<KIND>() || (j0 + 31 <= w.i_lo);
    bool skip, full;
    if constexpr (KIND == MASK_CAUSAL) {
        skip = c_skip || bs.x == KC_MAX;
        full = c_full && bs.y == 0;
    } else if constexpr (KIND == MASK_MULTI_CROSS) {
        skip = c_skip || bs.x >= w.act_max;
        full = c_full && bs.y < w.act_min;
    } else if constexpr (KIND == MASK_SESSION) {
        skip = bs.x == KC_MAX || (bs.z >= w.sess_max && (j0 > w.i_hi || j0 + 31 < w.istart_lo));
        full = bs.w < w.sess_min;
    } else {
        skip = bs.z >= w.sess_max || bs.x >= w.act_max;
        full = bs.w < w.sess_min && bs.y < w.act_min;
    }
    if (skip) return BLK_SKIP;
    if (full) return BLK_FULL;
    return c_full ? BLK_MASK : BLK_MASK_DIAG;
}
template <int KIND>
__device__ __forceinline__ WarpRange warp_range(int i, int L, int P, int act_i, int sess_i) {
    WarpRange w;
    const bool real = i < L;
    w.act_min = kind_uses_act<KIND>() ? __reduce_min_sync(0xffffffffu, real ? act_i : KC_MAX) : 0;
    w.act_max = kind_uses_act<KIND>() ? __reduce_max_sync(0xffffffffu, real ? act_i : -1) : 0;
    w.sess_min = kind_uses_sess<KIND>() ? __reduce_min_sync(0xffffffffu, real ? sess_i : KC_MAX) : 0;
    w.sess_max = kind_uses_sess<KIND>() ? __reduce_max_sync(0xffffffffu, real ? sess_i : -1) : 0;
    w.i_lo = i & ~31;
    w.i_hi = w.i_lo + 31;
    w.istart_lo = (w.i_lo / P) * P;
    return w;
}

__device__ __forceinline__ uint64_t desc_k(uint32_t saddr) { return umma_desc_sw128(saddr, 16, 1024); }
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo) { return umma_desc_sw128(saddr, lbo, 1024); }

// The issuing warps run converged: every lane waits on the barriers and builds the (warp-uniform) descriptors, one elected
// lane issues the MMAs and the commits.  (A role wrapped in `if (lane == 0)` makes every tcgen05.mma a divergent-code
// sequence — election loop, vector-to-uniform register moves, descriptor rebuild — of ~190 cycles, and a backward step
// has 32 of them.)  Descriptors are built once per operand tile; a k-step adds to the 16-byte address field: +2 for 32
// bytes along a K-major row, +128 for the 2 KB between 16-row groups of an MN-major tile.
// D[128 x 128] = A[128 x 64] B[128 x 64]^T, both K-major tiles (k = head dim)
__device__ __forceinline__ void issue_nt_128x128x64(uint32_t d_tmem, uint32_t sa, uint32_t sb, uint64_t* bar0,
                                                    uint64_t* bar1 = nullptr) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
    const uint64_t da = desc_k(sa), db = desc_k(sb);
    if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, k != 0);
        if (bar0 != nullptr) umma_commit(bar0);
        if (bar1 != nullptr) umma_commit(bar1);
    }
    __syncwarp();
}
// D[128 x 64] (+)= A[128 x 128] B[128 x 64]: A K-major in two 64-wide halves (16 KB apart), B MN-major [128 k-rows x 64]
__device__ __forceinline__ void issue_nn_128x64x128(uint32_t d_tmem, uint32_t sa, uint32_t sb, bool acc, uint64_t* bar0,
                                                    uint64_t* bar1 = nullptr, uint64_t* bar2 = nullptr,
                                                    uint64_t* bar3 = nullptr) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);
    const uint64_t da = desc_k(sa), db = desc_mn(sb, TILE_BYTES);
    if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            umma_bf16(d_tmem, da + (k >> 2) * (TILE_BYTES / 16) + (k & 3) * 2, db + k * 128, idesc, (acc || k != 0) ? 1u : 0u);
        umma_commit(bar0);
        if (bar1 != nullptr) umma_commit(bar1);
        if (bar2 != nullptr) umma_commit(bar2);
        if (bar3 != nullptr) umma_commit(bar3);
    }
    __syncwarp();
}
// D[128 x 64] (+)= A^T B with A stored [128 k-rows x 128 m] (two 64-wide halves, MN-major) and B [128 k-rows x 64] MN-major
__device__ __forceinline__ void issue_tn_128x64x128(uint32_t d_tmem, uint32_t sa, uint32_t sb, bool acc, uint64_t* bar0,
                                                    uint64_t* bar1 = nullptr) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
    const uint64_t da = desc_mn(sa, TILE_BYTES), db = desc_mn(sb, TILE_BYTES);
    if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16(d_tmem, da + k * 128, db + k * 128, idesc, (acc || k != 0) ? 1u : 0u);
        if (bar0 != nullptr) umma_commit(bar0);
        if (bar1 != nullptr) umma_commit(bar1);
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// packed fp32 helpers (sm_100: FFMA2 / FADD2 / FMUL2, FMNMX3)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float2 u2f2(uint32_t a, uint32_t b) { return make_float2(__uint_as_float(a), __uint_as_float(b)); }
// row max of 32 fp32 values held as raw words
__device__ __forceinline__ float max32(const uint32_t* s) {
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int c = 0; c < 32; c += 4) {
        m0 = max3(m0, __uint_as_float(s[c]), __uint_as_float(s[c + 1]));
        m1 = max3(m1, __uint_as_float(s[c + 2]), __uint_as_float(s[c + 3]));
    }
    return fmaxf(m0, m1);
}

// ---------------------------------------------------------------------------------------------------------------
// dropout keep words.  Word (bh, i, jw) covers keys [32 jw, 32 jw + 32) of query i; key 32 jw + 4 g + e <-> bit g + 8 e
// (so that one shift brings the flags of four consecutive keys to the four byte sign positions).  Lane bit c of eight
// random words r0..r7 (one Philox call, see keep_word) forms the 8-bit number R_c; the key is dropped iff R_c < thresh.
// ---------------------------------------------------------------------------------------------------------------
struct KeepGen {
    uint32_t k0, k1, site;
    uint32_t tm[8];   // bit k of the threshold, expanded to a full word
};
__device__ __forceinline__ KeepGen keep_gen(const DropParams& d) {
    KeepGen g;
    g.k0 = d.k0;
    g.k1 = d.k1;
    g.site = d.site;
#pragma unroll
    for (int k = 0; k < 8; ++k) g.tm[k] = 0u - ((d.thresh >> k) & 1u);
    return g;
}
__device__ __forceinline__ uint32_t keep_word(const KeepGen& g, uint32_t bh, uint32_t i, uint32_t jw) {
    // One Philox call per (query, 32-key block).  Its four words are the four HIGH bit planes of the 32 lane numbers R_c;
    // the four LOW planes are the same words rotated by 5, 13, 21 and 29 bits, i.e. lane c's low bit k is the high bit k
    // of lane c - rot_k: all eight bits of R_c still come from eight distinct (word, bit) cells of the generator's
    // output, so every R_c is exactly uniform on 0..255 and the drop probability is exactly thresh / 256; what is given
    // up is the independence between one lane's low bits and another lane's high bits, which a dropout mask never sees
    // (the low planes decide only when the high nibble ties with the threshold's).  Halves the generator's cost.
    const uint4 a = philox4x32_7(jw, i, bh, g.site, g.k0, g.k1);
    const uint32_t r[8] = {__funnelshift_l(a.x, a.x, 5), __funnelshift_l(a.y, a.y, 13), __funnelshift_l(a.z, a.z, 21),
                           __funnelshift_l(a.w, a.w, 29), a.x, a.y, a.z, a.w};
    uint32_t bo = 0u;   // borrow of R - thresh, least significant bit first: bo = (R < thresh) per lane
#pragma unroll
    for (int k = 0; k < 8; ++k) bo = (~r[k] & g.tm[k]) | (~(r[k] ^ g.tm[k]) & bo);
    return ~bo;
}
// zero the dropped elements of 32 probabilities held as 16 packed bf16 pairs (pair w = keys 2w, 2w + 1)
__device__ __forceinline__ void keep_apply_packed(uint32_t* pk, uint32_t kw) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const uint32_t sh = kw << (7 - g);
        pk[2 * g] &= prmt(sh, 0u, 0x9988u);
        pk[2 * g + 1] &= prmt(sh, 0u, 0xbbaau);
    }
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

// =================================================================================================================
// forward
// =================================================================================================================
constexpr int S_KT = 64;                                   // keys per tile
constexpr int S_HALF = S_KT * D * 2;                       // 8 KB: one [64 x 64] bf16 tile
constexpr int S_CODES = 640;                               // ka[64] | ks[64] | blk[2] int4 (+ pad)
constexpr int S_KST = 4;                                   // K (+ key codes) stages
constexpr int S_VST = 3;                                   // V stages
constexpr int S_OFF_Q = 0;                                 // [128 x 64] Q, later the O staging tile
constexpr int S_OFF_K = TILE_BYTES;
constexpr int S_OFF_V = S_OFF_K + S_KST * S_HALF;
constexpr int S_OFF_C = S_OFF_V + S_VST * S_HALF;          // key codes (they travel with K)
constexpr int S_OFF_BAR = S_OFF_C + S_KST * S_CODES;
constexpr int S_OFF_X = S_OFF_BAR + 128;                   // epilogue exchange of (m, l) between the two halves: 2 KB
constexpr int S_SMEM = S_OFF_X + 2048 + 1024;
constexpr int S_THREADS = 288;                             // 8 softmax warps + 1 issuer warp

struct FwdParams {
    int B, L, Lp, n_q, n_kv, P, q_tiles;
    const int* ka;
    const int* ks;
    const int4* blk;
    const int* act;
    const int* sess;
    float scale_log2;
    const float* vmean;
    float* lse;
    uint32_t* keep;   // [B * n_q][q_tiles][Lp / 32][128] keep words (written when dropout is on)
    DropParams drop;  // attention-probability dropout (8-bit threshold); thresh == 0: off
    Trace tr;
};

__device__ __forceinline__ int4 lds_int4(uint32_t saddr) {
    int4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
    return v;
}

// masked copy + row max of one 32-key block (slow path: the predicate per element; ka / ks = shared-memory addresses of
// the block's key codes)
template <int KIND, bool DIAG>
__device__ __forceinline__ float mask_max32(uint32_t* s, uint32_t ka, uint32_t ks, int act_i, int sess_i, int j0, int i,
                                            int istart) {
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
        int4 a4 = make_int4(0, 0, 0, 0), s4 = make_int4(0, 0, 0, 0);
        if (KIND != MASK_SESSION) a4 = lds_int4(ka + c4 * 16);
        if (kind_uses_sess<KIND>()) s4 = lds_int4(ks + c4 * 16);
        const int av[4] = {a4.x, a4.y, a4.z, a4.w};
        const int sv[4] = {s4.x, s4.y, s4.z, s4.w};
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = c4 * 4 + e;
            const bool ok = allow_tc<KIND, DIAG>(av[e], sv[e], act_i, sess_i, j0 + c, i, istart);
            v[e] = ok ? __uint_as_float(s[c]) : -INFINITY;
            s[c] = __float_as_uint(v[e]);
        }
        m0 = max3(m0, v[0], v[1]);
        m1 = max3(m1, v[2], v[3]);
    }
    return fmaxf(m0, m1);
}

// Dropout of the probabilities (SDPA dropout_p, Qwen3Multi/model.py:139): P is zeroed where the keep bit of (sequence,
// head, query, key) is clear; the row sum l keeps the undropped P (softmax normalisation happens before dropout) and the
// 1/keep scale is folded into the final O / l.
//
// Two CTAs per SM, 256 TMEM columns each.  S is double-buffered ([0,64) and [64,128); P, packed bf16 pairs, overwrites
// columns of the buffer its scores came from): S(n+2) is issued right behind PV(n) — the tensor pipe runs the MMAs
// of one thread in order, so the write follows PV(n)'s read of P(n) — which means S(n+1) is already in TMEM when the
// softmax warps finish tile n.  The 64 keys of a tile are split between two softmax warpgroups (8 warps, one query row
// and 32 keys per thread) that run INDEPENDENT online softmaxes — own running max, own row sum, own accumulator
// (O_A in [128,192) over the first halves of all tiles, O_B in [192,256) over the second halves) — merged once in the
// epilogue, so the warpgroups never exchange anything per tile and the SM holds 16 softmax warps.
template <int KIND, bool DROP>
__global__ void __launch_bounds__(S_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, FwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S_OFF_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;   // [4] K tile + key codes
    uint64_t* v_full = bars + 5;   // [3]
    uint64_t* s_full = bars + 8;   // [2] one per S buffer
    uint64_t* p_full = bars + 10;  // [2] P tile in TMEM (and O rescaled), one per S buffer: the softmax warps can be a
                                   // whole tile ahead of the issuer, and a parity wait must never fall two phases behind
    uint64_t* pv_done = bars + 12; // PV(n) complete
    uint64_t* o_full = bars + 13;  // the last PV complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
    int* range_slot = reinterpret_cast<int*>(bars + 15);   // [lo, hi) in 64-key tiles
    float* xch = reinterpret_cast<float*>(smem + S_OFF_X);  // epilogue exchange: [2 halves][m | l][128 rows]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // work item: heaviest query tiles first
    const int per = p.B * p.n_q;
    const int qt = p.q_tiles - 1 - (int)blockIdx.x / per;
    const int rem = (int)blockIdx.x % per;
    const int b = rem / p.n_q, h = rem % p.n_q;
    const int g = h / (p.n_q / p.n_kv);
    const int kt_all = (p.L + S_KT - 1) / S_KT;

    if (warp == 8) {
        // the issuer warp sets the barriers up itself, so the Q tile (which does not depend on the key range) is already
        // in flight while the TMEM allocation and the key-range scan below run
        if (lane == 0) {
            prefetch_tmap(&tmQ);
            prefetch_tmap(&tmK);
            prefetch_tmap(&tmV);
            prefetch_tmap(&tmO);
            for (int i = 0; i < 14; ++i) mbar_init(&bars[i], (i == 10 || i == 11) ? 256 : 1);
            fence_barrier_init();
            mbar_expect_tx(q_full, TILE_BYTES);
            tma_load_3d(smem + S_OFF_Q, &tmQ, q_full, h * D, qt * BT, b);
        }
        __syncwarp();
        tmem_alloc<256>(tmem_slot);
        // Key-tile range of this CTA: leading tiles without a valid key (left padding) and trailing tiles that no query of
        // the tile can see are never loaded.  Lane t looks at 64-key tiles t and t + 32.
        int q_smax = 0, q_amax = 0;   // over the tile's real queries
        if (!kind_causal<KIND>() || kind_uses_act<KIND>()) {
            int sm = -1, amx = -1;
            for (int r = lane; r < BT; r += 32) {
                const int i = qt * BT + r;
                if (i < p.L) {
                    if (kind_uses_sess<KIND>()) sm = max(sm, p.sess[(long long)b * p.L + i]);
                    if (kind_uses_act<KIND>()) amx = max(amx, p.act[(long long)b * p.L + i]);
                }
            }
            q_smax = __reduce_max_sync(0xffffffffu, sm);
            q_amax = __reduce_max_sync(0xffffffffu, amx);
        }
        const int i_max = min(qt * BT + BT - 1, p.L - 1);
        unsigned need_lo = 0, need_hi = 0;   // bit = this lane's tile holds a pair that may be allowed
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
            const int t = lane + rep * 32;
            bool need = false;
            if (t < kt_all) {
                const int4 s0 = p.blk[((long long)b * p.Lp >> 5) + 2 * t], s1 = p.blk[((long long)b * p.Lp >> 5) + 2 * t + 1];
                const int mn_a = min(s0.x, s1.x), mn_s = min(s0.z, s1.z);
                const bool any_valid = mn_a != KC_MAX;
                if constexpr (KIND == MASK_CAUSAL) need = any_valid && t * S_KT <= i_max;
                else if constexpr (KIND == MASK_MULTI_CROSS) need = mn_a < q_amax && t * S_KT <= i_max;
                else if constexpr (KIND == MASK_SESSION) need = any_valid && (mn_s < q_smax || t * S_KT <= i_max);
                else need = mn_s < q_smax && mn_a < q_amax;
            }
            const unsigned m = __ballot_sync(0xffffffffu, need);
            if (rep == 0) need_lo = m; else need_hi = m;
        }
        if (lane == 0) {
            const unsigned long long need = ((unsigned long long)need_hi << 32) | need_lo;
            range_slot[0] = need ? __ffsll((long long)need) - 1 : 0;
            range_slot[1] = need ? 64 - __clzll((long long)need) : 0;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int kt_lo = range_slot[0], kt_hi = range_slot[1];
    const int nkt = kt_hi - kt_lo;
    const uint32_t sq = smem_u32(smem + S_OFF_Q);
    constexpr uint32_t T_O = 128;

    if (warp == 8) {
        // ===================== issuer: TMA loads + MMAs (converged warp, elected lane issues) =====================
        auto load_k = [&](int n) {  // n-th tile of the range: K and its key codes
            if (elect_one()) {
                const int st = n % S_KST, j = kt_lo + n;
                uint8_t* sc = smem + S_OFF_C + st * S_CODES;
                mbar_expect_tx(&k_full[st], S_HALF + 512 + 32);
                tma_load_3d(smem + S_OFF_K + st * S_HALF, &tmK, &k_full[st], g * D, j * S_KT, b);
                bulk_load_1d(sc, p.ka + (long long)b * p.Lp + j * S_KT, 256, &k_full[st]);
                bulk_load_1d(sc + 256, p.ks + (long long)b * p.Lp + j * S_KT, 256, &k_full[st]);
                bulk_load_1d(sc + 512, p.blk + ((long long)b * p.Lp >> 5) + 2 * j, 32, &k_full[st]);
            }
            __syncwarp();
        };
        auto load_v = [&](int n) {
            if (elect_one()) {
                const int st = n % S_VST;
                mbar_expect_tx(&v_full[st], S_HALF);
                tma_load_3d(smem + S_OFF_V + st * S_HALF, &tmV, &v_full[st], g * D, (kt_lo + n) * S_KT, b);
            }
            __syncwarp();
        };
        auto issue_s = [&](int n) {  // S(n)[128 x 64] = Q K_n^T into buffer n & 1
            constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
            mbar_wait(&k_full[n % S_KST], (n / S_KST) & 1);
            tc_fence_after();
            const uint64_t da = desc_k(sq), db = desc_k(smem_u32(smem + S_OFF_K + (n % S_KST) * S_HALF));
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + (n & 1) * 64, da + 2 * k, db + 2 * k, idesc, k != 0);
                umma_commit(&s_full[n & 1]);
            }
            __syncwarp();
        };
        // O_A[128 x 64] (+)= P(n)[:, 0:32] V_n[0:32], O_B (+)= P(n)[:, 32:64] V_n[32:64]; each half's P (16 packed columns)
        // sits at the start of that half's 32 score columns
        auto issue_pv = [&](int n, bool last) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);
            const uint64_t db = desc_mn(smem_u32(smem + S_OFF_V + (n % S_VST) * S_HALF), S_HALF);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ts(tmem_base + T_O + (k >> 1) * 64, tmem_base + (n & 1) * 64 + (k >> 1) * 32 + (k & 1) * 8,
                                 db + k * 128, idesc, (n > 0 || (k & 1) != 0) ? 1u : 0u);
                umma_commit(pv_done);
                if (last) umma_commit(o_full);
            }
            __syncwarp();
        };
        if (nkt > 0) {
            for (int n = 0; n < S_KST && n < nkt; ++n) load_k(n);
            for (int n = 0; n < S_VST && n < nkt; ++n) load_v(n);
            mbar_wait(q_full, 0);
            issue_s(0);
            if (nkt > 1) issue_s(1);
            Trace tr = p.tr;
            if (blockIdx.x != 0 || lane != 0) tr.buf = nullptr;
            int tn = 0;
            for (int n = 0; n < nkt; ++n) {
                trace_pt(tr, 0, tn, 1);
                mbar_wait(&p_full[n & 1], (n >> 1) & 1);   // S(n) consumed, P(n) in TMEM
                trace_pt(tr, 0, tn, 2);
                // loads run two tiles ahead: K(n+4) into the stage of K(n) (S(n) has been consumed), V(n+2) into the stage
                // PV(n-1) read (issued a whole tile ago)
                if (n + S_KST < nkt) load_k(n + S_KST);
                if (n >= 1 && n + S_VST - 1 < nkt) {
                    mbar_wait(pv_done, (n - 1) & 1);
                    load_v(n + S_VST - 1);
                }
                mbar_wait(&v_full[n % S_VST], (n / S_VST) & 1);
                tc_fence_after();
                issue_pv(n, n + 1 == nkt);
                trace_pt(tr, 0, tn, 3);
                if (n + 2 < nkt) issue_s(n + 2);   // into the buffer PV(n) reads P from: ordered behind it
                trace_pt(tr, 0, tn, 4);
            }
        }
    } else {
        // ===================== softmax: one query row and one 32-key half of every tile per thread =====================
        const int hf = warp >> 2;                  // which half of the 64-key tiles (and which accumulator)
        const int wq = warp & 3;
        const int row = wq * 32 + lane;
        const uint32_t t_row = tmem_base + ((uint32_t)(wq * 32) << 16);
        const uint32_t t_o = t_row + T_O + hf * 64;
        const int i = qt * BT + row;
        int act_i = 0, sess_i = 0;
        if (i < p.L) {
            if (kind_uses_act<KIND>()) act_i = p.act[(long long)b * p.L + i];
            if (kind_uses_sess<KIND>()) sess_i = p.sess[(long long)b * p.L + i];
        }
        const int istart = (i / p.P) * p.P;
        const WarpRange wr = warp_range<KIND>(i, p.L, p.P, act_i, sess_i);
        const DropParams drop = drop_resolve(p.drop);
        KeepGen kg;
        if constexpr (DROP) kg = keep_gen(drop);
        const uint32_t bh = (uint32_t)(b * p.n_q + h);
        uint32_t* keep_row = nullptr;
        if constexpr (DROP) keep_row = p.keep + ((size_t)bh * p.q_tiles + qt) * (size_t)(p.Lp >> 5) * BT + row;
        float m = -INFINITY, l = 0.f;
        int kst = 0;
        Trace tr = p.tr;
        if (blockIdx.x != 0 || threadIdx.x != 0) tr.buf = nullptr;
        int tn = 0;
        trace_pt(tr, 1, tn, 19);
        for (int n = 0; n < nkt; ++n) {
            const int j0 = (kt_lo + n) * S_KT + hf * 32;   // first key of this thread's block
            const uint8_t* sc = smem + S_OFF_C + kst * S_CODES;
            const uint32_t t_s = t_row + (n & 1) * 64;
            trace_pt(tr, 1, tn, 20);
            // S(n) was issued after K(n) and its key codes had landed: one wait covers both
            mbar_wait(&s_full[n & 1], (n >> 1) & 1);
            trace_pt(tr, 1, tn, 22);
            tc_fence_after();
            const int mode = classify_block<KIND>(reinterpret_cast<const int4*>(sc + 512)[hf], wr, j0);
            uint32_t s[32];
            bool rescale = false;
            float m_old = m, m_new = m;
            if (mode != BLK_SKIP) {   // (a skipped block leaves m, l and the accumulator alone and contributes P = 0)
                tmem_ld_32x32(t_s + hf * 32, s);
                tmem_ld_wait();
                const uint32_t ka = smem_u32(sc) + hf * 128, ks = ka + 256;
                float mx;
                if (mode == BLK_FULL) mx = max32(s);
                else if (mode == BLK_MASK) mx = mask_max32<KIND, false>(s, ka, ks, act_i, sess_i, j0, i, istart);
                else mx = mask_max32<KIND, true>(s, ka, ks, act_i, sess_i, j0, i, istart);
                // lazy running max (log2 domain): raise it only from -inf or by more than 2^64
                const float m_tile = mx * p.scale_log2;
                if (m == -INFINITY) m_new = m_tile;
                else if (m_tile > m + 64.f) m_new = m_tile;
                rescale = (m != -INFINITY) && (m_new != m);
                if (rescale) l *= ex2_approx(m - m_new);
                m = m_new;
                const float neg_m = (m == -INFINITY) ? 0.f : -m;
                const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(neg_m, neg_m);
                float2 sum = make_float2(0.f, 0.f);
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                    const float2 x = __ffma2_rn(u2f2(s[c], s[c + 1]), sc2, nm2);
                    const float2 e = make_float2(ex2_approx(x.x), ex2_approx(x.y));
                    sum = __fadd2_rn(sum, e);
                    s[c >> 1] = pack_bf16(e.x, e.y);
                }
                l += sum.x + sum.y;
                if constexpr (DROP) {
                    const uint32_t jw = (uint32_t)(2 * (kt_lo + n) + hf);
                    const uint32_t kw = keep_word(kg, bh, (uint32_t)i, jw);
                    keep_row[(size_t)jw * BT] = kw;
                    keep_apply_packed(s, kw);
                }
            } else {
#pragma unroll
                for (int w = 0; w < 16; ++w) s[w] = 0u;
            }
            if (n > 0 && __any_sync(0xffffffffu, rescale)) {   // cold path: O must be stable, i.e. PV(n-1) complete
                mbar_wait(pv_done, (n - 1) & 1);
                tc_fence_after();
                rescale_o_row64(t_o, rescale ? ex2_approx(m_old - m_new) : 1.f);
            }
            trace_pt(tr, 1, tn, 23);
            tmem_st_32x16(t_s + hf * 32, s);   // over the first 16 of this half's own (already read) score columns
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&p_full[n & 1]);
            trace_pt(tr, 1, tn, 24);
            if (++kst == S_KST) kst = 0;
        }
        // ---- epilogue: merge the two halves, O / l (or the V column mean on uniform rows) -> bf16 -> smem (the dead Q
        // tile) -> TMA store.  Thread (row, hf) writes head-dim columns [32 hf, 32 hf + 32).
        xch[hf * 256 + row] = m;
        xch[hf * 256 + 128 + row] = l;
        if (nkt > 0) {
            mbar_wait(o_full, 0);
            tc_fence_after();
        }
        named_bar_sync(1, 256);
        trace_pt(tr, 1, tn, 25);
        const float m_o = xch[(hf ^ 1) * 256 + row], l_o = xch[(hf ^ 1) * 256 + 128 + row];
        const float m_all = fmaxf(m, m_o);
        const float w_me = (m == -INFINITY) ? 0.f : ex2_approx(m - m_all);
        const float w_ot = (m_o == -INFINITY) ? 0.f : ex2_approx(m_o - m_all);
        const float l_all = l * w_me + l_o * w_ot;
        const bool uniform = !(l_all > 0.f);
        const float inv = uniform ? 0.f : (DROP ? drop.scale : 1.f) / l_all;
        const float wa = (hf == 0 ? w_me : w_ot) * inv, wb = (hf == 0 ? w_ot : w_me) * inv;   // weights of O_A, O_B
        const float* vm = p.vmean + ((long long)b * p.n_kv + g) * D + hf * 32;
        {
            uint32_t oa[32], ob[32];
            if (nkt > 0) {
                tmem_ld_32x32(t_row + T_O + hf * 32, oa);
                tmem_ld_32x32(t_row + T_O + 64 + hf * 32, ob);
                tmem_ld_wait();
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    // an accumulator no PV ever wrote (its half was skipped in every tile) holds garbage: weight exactly 0
                    const float a = (wa != 0.f) ? __uint_as_float(oa[8 * q + e]) * wa : 0.f;
                    const float c = (wb != 0.f) ? __uint_as_float(ob[8 * q + e]) * wb : 0.f;
                    v[e] = uniform ? vm[8 * q + e] : a + c;
                }
                const int ch = hf * 4 + q;
                const bf16x8 ov = float_to_bf16x8(v);
                sts128(sq + row * 128 + ((ch ^ (row & 7)) << 4), ov.u[0], ov.u[1], ov.u[2], ov.u[3]);
            }
        }
        if (hf == 0 && i < p.L) p.lse[((long long)b * p.n_q + h) * p.L + i] = uniform ? INFINITY : (m_all + log2f(l_all));
        tc_fence_before();
        fence_proxy_async();
        named_bar_sync(1, 256);
        if (threadIdx.x == 0) {
            tma_store_3d(&tmO, smem + S_OFF_Q, h * D, qt * BT, b);
            bulk_commit();
            bulk_wait_read0();   // the CTA may retire once the store has read the staging tile
            trace_pt(tr, 1, tn, 26);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        // a CTA that never waited on q_full (empty key range) must not retire with the Q load still writing its smem
        if (nkt == 0) mbar_wait(q_full, 0);
        tmem_dealloc<256>(tmem_base);
    }
}

// =================================================================================================================
// backward
// =================================================================================================================
constexpr int B_KV_META = 2048;                            // ka[128] | ks[128] | blk[4] int4 (+ pad: stages stay 1 KB aligned)
constexpr int B_KV_STAGE = 2 * TILE_BYTES + B_KV_META;     // K | V | codes
constexpr int B_OFF_KV = 0;                                // 2 stages (items)
constexpr int B_QDO_ROWS = 2048;                           // lse'[128] | dsum'[128] | level[128] | session[128]
constexpr int B_QDO_KEEP = 2048;                           // keep words [4 blocks][128 rows]
constexpr int B_QDO_STAGE = 2 * TILE_BYTES + B_QDO_ROWS + B_QDO_KEEP;   // Q | dO | row scalars | keep words
constexpr int B_OFF_QDO = 2 * B_KV_STAGE;                  // 2 stages
constexpr int B_OFF_P = B_OFF_QDO + 2 * B_QDO_STAGE;       // [128 q x 128 k] bf16 in two 64-key halves
constexpr int B_OFF_DS = B_OFF_P + 2 * TILE_BYTES;
constexpr int B_OFF_STG = B_OFF_DS + 2 * TILE_BYTES;       // 16 KB staging: dQ half tiles (fp32), dK / dV tiles (bf16)
constexpr int B_OFF_BAR = B_OFF_STG + TILE_BYTES;
constexpr int B_SMEM = B_OFF_BAR + 256 + 1024;
constexpr int B_THREADS = 512;                             // producer, issuer, 2 spare | 8 softmax / dS warps | 4 drain warps
constexpr int DQ_TILE_FLOATS = BT * D;
static_assert(B_SMEM <= 227 * 1024, "backward shared memory");
static_assert(B_KV_STAGE % 1024 == 0 && B_QDO_STAGE % 1024 == 0 && B_OFF_P % 1024 == 0 && B_OFF_STG % 1024 == 0,
              "SWIZZLE_128B tiles need 1024-byte aligned bases");

struct BwdParams {
    int B, L, Lp, n_q, n_kv, P, q_tiles, k_tiles, total;
    const int* ka;
    const int* ks;
    const int4* blk;
    const int* qa;
    const int* qs;       // [B, Lp] query-side behaviour level / session (zero past L)
    const float* lse_p;  // [B, n_q, Lp] log2 domain minus log2(1/keep); +inf = uniform row or i >= L
    const float* dsum_p; // [B, n_q, Lp] rowsum(dO o O) * keep (uniform rows: unscaled)
    const unsigned* uni_bits;  // [B]: bit qt = query tile qt holds a uniform row
    const uint32_t* keep;      // the forward's keep words
    float scale, scale_log2, inv_L;
    float* dq_acc;       // [B, n_q, q_tiles][2 halves][128 rows][32 floats], 16-byte chunks XOR-swizzled by (row & 7)
    bf16* dq;            // [B * L, ld_d]: head h at columns [64 h, 64 h + 64)
    long long ld_d;
    int drop_on;
    Trace tr;
};

template <int KIND>
__device__ __forceinline__ bool bwd_decode(int w, const BwdParams& p, int& b, int& g, int& kt, unsigned& qmask,
                                           unsigned& lastmask) {
    // Work order.  A "group" = (sequence, kv head) is OWNED by one CTA, which walks the group's key tiles 0..k_tiles-1 back
    // to back (w = blockIdx.x + m * gridDim.x  <->  group blockIdx.x + (m / k_tiles) * gridDim.x, key tile m % k_tiles).
    // The group's Q / dO tiles stay in L2 across its key tiles, and because every contribution to a dQ tile comes from
    // the same CTA in key-tile order the fp32 dQ accumulator needs neither a zero fill nor a conversion pass: the first
    // contribution (key tile 0) is a plain store, the middle ones are reduce-adds, and the last one (lastmask) reads the
    // running sum back, adds its own tile and writes bf16 dQ.
    const int grid = (int)gridDim.x;
    const int m = w / grid;
    kt = m % p.k_tiles;
    const int grp = (m / p.k_tiles) * grid + w % grid;
    if (grp >= p.B * p.n_kv) return false;
    b = grp / p.n_kv;
    g = grp % p.n_kv;
    const unsigned all = (p.q_tiles >= 32) ? 0xffffffffu : ((1u << p.q_tiles) - 1u);
    if (kind_causal<KIND>()) {
        // query tile qt meets key tile kt when qt >= kt, or always when it holds a uniform row (P = 1/L on every key)
        const unsigned uni = p.uni_bits[b] & all;
        qmask = ((all >> kt) << kt) | uni;
        lastmask = (kt == p.k_tiles - 1) ? (uni | (1u << kt)) : ((1u << kt) & ~uni);
    } else {
        qmask = all;
        lastmask = (kt == p.k_tiles - 1) ? all : 0u;
    }
    return true;
}

// One 32-key block of a backward step, for one query row.  In: s = scores (raw words), dp = dO V^T.  Out: pk = the dV
// operand (P with dropout and 1/keep applied) and ds = the dK / dQ operand, both as 16 packed bf16 pairs.
//   Pz = exp2(s * scale_log2 - lse') = P / keep_prob (uniform rows: 1/L on every key j < L, no dropout)
//   dS = (Pz o keep) o dP - Pz * dsum'   with dsum' = rowsum(dO o O) * keep_prob (uniform rows: unscaled)
// MODE: BLK_FULL (no predicate, no uniform row in the warp), BLK_MASK / BLK_MASK_DIAG (predicate per element, uniform rows
// handled), BLK_SKIP is handled by the caller.
template <int KIND, bool DROP, int MODE>
__device__ __forceinline__ void bwd_block(const uint32_t* s, const uint32_t* dp, uint32_t* pk, uint32_t* ds, uint32_t ka,
                                          uint32_t ks, int act_i, int sess_i, int j0, int i, int istart, float scale_log2,
                                          float neg_lse, float neg_dsum, float pu, int lim_u, uint32_t kw) {
    const float2 sc2 = make_float2(scale_log2, scale_log2), nl2 = make_float2(neg_lse, neg_lse);
    const float2 nd2 = make_float2(neg_dsum, neg_dsum);
    float2 pz[16];
    if constexpr (MODE == BLK_FULL) {
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
            const float2 x = __ffma2_rn(u2f2(s[c], s[c + 1]), sc2, nl2);
            pz[c >> 1] = make_float2(ex2_approx(x.x), ex2_approx(x.y));
        }
    } else {
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
            int4 a4 = make_int4(0, 0, 0, 0), s4 = make_int4(0, 0, 0, 0);
            if (KIND != MASK_SESSION) a4 = lds_int4(ka + c4 * 16);
            if (kind_uses_sess<KIND>()) s4 = lds_int4(ks + c4 * 16);
            const int av[4] = {a4.x, a4.y, a4.z, a4.w};
            const int sv[4] = {s4.x, s4.y, s4.z, s4.w};
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = c4 * 4 + e;
                const bool ok = allow_tc<KIND, MODE == BLK_MASK_DIAG>(av[e], sv[e], act_i, sess_i, j0 + c, i, istart);
                const float x = fmaf(__uint_as_float(s[c]), scale_log2, neg_lse);
                v[e] = ex2_approx(ok ? x : -INFINITY);          // uniform rows: lse' = +inf -> 0
                v[e] = (c < lim_u) ? pu : v[e];                   // lim_u > 0 only on uniform rows
            }
            pz[c4 * 2] = make_float2(v[0], v[1]);
            pz[c4 * 2 + 1] = make_float2(v[2], v[3]);
        }
    }
#pragma unroll
    for (int w = 0; w < 16; ++w) pk[w] = pack_bf16(pz[w].x, pz[w].y);
    if constexpr (DROP) {
        keep_apply_packed(pk, kw);
#pragma unroll
        for (int w = 0; w < 16; ++w) {
            const float2 pzk = make_float2(__uint_as_float(pk[w] << 16), __uint_as_float(pk[w] & 0xffff0000u));
            const float2 t = __fmul2_rn(pz[w], nd2);
            const float2 d = __ffma2_rn(pzk, u2f2(dp[2 * w], dp[2 * w + 1]), t);
            ds[w] = pack_bf16(d.x, d.y);
        }
    } else {
#pragma unroll
        for (int w = 0; w < 16; ++w) {
            const float2 t = __fadd2_rn(u2f2(dp[2 * w], dp[2 * w + 1]), nd2);
            const float2 d = __fmul2_rn(pz[w], t);
            ds[w] = pack_bf16(d.x, d.y);
        }
    }
}

template <int KIND, bool DROP>
__global__ void __launch_bounds__(B_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                const __grid_constant__ CUtensorMap tmdK, const __grid_constant__ CUtensorMap tmdV, BwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B_OFF_BAR);
    uint64_t* kv_full = bars + 0;    // [2]
    uint64_t* kv_free = bars + 2;    // [2]
    uint64_t* qdo_full = bars + 4;   // [2]
    uint64_t* qdo_free = bars + 6;   // [2]
    uint64_t* sdp_full = bars + 8;   // S and dP of a step in TMEM (one commit behind both MMAs)
    uint64_t* sdp_free = bars + 9;   // all softmax threads have read S and dP
    uint64_t* pds_full = bars + 12;  // P and dS tiles written
    uint64_t* pds_free = bars + 13;  // dV, dK and dQ of the step complete: the P / dS tiles may be overwritten
    uint64_t* dq_full = bars + 15;   // [2]: dQ lives in two TMEM buffers, so dQ(n) does not wait for the drain of dQ(n-1)
    uint64_t* dq_free = bars + 17;   // [2]
    uint64_t* dkv_full = bars + 19;
    uint64_t* dkv_free = bars + 20;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);
    uint32_t* conv_flag = reinterpret_cast<uint32_t*>(bars + 22);   // items whose dQ accumulator writes have landed

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        *conv_flag = 0u;
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
            mbar_init(&dq_full[s], 1);
            mbar_init(&dq_free[s], 128);
        }
        mbar_init(sdp_full, 1);
        mbar_init(sdp_free, 256);
        mbar_init(pds_full, 256);
        mbar_init(pds_free, 1);
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
    constexpr uint32_t QDO_TX = 2 * TILE_BYTES + B_QDO_ROWS + (DROP ? B_QDO_KEEP : 0);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t item_n = 0, step_n = 0;
            for (int w = blockIdx.x; w < p.total; w += gridDim.x, ++item_n) {
                int b, g, kt;
                unsigned qmask, lastmask;
                if (!bwd_decode<KIND>(w, p, b, g, kt, qmask, lastmask)) break;
                const int ks = item_n & 1;
                uint8_t* skv = smem + B_OFF_KV + ks * B_KV_STAGE;
                mbar_wait(&kv_free[ks], ((item_n >> 1) & 1) ^ 1);
                mbar_expect_tx(&kv_full[ks], 2 * TILE_BYTES + 1024 + 64);
                tma_load_3d(skv, &tmK, &kv_full[ks], g * D, kt * BT, b);
                tma_load_3d(skv + TILE_BYTES, &tmV, &kv_full[ks], g * D, kt * BT, b);
                bulk_load_1d(skv + 2 * TILE_BYTES, p.ka + (long long)b * p.Lp + kt * BT, 512, &kv_full[ks]);
                bulk_load_1d(skv + 2 * TILE_BYTES + 512, p.ks + (long long)b * p.Lp + kt * BT, 512, &kv_full[ks]);
                bulk_load_1d(skv + 2 * TILE_BYTES + 1024, p.blk + ((long long)b * p.Lp >> 5) + 4 * kt, 64, &kv_full[ks]);
                for (int hh = 0; hh < 2; ++hh) {
                    for (unsigned qm = qmask; qm; qm &= qm - 1, ++step_n) {
                        const int qt = __ffs(qm) - 1;
                        const int st = step_n & 1;
                        mbar_wait(&qdo_free[st], ((step_n >> 1) & 1) ^ 1);
                        uint8_t* sq = smem + B_OFF_QDO + st * B_QDO_STAGE;
                        mbar_expect_tx(&qdo_full[st], QDO_TX);
                        tma_load_3d(sq, &tmQ, &qdo_full[st], (2 * g + hh) * D, qt * BT, b);
                        tma_load_3d(sq + TILE_BYTES, &tmdO, &qdo_full[st], (2 * g + hh) * D, qt * BT, b);
                        const long long bh = (long long)b * p.n_q + 2 * g + hh;
                        const long long ro = bh * p.Lp + qt * BT;
                        uint8_t* sr = sq + 2 * TILE_BYTES;
                        bulk_load_1d(sr, p.lse_p + ro, 512, &qdo_full[st]);
                        bulk_load_1d(sr + 512, p.dsum_p + ro, 512, &qdo_full[st]);
                        bulk_load_1d(sr + 1024, p.qa + (long long)b * p.Lp + qt * BT, 512, &qdo_full[st]);
                        bulk_load_1d(sr + 1536, p.qs + (long long)b * p.Lp + qt * BT, 512, &qdo_full[st]);
                        if constexpr (DROP)
                            bulk_load_1d(sr + B_QDO_ROWS,
                                         p.keep + ((size_t)(bh * p.q_tiles + qt) * (size_t)(p.Lp >> 5) + 4 * kt) * BT,
                                         B_QDO_KEEP, &qdo_full[st]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (converged warp, elected lane issues) =====================
        uint32_t item_n = 0, step_n = 0;
        const uint32_t sp = smem_u32(smem + B_OFF_P), sds = smem_u32(smem + B_OFF_DS);
        Trace tr = p.tr;
        if (blockIdx.x != 0 || lane != 0) tr.buf = nullptr;
        int tn = 0;
        for (int w = blockIdx.x; w < p.total; w += gridDim.x, ++item_n) {
            int b, g, kt;
            unsigned qmask, lastmask;
            if (!bwd_decode<KIND>(w, p, b, g, kt, qmask, lastmask)) break;
            const int N = 2 * __popc(qmask);
            const int ks = item_n & 1;
            const uint32_t sk = smem_u32(smem + B_OFF_KV + ks * B_KV_STAGE), sv = sk + TILE_BYTES;
            mbar_wait(&kv_full[ks], (item_n >> 1) & 1);
            {   // first step of the item: S and dP (the previous item's last step has released both buffers)
                const int st = step_n & 1;
                mbar_wait(&qdo_full[st], (step_n >> 1) & 1);
                if (step_n > 0) mbar_wait(sdp_free, (step_n - 1) & 1);
                tc_fence_after();
                const uint32_t sq = smem_u32(smem + B_OFF_QDO + st * B_QDO_STAGE);
                issue_nt_128x128x64(tmem_base + T_S, sq, sk, nullptr);
                issue_nt_128x128x64(tmem_base + T_DP, sq + TILE_BYTES, sv, sdp_full);
            }
            for (int n = 0; n < N; ++n, ++step_n) {
                const int st = step_n & 1;
                const uint32_t sq = smem_u32(smem + B_OFF_QDO + st * B_QDO_STAGE);
                const uint32_t sqn = smem_u32(smem + B_OFF_QDO + (st ^ 1) * B_QDO_STAGE);
                trace_pt(tr, 0, tn, 1);
                if (n + 1 < N) {
                    mbar_wait(&qdo_full[st ^ 1], ((step_n + 1) >> 1) & 1);
                    mbar_wait(sdp_free, step_n & 1);
                    trace_pt(tr, 0, tn, 2);
                    tc_fence_after();
                    issue_nt_128x128x64(tmem_base + T_S, sqn, sk, nullptr);                  // S(n+1) = Q K^T
                    issue_nt_128x128x64(tmem_base + T_DP, sqn + TILE_BYTES, sv, sdp_full);   // dP(n+1) = dO V^T
                }
                trace_pt(tr, 0, tn, 3);
                mbar_wait(pds_full, step_n & 1);
                trace_pt(tr, 0, tn, 4);
                if (n == 0 && item_n > 0) mbar_wait(dkv_free, (item_n - 1) & 1);  // dK/dV of the previous item drained
                tc_fence_after();
                issue_tn_128x64x128(tmem_base + T_DV, sp, sq + TILE_BYTES, n > 0, nullptr);  // dV += P^T dO
                // Q / dO of this step are dead once dK is done: dQ reads dS and K only
                issue_tn_128x64x128(tmem_base + T_DK, sds, sq, n > 0, &qdo_free[st]);        // dK += dS^T Q
                const uint32_t db = step_n & 1, du = step_n >> 1;
                if (du > 0) {
                    mbar_wait(&dq_free[db], (du - 1) & 1);
                    tc_fence_after();
                }
                const bool last = n + 1 == N;
                trace_pt(tr, 0, tn, 5);
                issue_nn_128x64x128(tmem_base + T_DQ + db * 64, sds, sk, false, &dq_full[db], pds_free,  // dQ = dS K
                                    last ? dkv_full : nullptr, last ? &kv_free[ks] : nullptr);
            }
        }
    } else if (warp == 2 || warp == 3) {
        // ===================== converter warps: finished dQ tiles, fp32 running sum (L2) -> bf16 rows =====================
        // Runs one item behind the drain warps, off every critical path: the drain thread publishes the number of items
        // whose accumulator writes have landed; a tile is converted by the item that made its last contribution.
        const int t = (int)threadIdx.x - 64;
        uint32_t item_n = 0;
        Trace tr = p.tr;
        if (blockIdx.x != 0 || t != 0) tr.buf = nullptr;
        int tn = 0;
        for (int w = blockIdx.x; w < p.total; w += gridDim.x, ++item_n) {
            int b, g, kt;
            unsigned qmask, lastmask;
            if (!bwd_decode<KIND>(w, p, b, g, kt, qmask, lastmask)) break;
            if (lastmask == 0u) continue;
            trace_pt(tr, 2, tn, 40);
            while (*reinterpret_cast<volatile uint32_t*>(conv_flag) < item_n + 1) __nanosleep(256);
            __threadfence_block();
            trace_pt(tr, 2, tn, 41);
            for (int hh = 0; hh < 2; ++hh) {
                for (unsigned qm = lastmask; qm; qm &= qm - 1) {
                    const int qt = __ffs(qm) - 1;
                    const float* tile = p.dq_acc + (((long long)b * p.n_q + 2 * g + hh) * p.q_tiles + qt) * DQ_TILE_FLOATS;
                    bf16* out = p.dq + ((long long)b * p.L + qt * BT) * p.ld_d + (2 * g + hh) * D;
                    const int rows = min(BT, p.L - qt * BT);
                    // two batches of eight 32-byte chunks per thread: all loads of a batch in flight before the first store
                    // (the compiler cannot hoist loads over the stores on its own: the pointers may alias)
#pragma unroll 1
                    for (int it0 = 0; it0 < 16; it0 += 8) {
                        float4 v[8][2];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int id = (it0 + j) * 64 + t, r = id >> 3, c8 = id & 7;   // 8 threads per 128-byte output row
                            const float* src = tile + (c8 >> 2) * (DQ_TILE_FLOATS / 2) + r * 32;
                            const int ch = (c8 & 3) * 2;
                            if (r < rows) {
                                v[j][0] = __ldcg(reinterpret_cast<const float4*>(src + ((ch ^ (r & 7)) << 2)));
                                v[j][1] = __ldcg(reinterpret_cast<const float4*>(src + (((ch + 1) ^ (r & 7)) << 2)));
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int id = (it0 + j) * 64 + t, r = id >> 3, c8 = id & 7;
                            if (r < rows) {
                                const float f[8] = {v[j][0].x, v[j][0].y, v[j][0].z, v[j][0].w,
                                                    v[j][1].x, v[j][1].y, v[j][1].z, v[j][1].w};
                                *reinterpret_cast<bf16x8*>(out + (long long)r * p.ld_d + c8 * 8) = float_to_bf16x8(f);
                            }
                        }
                    }
                }
            }
        }
    } else if (warp >= 4 && warp < 12) {
        // ===================== softmax / dS warps: warpgroup k (warps 4+4k..7+4k) owns key columns [64k, 64k+64) of the
        // tile, processed as two 32-key blocks; one query row per thread =====================
        const int wgi = (warp - 4) >> 2;
        const int wq = warp & 3;
        const int row = wq * 32 + lane;
        const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
        const uint32_t t_s = tmem_base + lane_off + T_S + wgi * 64;
        const uint32_t t_dp = tmem_base + lane_off + T_DP + wgi * 64;
        // P / dS live as two 64-key halves; a 32-key block is four 16-byte chunks of this thread's 128-byte row
        const uint32_t sP = smem_u32(smem + B_OFF_P + wgi * TILE_BYTES) + row * 128;
        const uint32_t sDS = smem_u32(smem + B_OFF_DS + wgi * TILE_BYTES) + row * 128;
        uint32_t item_n = 0, step_n = 0;
        Trace tr = p.tr;
        if (blockIdx.x != 0 || lane != 0 || wq != 0 || wgi != 0) tr.buf = nullptr;
        int tn = 0;
        const int trole = 1 + wgi;
        for (int w = blockIdx.x; w < p.total; w += gridDim.x, ++item_n) {
            int b, g, kt;
            unsigned qmask, lastmask;
            if (!bwd_decode<KIND>(w, p, b, g, kt, qmask, lastmask)) break;
            const int ks = item_n & 1;
            mbar_wait(&kv_full[ks], (item_n >> 1) & 1);  // key codes
            const uint8_t* meta = smem + B_OFF_KV + ks * B_KV_STAGE + 2 * TILE_BYTES;
            const uint32_t ka = smem_u32(meta) + wgi * 256, kss = ka + 512;
            const int4 bsum0 = reinterpret_cast<const int4*>(meta + 1024)[wgi * 2];
            const int4 bsum1 = reinterpret_cast<const int4*>(meta + 1024)[wgi * 2 + 1];
            const int jbase = kt * BT + wgi * 64;
            for (int hh = 0; hh < 2; ++hh) {
                for (unsigned qm = qmask; qm; qm &= qm - 1, ++step_n) {
                    const int qt = __ffs(qm) - 1;
                    const int i = qt * BT + row;
                    // per-row scalars travel with the Q / dO stage (padded copies: lse' = +inf, dsum' = 0 past L)
                    const int st = step_n & 1;
                    trace_pt(tr, trole, tn, 10);
                    // S / dP of the step were issued after the Q / dO stage (row scalars included) had landed
                    mbar_wait(sdp_full, step_n & 1);
                    tc_fence_after();
                    const float* rowf = reinterpret_cast<const float*>(smem + B_OFF_QDO + st * B_QDO_STAGE + 2 * TILE_BYTES);
                    const float lse_i = rowf[row], dsum_i = rowf[128 + row];
                    const int act_i = reinterpret_cast<const int*>(rowf)[256 + row];
                    const int sess_i = reinterpret_cast<const int*>(rowf)[384 + row];
                    const uint32_t* keep_s = reinterpret_cast<const uint32_t*>(rowf) + 512 + wgi * 2 * BT + row;
                    const bool uni = (i < p.L) && (lse_i == INFINITY);
                    const bool wuni = __any_sync(0xffffffffu, uni);
                    const int istart = (i / p.P) * p.P;
                    const WarpRange wr = warp_range<KIND>(i, p.L, p.P, act_i, sess_i);
                    const float pu = uni ? p.inv_L : 0.f;
                    const float neg_lse = -lse_i;       // -inf on uniform / padding rows: exp2 -> 0
                    const float neg_dsum = -dsum_i;
                    trace_pt(tr, trole, tn, 12);
#pragma unroll 1
                    for (int bk = 0; bk < 2; ++bk) {
                        const int j0 = jbase + bk * 32;
                        int mode = classify_block<KIND>(bk == 0 ? bsum0 : bsum1, wr, j0);
                        // uniform rows put 1/L on every key below L, whatever the predicate says
                        if (wuni && j0 < p.L && (mode == BLK_SKIP || mode == BLK_FULL)) mode = BLK_MASK_DIAG;
                        uint32_t pk[16], ds[16];
                        if (mode == BLK_SKIP) {
#pragma unroll
                            for (int x = 0; x < 16; ++x) pk[x] = ds[x] = 0u;
                            if (bk == 1) {   // nothing to read from TMEM: release S and dP
                                tc_fence_before();
                                mbar_arrive(sdp_free);
                            }
                        } else {
                            uint32_t s[32], dp[32];
                            tmem_ld_32x32(t_s + bk * 32, s);
                            tmem_ld_32x32(t_dp + bk * 32, dp);
                            tmem_ld_wait();
                            trace_pt(tr, trole, tn, 13);
                            if (bk == 1) {
                                tc_fence_before();
                                mbar_arrive(sdp_free);
                            }
                            uint32_t kw = 0xffffffffu;
                            if constexpr (DROP) kw = uni ? 0xffffffffu : keep_s[bk * BT];
                            const int lim_u = uni ? (p.L - j0) : 0;
                            if (mode == BLK_FULL)
                                bwd_block<KIND, DROP, BLK_FULL>(s, dp, pk, ds, ka + bk * 128, kss + bk * 128, act_i, sess_i, j0, i,
                                                                istart, p.scale_log2, neg_lse, neg_dsum, pu, lim_u, kw);
                            else if (mode == BLK_MASK)
                                bwd_block<KIND, DROP, BLK_MASK>(s, dp, pk, ds, ka + bk * 128, kss + bk * 128, act_i, sess_i, j0, i,
                                                                istart, p.scale_log2, neg_lse, neg_dsum, pu, lim_u, kw);
                            else
                                bwd_block<KIND, DROP, BLK_MASK_DIAG>(s, dp, pk, ds, ka + bk * 128, kss + bk * 128, act_i, sess_i, j0,
                                                                     i, istart, p.scale_log2, neg_lse, neg_dsum, pu, lim_u, kw);
                        }
                        trace_pt(tr, trole, tn, 14);
                        if (bk == 0 && step_n > 0) mbar_wait(pds_free, (step_n - 1) & 1);
                        trace_pt(tr, trole, tn, 15);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint32_t off = (uint32_t)(((bk * 4 + q) ^ (row & 7)) << 4);
                            sts128(sP + off, pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                            sts128(sDS + off, ds[4 * q], ds[4 * q + 1], ds[4 * q + 2], ds[4 * q + 3]);
                        }
                    }
                    fence_proxy_async();
                    mbar_arrive(pds_full);
                    trace_pt(tr, trole, tn, 16);
                }
            }
        }
    } else if (warp >= 12) {
        // ===================== drain warps: dQ tiles -> TMA fp32 reduce-add; dK / dV -> bf16 TMA store ================
        const int wq = warp & 3;
        const int row = wq * 32 + lane;
        const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
        const uint32_t t_dq = tmem_base + lane_off + T_DQ;
        uint8_t* stg = smem + B_OFF_STG;
        const uint32_t stg_row = smem_u32(stg) + row * 128;
        uint32_t item_n = 0, step_n = 0;
        Trace tr = p.tr;
        if (blockIdx.x != 0 || row != 0) tr.buf = nullptr;
        int tn = 0;
        for (int w = blockIdx.x; w < p.total; w += gridDim.x, ++item_n) {
            int b, g, kt;
            unsigned qmask, lastmask;
            if (!bwd_decode<KIND>(w, p, b, g, kt, qmask, lastmask)) break;
            for (int hh = 0; hh < 2; ++hh) {
                for (unsigned qm = qmask; qm; qm &= qm - 1, ++step_n) {
                    const int qt = __ffs(qm) - 1;
                    const uint32_t db = step_n & 1, du = step_n >> 1;
                    trace_pt(tr, 3, tn, 30);
                    float* dst = p.dq_acc + (((long long)b * p.n_q + 2 * g + hh) * p.q_tiles + qt) * DQ_TILE_FLOATS;
                    const bool first = kt == 0;
                    mbar_wait(&dq_full[db], du & 1);
                    trace_pt(tr, 3, tn, 31);
                    tc_fence_after();
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
                            // key tile 0 opens the running sum with a plain store (no zero fill of the accumulator)
                            if (first) bulk_store_1d(dst + half * (DQ_TILE_FLOATS / 2), stg, DQ_TILE_FLOATS * 2);
                            else bulk_reduce_add_f32(dst + half * (DQ_TILE_FLOATS / 2), stg, DQ_TILE_FLOATS * 2);
                            bulk_commit();
                        }
                    }
                    trace_pt(tr, 3, tn, 32);
                }
            }
            // ---- item epilogue: dK then dV -> bf16 -> staging -> TMA store (rows past L are clipped)
            mbar_wait(dkv_full, item_n & 1);
            tc_fence_after();
#pragma unroll 1
            for (int which = 0; which < 2; ++which) {
                const uint32_t t_acc = tmem_base + lane_off + (which == 0 ? T_DK : T_DV);
                const float mul = (which == 0) ? p.scale : 1.f;
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
            // The item's dQ stores / reduce-adds (every bulk group but the two stores just committed) must have landed before
            // the next key tile of the group adds to the same accumulator tiles, and before the converter warps read the
            // tiles this item completed.
            if (row == 0) {
                bulk_wait_group2();
                fence_proxy_async_all();
                __threadfence_block();
                *reinterpret_cast<volatile uint32_t*>(conv_flag) = item_n + 1;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// =================================================================================================================
// decode: one new token per beam row against the user's prompt K/V (constrained beam search, generation.py)
// =================================================================================================================
// CTA = (user, kv head).  The 2 x beams query vectors of the group (head hh of the group at tile rows [64 hh, 64 hh +
// beams)) meet the user's prompt keys — stored once per user, shared by all beams — as tcgen05 tiles: S = Q K^T over 64
// keys into TMEM, online softmax with one query row per thread, P (bf16) through shared memory, O += P V.  The beam's own
// generated keys (<= 4, reached through the ancestry table) are folded in by the row's thread in the epilogue.
constexpr int DC_KT = 64;
constexpr int DC_THREADS = 192;                          // 4 softmax warps | MMA issuer | TMA producer
constexpr int DC_OFF_Q = 0;                              // [128 x 64]
constexpr int DC_OFF_K = TILE_BYTES;                     // 2 stages of [64 x 64]
constexpr int DC_OFF_V = DC_OFF_K + 2 * S_HALF;
constexpr int DC_OFF_P = DC_OFF_V + 2 * S_HALF;          // [128 x 64] bf16, K-major
constexpr int DC_OFF_OK = DC_OFF_P + TILE_BYTES;         // one bit per prompt key (<= 4096 keys)
constexpr int DC_OFF_TL = DC_OFF_OK + 512;               // list of key tiles that hold an allowed key (<= 64) + count
constexpr int DC_OFF_BAR = DC_OFF_TL + 512;
constexpr int DC_SMEM = DC_OFF_BAR + 128 + 1024;

struct DecParams {
    const bf16* qcur;        // [R, ld_g], q head h at column h*64
    const bf16* gen_k;       // generated keys: step s, slot r at gen_k + s*gen_step_stride + r*ld_g (+ kvh*64)
    const bf16* gen_v;
    long long gen_step_stride, ld_g;
    const int* anc;          // [R, S_max] slot of this row's ancestor at step s
    int B, beams, L0, n_gen, n_q, n_kv, S_max;
    const int* am;           // [B, L0]
    const int* act;          // [B, L0] or nullptr
    const int* sess;         // [B, L0] or nullptr
    int kind;
    const float* vmean;      // [B, n_kv, 64] mean of ALL L0 prompt values (cross kinds)
    float scale_log2;
    bf16* o;                 // [R, ld_o]
    long long ld_o;
};

__global__ void __launch_bounds__(DC_THREADS, 3)
attn_decode_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, DecParams a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DC_OFF_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;   // [2]
    uint64_t* v_full = bars + 3;   // [2]
    uint64_t* s_full = bars + 5;
    uint64_t* s_free = bars + 6;   // all softmax threads have read S
    uint64_t* p_full = bars + 7;   // P stored, O rescaled
    uint64_t* pv_done = bars + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
    uint32_t* ok_bits = reinterpret_cast<uint32_t*>(smem + DC_OFF_OK);
    int* tiles = reinterpret_cast<int*>(smem + DC_OFF_TL);   // [0] = count, [1..] = tile indices

    const int u = blockIdx.x, kvh = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool cross = (a.kind == MASK_MULTI_CROSS || a.kind == MASK_SESSION_CROSS);
    const int kt_all = (a.L0 + DC_KT - 1) / DC_KT;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmQ);
        prefetch_tmap(&tmK);
        prefetch_tmap(&tmV);
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&v_full[i], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_free, 128);
        mbar_init(p_full, 128);
        mbar_init(pv_done, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc<128>(tmem_slot);
    {   // allowed-key bits of the new token's row against the prompt (Qwen3Multi/model.py:605-617, 717-728): every valid
        // key, restricted by behaviour level / session of the last prompt token for the cross kinds
        int act_last = 0, sess_last = 0;
        if (cross) {
            act_last = a.act[(long long)u * a.L0 + a.L0 - 1];
            if (a.sess) sess_last = a.sess[(long long)u * a.L0 + a.L0 - 1];
        }
        for (int j0 = warp * 32; j0 < kt_all * DC_KT; j0 += DC_THREADS) {
            const int j = j0 + lane;
            int ok = 0;
            if (j < a.L0) {
                const long long idx = (long long)u * a.L0 + j;
                ok = a.am[idx];
                if (cross) {
                    ok = ok && (a.act[idx] < act_last);
                    if (a.kind == MASK_SESSION_CROSS) ok = ok && (a.sess[idx] < sess_last);
                }
            }
            const unsigned w = __ballot_sync(0xffffffffu, ok != 0);
            if (lane == 0) ok_bits[j0 >> 5] = w;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {   // key tiles with at least one allowed key
        int n = 0;
        for (int t = 0; t < kt_all; ++t)
            if ((ok_bits[2 * t] | ok_bits[2 * t + 1]) != 0u) tiles[1 + n++] = t;
        tiles[0] = n;
    }
    __syncthreads();
    const uint32_t tmem_base = *tmem_slot;
    const int nkt = tiles[0];
    const uint32_t sq = smem_u32(smem + DC_OFF_Q);
    const uint32_t sp = smem_u32(smem + DC_OFF_P);
    constexpr uint32_t T_O = 64;

    if (warp == 5) {
        // ===================== TMA producer =====================
        if (lane == 0 && nkt > 0) {
            mbar_expect_tx(q_full, 2 * a.beams * 128);
            tma_load_2d(smem + DC_OFF_Q, &tmQ, q_full, (2 * kvh) * D, u * a.beams);
            tma_load_2d(smem + DC_OFF_Q + 64 * 128, &tmQ, q_full, (2 * kvh + 1) * D, u * a.beams);
            for (int n = 0; n < nkt; ++n) {
                const int st = n & 1, t = tiles[1 + n];
                if (n >= 2) mbar_wait(pv_done, n & 1);   // PV(n-2) done: both its K (S(n-2) came first) and V are dead
                mbar_expect_tx(&k_full[st], S_HALF);
                tma_load_3d(smem + DC_OFF_K + st * S_HALF, &tmK, &k_full[st], kvh * D, t * DC_KT, u);
                mbar_expect_tx(&v_full[st], S_HALF);
                tma_load_3d(smem + DC_OFF_V + st * S_HALF, &tmV, &v_full[st], kvh * D, t * DC_KT, u);
            }
        }
    } else if (warp == 4) {
        // ===================== MMA issuer (converged warp, elected lane issues) =====================
        auto issue_s = [&](int n) {   // S(n)[128 x 64] = Q K_n^T
            constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
            mbar_wait(&k_full[n & 1], (n >> 1) & 1);
            if (n > 0) mbar_wait(s_free, (n - 1) & 1);
            tc_fence_after();
            const uint64_t da = desc_k(sq), db = desc_k(smem_u32(smem + DC_OFF_K + (n & 1) * S_HALF));
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, k != 0);
                umma_commit(s_full);
            }
            __syncwarp();
        };
        if (nkt > 0) {
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int n = 0; n < nkt; ++n) {
                if (n + 1 < nkt) issue_s(n + 1);
                mbar_wait(p_full, n & 1);
                mbar_wait(&v_full[n & 1], (n >> 1) & 1);
                tc_fence_after();
                constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);
                const uint64_t da = desc_k(sp), db = desc_mn(smem_u32(smem + DC_OFF_V + (n & 1) * S_HALF), S_HALF);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_base + T_O, da + 2 * k, db + k * 128, idesc, (n > 0 || k != 0) ? 1u : 0u);
                    umma_commit(pv_done);
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== softmax warps: one query row per thread =====================
        const int row = warp * 32 + lane;
        const int beam = row & 63, hh = row >> 6;
        const bool valid = beam < a.beams;
        const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
        float m = -INFINITY, l = 0.f;
        for (int n = 0; n < nkt; ++n) {
            const int t = tiles[1 + n];
            mbar_wait(s_full, n & 1);
            tc_fence_after();
            uint32_t s0[32], s1[32];
            tmem_ld_32x32(t_row, s0);
            tmem_ld_32x32(t_row + 32, s1);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(s_free);
            const uint32_t w0 = ok_bits[2 * t], w1 = ok_bits[2 * t + 1];
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const float v0 = ((w0 >> c) & 1u) ? __uint_as_float(s0[c]) : -INFINITY;
                const float v1 = ((w1 >> c) & 1u) ? __uint_as_float(s1[c]) : -INFINITY;
                s0[c] = __float_as_uint(v0);
                s1[c] = __float_as_uint(v1);
                mx = max3(mx, v0, v1);
            }
            const float m_new = fmaxf(m, mx * a.scale_log2);
            const float alpha = (m == -INFINITY) ? 0.f : ex2_approx(m - m_new);   // (m_new finite: the tile has an allowed key)
            const float neg_m = -m_new;
            float sum = 0.f;
            uint32_t pk[32];
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
                const float e0 = ex2_approx(fmaf(__uint_as_float(s0[c]), a.scale_log2, neg_m));
                const float e1 = ex2_approx(fmaf(__uint_as_float(s0[c + 1]), a.scale_log2, neg_m));
                const float f0 = ex2_approx(fmaf(__uint_as_float(s1[c]), a.scale_log2, neg_m));
                const float f1 = ex2_approx(fmaf(__uint_as_float(s1[c + 1]), a.scale_log2, neg_m));
                sum += (e0 + e1) + (f0 + f1);
                pk[c >> 1] = pack_bf16(e0, e1);
                pk[16 + (c >> 1)] = pack_bf16(f0, f1);
            }
            l = l * alpha + sum;
            m = m_new;
            if (n > 0) {   // PV(n-1) has read P and written O: rescale O, then overwrite P
                mbar_wait(pv_done, (n - 1) & 1);
                tc_fence_after();
                if (__any_sync(0xffffffffu, alpha != 1.f)) rescale_o_row64(t_row + T_O, alpha);
            }
#pragma unroll
            for (int ch = 0; ch < 8; ++ch)
                sts128(sp + row * 128 + ((ch ^ (row & 7)) << 4), pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(p_full);
        }
        if (nkt > 0) {
            mbar_wait(pv_done, (nkt - 1) & 1);
            tc_fence_after();
        }
        // ---- epilogue: the beam's own generated keys (current token included), normalisation, store
        float o[64];
        {
            uint32_t r0[32], r1[32];
            if (nkt > 0) {
                tmem_ld_32x32(t_row + T_O, r0);
                tmem_ld_32x32(t_row + T_O + 32, r1);
                tmem_ld_wait();
            }
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                o[c] = nkt > 0 ? __uint_as_float(r0[c]) : 0.f;
                o[32 + c] = nkt > 0 ? __uint_as_float(r1[c]) : 0.f;
            }
        }
        if (valid) {
            const int r = u * a.beams + beam, h = kvh * 2 + hh;
            const bf16* qp = a.qcur + (long long)r * a.ld_g + h * D;
            if (!cross) {                       // generated columns are masked for the cross rows (model.py:605-617)
                for (int s = 0; s < a.n_gen; ++s) {
                    const int slot = a.anc[(long long)r * a.S_max + s];
                    const bf16* kp = a.gen_k + s * a.gen_step_stride + (long long)slot * a.ld_g + kvh * D;
                    const bf16* vp = a.gen_v + s * a.gen_step_stride + (long long)slot * a.ld_g + kvh * D;
                    float sc = 0.f;
#pragma unroll
                    for (int c8 = 0; c8 < 8; ++c8) {
                        float q8[8], k8[8];
                        bf16x8_to_float(*reinterpret_cast<const bf16x8*>(qp + c8 * 8), q8);
                        bf16x8_to_float(*reinterpret_cast<const bf16x8*>(kp + c8 * 8), k8);
#pragma unroll
                        for (int e = 0; e < 8; ++e) sc = fmaf(q8[e], k8[e], sc);
                    }
                    sc *= a.scale_log2;
                    const float m_new = fmaxf(m, sc);
                    const float alpha = (m == -INFINITY) ? 0.f : ex2_approx(m - m_new), pe = ex2_approx(sc - m_new);
                    l = l * alpha + pe;
#pragma unroll
                    for (int c8 = 0; c8 < 8; ++c8) {
                        float v8[8];
                        bf16x8_to_float(*reinterpret_cast<const bf16x8*>(vp + c8 * 8), v8);
#pragma unroll
                        for (int e = 0; e < 8; ++e) o[c8 * 8 + e] = o[c8 * 8 + e] * alpha + pe * v8[e];
                    }
                    m = m_new;
                }
            }
            if (l > 0.f) {
                const float inv = 1.f / l;
#pragma unroll
                for (int c = 0; c < 64; ++c) o[c] *= inv;
            } else {
                // no allowed key: uniform over ALL cached keys (quirk Q1) = (L0 * mean(prompt V) + sum gen V) / (L0 + n_gen)
                const float* vm = a.vmean + ((long long)u * a.n_kv + kvh) * D;
                const float inv = 1.0f / (float)(a.L0 + a.n_gen);
#pragma unroll
                for (int c = 0; c < 64; ++c) o[c] = vm[c] * (float)a.L0;
                for (int s = 0; s < a.n_gen; ++s) {
                    const int slot = a.anc[(long long)r * a.S_max + s];
                    const bf16* vp = a.gen_v + s * a.gen_step_stride + (long long)slot * a.ld_g + kvh * D;
#pragma unroll
                    for (int c8 = 0; c8 < 8; ++c8) {
                        float v8[8];
                        bf16x8_to_float(*reinterpret_cast<const bf16x8*>(vp + c8 * 8), v8);
#pragma unroll
                        for (int e = 0; e < 8; ++e) o[c8 * 8 + e] += v8[e];
                    }
                }
#pragma unroll
                for (int c = 0; c < 64; ++c) o[c] *= inv;
            }
            bf16* op = a.o + (long long)r * a.ld_o + h * D;
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) *reinterpret_cast<bf16x8*>(op + c8 * 8) = float_to_bf16x8(o + c8 * 8);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<128>(tmem_base);
}

// Padded per-row inputs of the backward: dsum_p[b,h,i] = keep_prob * sum_d dO*O and lse_p[b,h,i] = lse + log2(keep_prob)
// (so that exp2(s - lse_p) = P / keep_prob); uniform rows keep dsum unscaled and lse = +inf; rows i >= L: 0 / +inf.
// uni_bits[b] |= 1 << (i / 128) for uniform rows (head 0 decides: the mask is head-independent).
constexpr int PREP_G = 4;   // (row, head) groups per 8-lane team: all eight 16-byte loads in flight before the first use
__global__ void attn_bwd_prep_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o, long long ld_o, int B, int L,
                                     int Lp, int n_q, const float* __restrict__ lse, float keep_prob, float log2_keep,
                                     float* __restrict__ dsum_p, float* __restrict__ lse_p, unsigned* __restrict__ uni_bits) {
    // grid = (teams of one sequence / 32, B); 8 threads per (padded row, head) group, PREP_G consecutive groups per team,
    // 32-bit index math only
    const int b = blockIdx.y;
    const int sub = threadIdx.x & 7;
    const int team = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 3);
    const int n_grp = Lp * n_q;
    bf16x8 av[PREP_G], dv[PREP_G];
#pragma unroll
    for (int j = 0; j < PREP_G; ++j) {
        const int grp = team * PREP_G + j;
        const bool live = grp < n_grp;
        const int i = live ? grp / n_q : 0;
        const int h = live ? grp - i * n_q : 0;
        const long long row = (live && i < L) ? (long long)b * L + i : 0;
        av[j] = *reinterpret_cast<const bf16x8*>(o + row * ld_o + h * D + sub * 8);
        dv[j] = *reinterpret_cast<const bf16x8*>(d_o + row * ld_o + h * D + sub * 8);
    }
#pragma unroll
    for (int j = 0; j < PREP_G; ++j) {
        const int grp = team * PREP_G + j;
        const bool live = grp < n_grp;
        const int i = live ? grp / n_q : 0;
        const int h = live ? grp - i * n_q : 0;
        const bool real = live && i < L;
        float a[8], d[8];
        bf16x8_to_float(av[j], a);
        bf16x8_to_float(dv[j], d);
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
                if (ls == INFINITY) {
                    if (h == 0) atomicOr(&uni_bits[b], 1u << (i / BT));
                } else {
                    ls += log2_keep;
                    s *= keep_prob;
                }
            }
            dsum_p[pi] = real ? s : 0.f;
            lse_p[pi] = ls;
        }
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
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}

inline long long align256(long long x) { return (x + 255) / 256 * 256; }

struct MetaLayout {
    int Lp, k_tiles;
    long long off_ka, off_ks, off_qa, off_qs, off_blk, bytes;
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
    m.off_blk = 4 * arr;
    m.bytes = m.off_blk + align256((long long)B * (m.Lp / 32) * 16);
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
                                                        reinterpret_cast<int4*>(ws + ml.off_blk));
    GAMER_LAUNCH_CHECK();
    return 0;
}

template <int KIND, bool DROP>
int launch_fwd_t(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to,
                 const FwdParams& p, cudaStream_t stream) {
    static PerDeviceOnce cfg;
    if (cfg.need())
        GAMER_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<KIND, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, S_SMEM));
    attn_fwd_kernel<KIND, DROP><<<p.B * p.n_q * p.q_tiles, S_THREADS, S_SMEM, stream>>>(tq, tk, tv, to, p);
    GAMER_LAUNCH_CHECK();
    return 0;
}
template <int KIND>
int launch_fwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to,
               const FwdParams& p, cudaStream_t stream) {
    return p.drop.thresh ? launch_fwd_t<KIND, true>(tq, tk, tv, to, p, stream)
                         : launch_fwd_t<KIND, false>(tq, tk, tv, to, p, stream);
}

template <int KIND, bool DROP>
int launch_bwd_t(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
                 const CUtensorMap& tdk, const CUtensorMap& tdv, const BwdParams& p, cudaStream_t stream) {
    static PerDeviceOnce cfg;
    if (cfg.need())
        GAMER_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<KIND, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM));
    const int sms = sm_count();
    const int n_groups = p.B * p.n_kv;
    const int grid = n_groups < sms ? n_groups : sms;
    BwdParams pp = p;
    pp.total = (n_groups + grid - 1) / grid * grid * p.k_tiles;   // walk index bound: every CTA runs the same number of rounds
    attn_bwd_kernel<KIND, DROP><<<grid, B_THREADS, B_SMEM, stream>>>(tq, tk, tv, tdo, tdk, tdv, pp);
    GAMER_LAUNCH_CHECK();
    return 0;
}
template <int KIND>
int launch_bwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
               const CUtensorMap& tdk, const CUtensorMap& tdv, const BwdParams& p, cudaStream_t stream) {
    return p.drop_on ? launch_bwd_t<KIND, true>(tq, tk, tv, tdo, tdk, tdv, p, stream)
                     : launch_bwd_t<KIND, false>(tq, tk, tv, tdo, tdk, tdv, p, stream);
}

}  // namespace

// debug hook: subsequent launches record a timeline of CTA 0 into buf (4 roles x cap x (tag, clock) int64 pairs)
extern "C" int gamer_attn_set_trace(void* buf, int cap) {
    g_trace.buf = reinterpret_cast<long long*>(buf);
    g_trace.cap = cap;
    return 0;
}

bool attn_tc_supported(int L, int n_q, int n_kv, int head_dim) {
    return head_dim == D && n_kv > 0 && n_q == 2 * n_kv && L >= 1 && (L + BT - 1) / BT <= 32;
}

long long attn_tc_fwd_ws_bytes(int B, int L) { return meta_layout(B, L).bytes; }

long long attn_tc_keep_bytes(int B, int L, int n_q) {
    const MetaLayout ml = meta_layout(B, L);
    return (long long)B * n_q * ml.k_tiles * (ml.Lp / 32) * BT * 4;
}

int attn_tc_fwd(const void* q, const void* k, const void* v, long long ld, int B, int L, int n_q, int n_kv, int kind, int P,
                const int* am, const int* act, const int* sess, float scale, const float* vmean, void* ws, void* o,
                long long ld_o, float* lse, const gamer_dropout_t* drop, void* keep, cudaStream_t stream) {
    const MetaLayout ml = meta_layout(B, L);
    uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
    if (int e = build_meta(kind, am, act, sess, B, L, w8, ml, stream)) return e;
    CUtensorMap tq, tk, tv, to;
    if (int e = make_tmap_seq(&tq, q, B, L, n_q * D, ld)) return e;
    if (int e = make_tmap_seq(&tk, k, B, L, n_kv * D, ld, S_KT)) return e;
    if (int e = make_tmap_seq(&tv, v, B, L, n_kv * D, ld, S_KT)) return e;
    if (int e = make_tmap_seq(&to, o, B, L, n_q * D, ld_o)) return e;
    FwdParams p{};
    p.B = B; p.L = L; p.Lp = ml.Lp; p.n_q = n_q; p.n_kv = n_kv; p.P = P; p.q_tiles = ml.k_tiles;
    p.ka = reinterpret_cast<const int*>(w8 + ml.off_ka);
    p.ks = reinterpret_cast<const int*>(w8 + ml.off_ks);
    p.blk = reinterpret_cast<const int4*>(w8 + ml.off_blk);
    p.act = act; p.sess = sess; p.scale_log2 = scale * 1.4426950408889634f; p.vmean = vmean; p.lse = lse;
    p.keep = reinterpret_cast<uint32_t*>(keep);
    p.drop = make_drop(drop, 8);
    p.tr = g_trace;
    GAMER_REQUIRE(p.drop.thresh == 0 || keep != nullptr,
                  "attention dropout needs the keep-word buffer (gamer_attn_keep_bytes) the backward reads back");
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
                const void* keep, cudaStream_t stream) {
    const BwdLayout bl = bwd_layout(B, L, n_q);
    uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
    if (int e = build_meta(kind, am, act, sess, B, L, w8, bl.ml, stream)) return e;
    float* dsum = reinterpret_cast<float*>(w8 + bl.off_dsum);
    float* lse_p = reinterpret_cast<float*>(w8 + bl.off_lse);
    unsigned* uni = reinterpret_cast<unsigned*>(w8 + bl.off_uni);
    float* acc = reinterpret_cast<float*>(w8 + bl.off_acc);
    const DropParams dp = make_drop(drop, 8);
    GAMER_REQUIRE(dp.thresh == 0 || keep != nullptr, "attention dropout: the backward needs the forward's keep words");
    // (the dQ accumulator needs no zero fill: the first contribution to a tile is a store)
    GAMER_CHECK_CUDA(cudaMemsetAsync(uni, 0, (size_t)B * 4, stream));
    {
        const float keep_prob = 1.0f / dp.scale;
        const dim3 grid(((bl.ml.Lp * n_q + PREP_G - 1) / PREP_G * 8 + 255) / 256, B);
        attn_bwd_prep_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(o), reinterpret_cast<const bf16*>(d_o),
                                                       ld_o, B, L, bl.ml.Lp, n_q, lse, keep_prob, log2f(keep_prob), dsum,
                                                       lse_p, uni);
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
    p.blk = reinterpret_cast<const int4*>(w8 + bl.ml.off_blk);
    p.qa = reinterpret_cast<const int*>(w8 + bl.ml.off_qa);
    p.qs = reinterpret_cast<const int*>(w8 + bl.ml.off_qs);
    p.lse_p = lse_p; p.dsum_p = dsum; p.uni_bits = uni;
    p.keep = reinterpret_cast<const uint32_t*>(keep);
    p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f; p.inv_L = 1.0f / (float)L; p.dq_acc = acc;
    p.dq = reinterpret_cast<bf16*>(dq); p.ld_d = ld_d;
    p.drop_on = dp.thresh != 0;
    p.tr = g_trace;
    int e;
    switch (kind) {
        case 0: e = launch_bwd<0>(tq, tk, tv, tdo, tdk, tdv, p, stream); break;
        case 1: e = launch_bwd<1>(tq, tk, tv, tdo, tdk, tdv, p, stream); break;
        case 2: e = launch_bwd<2>(tq, tk, tv, tdo, tdk, tdv, p, stream); break;
        default: e = launch_bwd<3>(tq, tk, tv, tdo, tdk, tdv, p, stream); break;
    }
    if (e) return e;
    return 0;
}

int attn_tc_decode(const void* qcur, const void* pk, const void* pv, long long ld_p, const void* gen_k, const void* gen_v,
                   long long gen_step_stride, long long ld_g, const int* anc, int B, int beams, int L0, int n_gen, int n_q,
                   int n_kv, int S_max, const int* am, const int* act, const int* sess, int kind, const float* vmean,
                   float scale, void* o, long long ld_o, cudaStream_t stream) {
    CUtensorMap tq, tk, tv;
    {   // q rows of this step: [R, ld_g] viewed as one "sequence"; box = [64 columns x beams rows]
        EncodeTiledFn fn = encode_fn();
        GAMER_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
        GAMER_REQUIRE((reinterpret_cast<uintptr_t>(qcur) & 15) == 0 && (ld_g * 2) % 16 == 0, "decode q rows must be 16-byte aligned");
        cuuint64_t dims[2] = {(cuuint64_t)n_q * D, (cuuint64_t)B * beams};
        cuuint64_t strides[1] = {(cuuint64_t)ld_g * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)beams};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = fn(&tq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(qcur), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        GAMER_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (decode q) failed with %d", (int)r);
    }
    if (int e = make_tmap_seq(&tk, pk, B, L0, n_kv * D, ld_p, DC_KT)) return e;
    if (int e = make_tmap_seq(&tv, pv, B, L0, n_kv * D, ld_p, DC_KT)) return e;
    DecParams a{};
    a.qcur = reinterpret_cast<const bf16*>(qcur);
    a.gen_k = reinterpret_cast<const bf16*>(gen_k); a.gen_v = reinterpret_cast<const bf16*>(gen_v);
    a.gen_step_stride = gen_step_stride; a.ld_g = ld_g; a.anc = anc;
    a.B = B; a.beams = beams; a.L0 = L0; a.n_gen = n_gen; a.n_q = n_q; a.n_kv = n_kv; a.S_max = S_max;
    a.am = am; a.act = act; a.sess = sess; a.kind = kind; a.vmean = vmean;
    a.scale_log2 = scale * 1.4426950408889634f;
    a.o = reinterpret_cast<bf16*>(o); a.ld_o = ld_o;
    static PerDeviceOnce cfg;
    if (cfg.need())
        GAMER_CHECK_CUDA(cudaFuncSetAttribute(attn_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DC_SMEM));
    attn_decode_kernel<<<dim3(B, n_kv), DC_THREADS, DC_SMEM, stream>>>(tq, tk, tv, a);
    GAMER_LAUNCH_CHECK();
    return 0;
}
