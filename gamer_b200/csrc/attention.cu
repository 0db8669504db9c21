// K6: session / behaviour-masked attention, forward and backward, GQA 6:3, head_dim 64.
//
// The multi-level mask is a predicate evaluated per (query i, key j) from three int arrays (attention mask, behaviour
// level `actions`, `session_ids`) — the reference materialises two [B,1,L,L] fp32 tensors instead
// (SeqRec/models/generative/Qwen3Multi/model.py:573-741, Qwen3SessionMoe/model.py:416-468):
//   CAUSAL        j<=i & am[j]
//   MULTI_CROSS   j<=i & act[j]<act[i] & am[j]
//   SESSION       (item(j)==item(i) & j<=i | sess[j]<sess[i]) & am[j]
//   SESSION_CROSS sess[j]<sess[i] & act[j]<act[i] & am[j]
// A query row with no allowed key is, in the reference, a softmax over an all-finfo.min row = uniform over ALL L keys
// (quirk Q1); here such rows take the column mean of V in the forward (lse = +inf marks them) and P = 1/L in the
// backward (the analytic gradient of that uniform softmax, = the reference's eager-attention gradient).
//
// Math runs on mma.sync.m16n8k16 bf16 tiles with fp32 accumulation and an online softmax in the exp2 domain.
#include <stdlib.h>

#include "attention_tc.cuh"
#include "common.cuh"

namespace {

constexpr int D = 64;        // head dim
constexpr int BQ = 64;       // query rows per CTA (4 warps x 16)
constexpr int BK = 64;       // keys per tile
constexpr int TILE_BYTES = 64 * 128;

struct AttnArgs {
    const bf16* q;   // [B*L, ld] + h*64
    const bf16* k;   // [B*L, ld] + kvh*64
    const bf16* v;
    long long ld;    // row stride of q/k/v (elements)
    int B, L, n_q, n_kv, P;
    const int* am;   // [B, L]
    const int* act;  // [B, L] or nullptr
    const int* sess; // [B, L] or nullptr
    float scale_log2;  // head_dim^-0.5 * log2(e)
    const float* vmean;  // [B, n_kv, 64]
    bf16* o;         // [B*L, ld_o] + h*64
    long long ld_o;
    float* lse;      // [B, n_q, L]   (log2 domain; +inf = uniform row)
    // backward
    const bf16* d_o;   // [B*L, ld_o]
    const float* dsum; // [B, n_q, L]  rowsum(dO * O)
    const int* uni_flag;  // [B, q_tiles]
    bf16* dq;        // [B*L, ld_d] + h*64
    bf16* dk;
    bf16* dv;
    long long ld_d;
    float scale;     // head_dim^-0.5
};

__device__ __forceinline__ uint32_t swz(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
    const int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// 64 rows x 64 bf16 tile (row stride ld) -> swizzled smem; rows >= n_valid are zero-filled.  128 threads.
__device__ __forceinline__ void load_tile(uint32_t sbase, const bf16* g, long long ld, int n_valid) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int idx = threadIdx.x + it * 128;
        const int row = idx >> 3, chunk = idx & 7;
        const bool ok = row < n_valid;
        cp_async16(sbase + swz(row, chunk), g + (long long)(ok ? row : 0) * ld + chunk * 8, ok);
    }
}

// A fragments (16 rows x 64 k) of rows [r0, r0+16) of a tile: f[kk][4]
__device__ __forceinline__ void load_a_frags(uint32_t sbase, int r0, uint32_t f[4][4]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
        ldsm_x4(f[kk][0], f[kk][1], f[kk][2], f[kk][3], sbase + swz(r0 + (lane & 15), kk * 2 + (lane >> 4)));
}

// acc[nb][4] (16 x 64) += A(16 x 64 k, frags) * B^T where the B tile is stored [n][k] row-major (n = 64 rows)
__device__ __forceinline__ void gemm_a_bnk(float acc[8][4], const uint32_t a[4][4], uint32_t sB) {
    const int lane = threadIdx.x & 31;
    const int mi = lane >> 3, lr = lane & 7;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4(b0, b1, b2, b3, sB + swz(np * 16 + (mi >> 1) * 8 + lr, kk * 2 + (mi & 1)));
            mma16816(acc[2 * np], a[kk], b0, b1);
            mma16816(acc[2 * np + 1], a[kk], b2, b3);
        }
    }
}

// acc[nb][4] (16 x 64 n) += A(16 x 64 k, frags) * B where the B tile is stored [k][n] row-major (k = 64 rows)
__device__ __forceinline__ void gemm_a_bkn(float acc[8][4], const uint32_t a[4][4], uint32_t sB) {
    const int lane = threadIdx.x & 31;
    const int mi = lane >> 3, lr = lane & 7;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4_t(b0, b1, b2, b3, sB + swz(kk * 16 + (mi & 1) * 8 + lr, np * 2 + (mi >> 1)));
            mma16816(acc[2 * np], a[kk], b0, b1);
            mma16816(acc[2 * np + 1], a[kk], b2, b3);
        }
    }
}

// C-layout accumulators (16 x 64) -> bf16 A fragments for a following GEMM whose k runs over these 64 columns
__device__ __forceinline__ void acc_to_a_frags(const float acc[8][4], uint32_t a[4][4]) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        a[kk][0] = pack_bf16(acc[2 * kk][0], acc[2 * kk][1]);
        a[kk][1] = pack_bf16(acc[2 * kk][2], acc[2 * kk][3]);
        a[kk][2] = pack_bf16(acc[2 * kk + 1][0], acc[2 * kk + 1][1]);
        a[kk][3] = pack_bf16(acc[2 * kk + 1][2], acc[2 * kk + 1][3]);
    }
}

template <int KIND>
__device__ __forceinline__ bool allow(int i, int j, int act_i, int act_j, int sess_i, int sess_j, int am_j, int P) {
    if (KIND == MASK_CAUSAL) return (j <= i) && am_j;
    if (KIND == MASK_MULTI_CROSS) return (j <= i) && (act_j < act_i) && am_j;
    if (KIND == MASK_SESSION) return (((j <= i) && (j / P == i / P)) || (sess_j < sess_i)) && am_j;
    return (sess_j < sess_i) && (act_j < act_i) && am_j;
}
template <int KIND>
__device__ __forceinline__ constexpr bool kind_is_causal() { return KIND == MASK_CAUSAL || KIND == MASK_MULTI_CROSS; }

// key/query metadata for one 64-token tile -> smem ints [3][64]: am, act, sess
__device__ __forceinline__ void load_meta(int* sm, const AttnArgs& a, int b, int t0) {
    for (int x = threadIdx.x; x < 3 * 64; x += blockDim.x) {
        const int which = x >> 6, r = x & 63;
        const int t = t0 + r;
        int v = 0;
        if (t < a.L) {
            const long long idx = (long long)b * a.L + t;
            v = which == 0 ? a.am[idx] : (which == 1 ? (a.act ? a.act[idx] : 0) : (a.sess ? a.sess[idx] : 0));
        }
        sm[x] = v;
    }
}

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(128) attn_fwd_kernel(AttnArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sQ = smem_u32(smem);
    const uint32_t sK0 = sQ + TILE_BYTES;           // 2 stages of K then 2 stages of V
    const uint32_t sV0 = sK0 + 2 * TILE_BYTES;
    int* sMeta = reinterpret_cast<int*>(smem + 5 * TILE_BYTES);  // [2][3][64]

    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int kvh = h / (a.n_q / a.n_kv);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i0 = qt * BQ;
    const bf16* qg = a.q + ((long long)b * a.L + i0) * a.ld + h * D;
    const bf16* kg = a.k + (long long)b * a.L * a.ld + kvh * D;
    const bf16* vg = a.v + (long long)b * a.L * a.ld + kvh * D;
    const int n_tiles_all = (a.L + BK - 1) / BK;
    const int n_tiles = kind_is_causal<KIND>() ? min(qt + 1, n_tiles_all) : n_tiles_all;

    load_tile(sQ, qg, a.ld, a.L - i0);
    load_tile(sK0, kg, a.ld, a.L);
    load_tile(sV0, vg, a.ld, a.L);
    cp_async_commit();
    load_meta(sMeta, a, b, 0);

    const int r0 = i0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
    int act_i[2] = {0, 0}, sess_i[2] = {0, 0};
    {
        const int rr[2] = {r0, r1};
#pragma unroll
        for (int x = 0; x < 2; ++x)
            if (rr[x] < a.L) {
                const long long idx = (long long)b * a.L + rr[x];
                act_i[x] = a.act ? a.act[idx] : 0;
                sess_i[x] = a.sess ? a.sess[idx] : 0;
            }
    }

    uint32_t qf[4][4];
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};

    for (int t = 0; t < n_tiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < n_tiles) {
            const int j1 = (t + 1) * BK;
            load_tile(sK0 + (buf ^ 1) * TILE_BYTES, kg + (long long)j1 * a.ld, a.ld, a.L - j1);
            load_tile(sV0 + (buf ^ 1) * TILE_BYTES, vg + (long long)j1 * a.ld, a.ld, a.L - j1);
            cp_async_commit();
            load_meta(sMeta + (buf ^ 1) * 192, a, b, j1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (t == 0) load_a_frags(sQ, warp * 16, qf);

        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
        gemm_a_bnk(s, qf, sK0 + buf * TILE_BYTES);

        const int* meta = sMeta + buf * 192;
        const int j0 = t * BK;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int jj = nb * 8 + (lane & 3) * 2 + (c & 1);
                const int j = j0 + jj;
                const int x = c >> 1;
                const int i = x ? r1 : r0;
                const bool ok = (j < a.L) && allow<KIND>(i, j, act_i[x], meta[64 + jj], sess_i[x], meta[128 + jj], meta[jj], a.P);
                const float v = ok ? s[nb][c] * a.scale_log2 : -INFINITY;
                s[nb][c] = v;
                mx[x] = fmaxf(mx[x], v);
            }
        }
#pragma unroll
        for (int x = 0; x < 2; ++x) {
            mx[x] = fmaxf(mx[x], __shfl_xor_sync(0xffffffffu, mx[x], 1));
            mx[x] = fmaxf(mx[x], __shfl_xor_sync(0xffffffffu, mx[x], 2));
        }
        float alpha[2], mnew[2];
#pragma unroll
        for (int x = 0; x < 2; ++x) {
            mnew[x] = fmaxf(mrow[x], mx[x]);
            alpha[x] = (mnew[x] == -INFINITY) ? 1.f : exp2f(mrow[x] - mnew[x]);
            mrow[x] = mnew[x];
        }
        float rs[2] = {0.f, 0.f};
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int x = c >> 1;
                const float p = (mnew[x] == -INFINITY) ? 0.f : exp2f(s[nb][c] - mnew[x]);
                s[nb][c] = p;
                rs[x] += p;
            }
        }
#pragma unroll
        for (int x = 0; x < 2; ++x) lrow[x] = lrow[x] * alpha[x] + rs[x];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            o[nb][0] *= alpha[0];
            o[nb][1] *= alpha[0];
            o[nb][2] *= alpha[1];
            o[nb][3] *= alpha[1];
        }
        uint32_t pf[4][4];
        acc_to_a_frags(s, pf);
        gemm_a_bkn(o, pf, sV0 + buf * TILE_BYTES);
        __syncthreads();
    }

    // finalize
#pragma unroll
    for (int x = 0; x < 2; ++x) {
        lrow[x] += __shfl_xor_sync(0xffffffffu, lrow[x], 1);
        lrow[x] += __shfl_xor_sync(0xffffffffu, lrow[x], 2);
    }
    const float* vm = a.vmean + ((long long)b * a.n_kv + kvh) * D;
#pragma unroll
    for (int x = 0; x < 2; ++x) {
        const int i = x ? r1 : r0;
        if (i >= a.L) continue;
        const bool uniform = !(lrow[x] > 0.f);
        const float inv = uniform ? 0.f : 1.f / lrow[x];
        bf16* op = a.o + ((long long)b * a.L + i) * a.ld_o + h * D;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            const int d = nb * 8 + (lane & 3) * 2;
            float v0 = o[nb][2 * x] * inv, v1 = o[nb][2 * x + 1] * inv;
            if (uniform) {
                v0 = vm[d];
                v1 = vm[d + 1];
            }
            *reinterpret_cast<uint32_t*>(op + d) = pack_bf16(v0, v1);
        }
        if ((lane & 3) == 0)
            a.lse[((long long)b * a.n_q + h) * a.L + i] = uniform ? INFINITY : (mrow[x] + log2f(lrow[x]));
    }
}

// vmean[b, kvh, d] = mean over all L keys of V (every key: future and padded ones included — quirk Q1)
__global__ void v_colmean_kernel(const bf16* __restrict__ v, long long ld, int L, int n_kv, float* __restrict__ vmean) {
    __shared__ float part[4][64];
    const int b = blockIdx.x / n_kv, kvh = blockIdx.x % n_kv;
    const int d = threadIdx.x & 63, g = threadIdx.x >> 6;
    const bf16* vp = v + (long long)b * L * ld + kvh * D + d;
    float acc = 0.f;
    for (int j = g; j < L; j += 4) acc += __bfloat162float(vp[(long long)j * ld]);
    part[g][d] = acc;
    __syncthreads();
    if (g == 0) vmean[(long long)blockIdx.x * D + d] = (part[0][d] + part[1][d] + part[2][d] + part[3][d]) / (float)L;
}

// ------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------
// dsum[b,h,i] = sum_d dO*O ; uni_flag[b, qt] = any uniform row in query tile qt
__global__ void attn_bwd_prep_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o, long long ld_o, int B,
                                     int L, int n_q, const float* __restrict__ lse, float* __restrict__ dsum,
                                     int* __restrict__ uni_flag, int q_tiles) {
    const long long total = (long long)B * L * n_q;
    const int sub = threadIdx.x & 7;
    const long long g0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const long long gs = ((long long)gridDim.x * blockDim.x) >> 3;
    const long long iters = (total + gs - 1) / gs;
    for (long long it = 0; it < iters; ++it) {
        const long long gi = g0 + it * gs;
        const bool live = gi < total;
        const long long row = live ? gi / n_q : 0;
        const int h = live ? (int)(gi % n_q) : 0;
        float a[8], d[8];
        bf16x8_to_float(*reinterpret_cast<const bf16x8*>(o + row * ld_o + h * D + sub * 8), a);
        bf16x8_to_float(*reinterpret_cast<const bf16x8*>(d_o + row * ld_o + h * D + sub * 8), d);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += a[i] * d[i];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (live && sub == 0) {
            const int b = (int)(row / L), i = (int)(row % L);
            const long long li = ((long long)b * n_q + h) * L + i;
            dsum[li] = s;
            if (h == 0 && lse[li] == INFINITY) uni_flag[b * q_tiles + i / BQ] = 1;
        }
    }
}

// dQ: same tiling as the forward.  dS = P o (dP - dsum), dQ = scale * dS K
template <int KIND>
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(AttnArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sQ = smem_u32(smem);
    const uint32_t sdO = sQ + TILE_BYTES;
    const uint32_t sK0 = sdO + TILE_BYTES;
    const uint32_t sV0 = sK0 + 2 * TILE_BYTES;
    int* sMeta = reinterpret_cast<int*>(smem + 6 * TILE_BYTES);

    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int kvh = h / (a.n_q / a.n_kv);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i0 = qt * BQ;
    const int q_tiles = (a.L + BQ - 1) / BQ;
    const bf16* qg = a.q + ((long long)b * a.L + i0) * a.ld + h * D;
    const bf16* dog = a.d_o + ((long long)b * a.L + i0) * a.ld_o + h * D;
    const bf16* kg = a.k + (long long)b * a.L * a.ld + kvh * D;
    const bf16* vg = a.v + (long long)b * a.L * a.ld + kvh * D;
    const int n_tiles_all = (a.L + BK - 1) / BK;
    const bool any_uniform = a.uni_flag[b * q_tiles + qt] != 0;
    const int n_tiles = (kind_is_causal<KIND>() && !any_uniform) ? min(qt + 1, n_tiles_all) : n_tiles_all;

    load_tile(sQ, qg, a.ld, a.L - i0);
    load_tile(sdO, dog, a.ld_o, a.L - i0);
    load_tile(sK0, kg, a.ld, a.L);
    load_tile(sV0, vg, a.ld, a.L);
    cp_async_commit();
    load_meta(sMeta, a, b, 0);

    const int r0 = i0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
    int act_i[2] = {0, 0}, sess_i[2] = {0, 0};
    float lse_i[2] = {0.f, 0.f}, ds_i[2] = {0.f, 0.f};
    {
        const int rr[2] = {r0, r1};
#pragma unroll
        for (int x = 0; x < 2; ++x)
            if (rr[x] < a.L) {
                const long long idx = (long long)b * a.L + rr[x];
                act_i[x] = a.act ? a.act[idx] : 0;
                sess_i[x] = a.sess ? a.sess[idx] : 0;
                const long long li = ((long long)b * a.n_q + h) * a.L + rr[x];
                lse_i[x] = a.lse[li];
                ds_i[x] = a.dsum[li];
            }
    }
    const float inv_L = 1.0f / (float)a.L;

    uint32_t qf[4][4], dof[4][4];
    float dq[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;

    for (int t = 0; t < n_tiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < n_tiles) {
            const int j1 = (t + 1) * BK;
            load_tile(sK0 + (buf ^ 1) * TILE_BYTES, kg + (long long)j1 * a.ld, a.ld, a.L - j1);
            load_tile(sV0 + (buf ^ 1) * TILE_BYTES, vg + (long long)j1 * a.ld, a.ld, a.L - j1);
            cp_async_commit();
            load_meta(sMeta + (buf ^ 1) * 192, a, b, j1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (t == 0) {
            load_a_frags(sQ, warp * 16, qf);
            load_a_frags(sdO, warp * 16, dof);
        }
        float s[8][4], dp[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
            dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
        }
        gemm_a_bnk(s, qf, sK0 + buf * TILE_BYTES);
        gemm_a_bnk(dp, dof, sV0 + buf * TILE_BYTES);
        const int* meta = sMeta + buf * 192;
        const int j0 = t * BK;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int jj = nb * 8 + (lane & 3) * 2 + (c & 1);
                const int j = j0 + jj;
                const int x = c >> 1;
                const int i = x ? r1 : r0;
                float p;
                if (lse_i[x] == INFINITY) {
                    p = (j < a.L) ? inv_L : 0.f;
                } else {
                    const bool ok = (j < a.L) && allow<KIND>(i, j, act_i[x], meta[64 + jj], sess_i[x], meta[128 + jj], meta[jj], a.P);
                    p = ok ? exp2f(s[nb][c] * a.scale_log2 - lse_i[x]) : 0.f;
                }
                s[nb][c] = p * (dp[nb][c] - ds_i[x]);
            }
        }
        uint32_t dsf[4][4];
        acc_to_a_frags(s, dsf);
        gemm_a_bkn(dq, dsf, sK0 + buf * TILE_BYTES);
        __syncthreads();
    }
#pragma unroll
    for (int x = 0; x < 2; ++x) {
        const int i = x ? r1 : r0;
        if (i >= a.L) continue;
        bf16* op = a.dq + ((long long)b * a.L + i) * a.ld_d + h * D;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            const int d = nb * 8 + (lane & 3) * 2;
            *reinterpret_cast<uint32_t*>(op + d) = pack_bf16(dq[nb][2 * x] * a.scale, dq[nb][2 * x + 1] * a.scale);
        }
    }
}

// dK/dV: one CTA per (key tile, kv head, batch); loops over the GQA group's query heads and the query tiles.
template <int KIND>
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(AttnArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sK = smem_u32(smem);
    const uint32_t sV = sK + TILE_BYTES;
    const uint32_t sQ0 = sV + TILE_BYTES;            // 2 stages Q, then 2 stages dO
    const uint32_t sdO0 = sQ0 + 2 * TILE_BYTES;
    int* sMeta = reinterpret_cast<int*>(smem + 6 * TILE_BYTES);      // [2][3][64] query metadata
    float* sStat = reinterpret_cast<float*>(sMeta + 2 * 192);        // [2][2][64] lse, dsum

    const int kt = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
    const int group = a.n_q / a.n_kv;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j0 = kt * BK;
    const int q_tiles = (a.L + BQ - 1) / BQ;
    const bf16* kg = a.k + ((long long)b * a.L + j0) * a.ld + kvh * D;
    const bf16* vg = a.v + ((long long)b * a.L + j0) * a.ld + kvh * D;

    load_tile(sK, kg, a.ld, a.L - j0);
    load_tile(sV, vg, a.ld, a.L - j0);
    cp_async_commit();

    // this thread's two key rows
    const int c0 = j0 + warp * 16 + (lane >> 2), c1 = c0 + 8;
    int act_j[2] = {0, 0}, sess_j[2] = {0, 0}, am_j[2] = {0, 0};
    {
        const int cc[2] = {c0, c1};
#pragma unroll
        for (int x = 0; x < 2; ++x)
            if (cc[x] < a.L) {
                const long long idx = (long long)b * a.L + cc[x];
                am_j[x] = a.am[idx];
                act_j[x] = a.act ? a.act[idx] : 0;
                sess_j[x] = a.sess ? a.sess[idx] : 0;
            }
    }
    const float inv_L = 1.0f / (float)a.L;

    // work list: (head in group, query tile); causal kinds skip tiles strictly above the diagonal unless they hold a
    // uniform row (which attends to every key)
    const int n_work = group * q_tiles;
    auto needed = [&](int w) {
        const int qt = w % q_tiles;
        if (!kind_is_causal<KIND>()) return true;
        return qt >= kt || a.uni_flag[b * q_tiles + qt] != 0;
    };
    auto issue = [&](int w, int buf) {
        const int hh = kvh * group + w / q_tiles, qt = w % q_tiles;
        const int i0 = qt * BQ;
        load_tile(sQ0 + buf * TILE_BYTES, a.q + ((long long)b * a.L + i0) * a.ld + hh * D, a.ld, a.L - i0);
        load_tile(sdO0 + buf * TILE_BYTES, a.d_o + ((long long)b * a.L + i0) * a.ld_o + hh * D, a.ld_o, a.L - i0);
        cp_async_commit();
        load_meta(sMeta + buf * 192, a, b, i0);
        for (int x = threadIdx.x; x < 128; x += blockDim.x) {
            const int which = x >> 6, r = x & 63;
            const int i = i0 + r;
            float v = 0.f;
            if (i < a.L) {
                const long long li = ((long long)b * a.n_q + hh) * a.L + i;
                v = which == 0 ? a.lse[li] : a.dsum[li];
            }
            sStat[buf * 128 + x] = v;
        }
    };
    int w = 0;
    while (w < n_work && !needed(w)) ++w;

    uint32_t kf[4][4], vf[4][4];
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
        dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    if (w < n_work) issue(w, 0);
    int buf = 0;
    bool first = true;
    while (w < n_work) {
        int wn = w + 1;
        while (wn < n_work && !needed(wn)) ++wn;
        if (wn < n_work) {
            issue(wn, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (first) {
            load_a_frags(sK, warp * 16, kf);
            load_a_frags(sV, warp * 16, vf);
            first = false;
        }
        const int qt = w % q_tiles;
        const int i0 = qt * BQ;
        const int* meta = sMeta + buf * 192;
        const float* stat = sStat + buf * 128;
        // S^T[key][query] and dP^T[key][query]
        float st[8][4], dpt[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
            dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
        }
        gemm_a_bnk(st, kf, sQ0 + buf * TILE_BYTES);
        gemm_a_bnk(dpt, vf, sdO0 + buf * TILE_BYTES);
        float pt[8][4];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int ii = nb * 8 + (lane & 3) * 2 + (c & 1);   // query within tile
                const int i = i0 + ii;
                const int x = c >> 1;                               // key row selector
                const int j = x ? c1 : c0;
                const float lse = stat[ii];
                float p;
                if (i >= a.L || j >= a.L) {
                    p = 0.f;
                } else if (lse == INFINITY) {
                    p = inv_L;
                } else {
                    const bool ok = allow<KIND>(i, j, meta[64 + ii], act_j[x], meta[128 + ii], sess_j[x], am_j[x], a.P);
                    p = ok ? exp2f(st[nb][c] * a.scale_log2 - lse) : 0.f;
                }
                pt[nb][c] = p;
                st[nb][c] = p * (dpt[nb][c] - stat[64 + ii]);
            }
        }
        uint32_t pf[4][4], dsf[4][4];
        acc_to_a_frags(pt, pf);
        acc_to_a_frags(st, dsf);
        gemm_a_bkn(dv, pf, sdO0 + buf * TILE_BYTES);
        gemm_a_bkn(dk, dsf, sQ0 + buf * TILE_BYTES);
        __syncthreads();
        w = wn;
        buf ^= 1;
    }
    if (first) cp_async_wait<0>();
#pragma unroll
    for (int x = 0; x < 2; ++x) {
        const int j = x ? c1 : c0;
        if (j >= a.L) continue;
        bf16* kp = a.dk + ((long long)b * a.L + j) * a.ld_d + kvh * D;
        bf16* vp = a.dv + ((long long)b * a.L + j) * a.ld_d + kvh * D;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            const int d = nb * 8 + (lane & 3) * 2;
            *reinterpret_cast<uint32_t*>(kp + d) = pack_bf16(dk[nb][2 * x] * a.scale, dk[nb][2 * x + 1] * a.scale);
            *reinterpret_cast<uint32_t*>(vp + d) = pack_bf16(dv[nb][2 * x], dv[nb][2 * x + 1]);
        }
    }
}

constexpr int FWD_SMEM = 5 * TILE_BYTES + 2 * 192 * 4;
constexpr int DQ_SMEM = 6 * TILE_BYTES + 2 * 192 * 4;
constexpr int DKV_SMEM = 6 * TILE_BYTES + 2 * 192 * 4 + 2 * 128 * 4;

template <typename K>
int set_smem(K kern, int bytes) {
    GAMER_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    return 0;
}

int check_common(int n_q, int n_kv, int head_dim, int kind, const int* act, const int* sess) {
    GAMER_REQUIRE(head_dim == D, "attention kernels are specialised for head_dim 64 (got %d)", head_dim);
    GAMER_REQUIRE(n_kv > 0 && n_q % n_kv == 0, "n_q must be a multiple of n_kv");
    GAMER_REQUIRE(kind >= 0 && kind <= 3, "unknown mask kind %d", kind);
    GAMER_REQUIRE(kind == MASK_CAUSAL || kind == MASK_SESSION || act != nullptr, "mask kind %d needs `actions`", kind);
    GAMER_REQUIRE(kind == MASK_CAUSAL || kind == MASK_MULTI_CROSS || sess != nullptr, "mask kind %d needs `session_ids`", kind);
    return 0;
}

// tcgen05 path (attention_tc.cu) unless the shape is outside its specialisation or GAMER_ATTN_LEGACY=1 (A/B testing)
bool use_tc(int L, int n_q, int n_kv, int head_dim) {
    static int legacy = -1;
    if (legacy < 0) {
        const char* e = getenv("GAMER_ATTN_LEGACY");
        legacy = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return legacy == 0 && attn_tc_supported(L, n_q, n_kv, head_dim);
}
long long vmean_bytes(int B, int n_kv) { return ((long long)B * n_kv * D * sizeof(float) + 255) / 256 * 256; }

}  // namespace

extern "C" long long gamer_attn_workspace_bytes(int B, int L, int n_q, int n_kv) {
    // vmean [B, n_kv, 64] fp32 first (callers read it back), then the tensor-core path's key codes
    return vmean_bytes(B, n_kv) + attn_tc_fwd_ws_bytes(B, L);
}

extern "C" int gamer_attn_fwd(const void* q, const void* k, const void* v, long long ld, int B, int L, int n_q, int n_kv,
                              int head_dim, int mask_kind, int tokens_per_item, const int* am, const int* act,
                              const int* sess, float scale, void* workspace, void* o, long long ld_o, float* lse,
                              const gamer_dropout_t* drop, cudaStream_t stream) {
    if (int e = check_common(n_q, n_kv, head_dim, mask_kind, act, sess)) return e;
    if (B == 0 || L == 0) return 0;
    AttnArgs a{};
    a.q = reinterpret_cast<const bf16*>(q); a.k = reinterpret_cast<const bf16*>(k); a.v = reinterpret_cast<const bf16*>(v);
    a.ld = ld; a.B = B; a.L = L; a.n_q = n_q; a.n_kv = n_kv; a.P = tokens_per_item;
    a.am = am; a.act = act; a.sess = sess; a.scale_log2 = scale * 1.4426950408889634f; a.scale = scale;
    a.vmean = reinterpret_cast<const float*>(workspace);
    a.o = reinterpret_cast<bf16*>(o); a.ld_o = ld_o; a.lse = lse;
    v_colmean_kernel<<<B * n_kv, 256, 0, stream>>>(a.v, ld, L, n_kv, reinterpret_cast<float*>(workspace));
    GAMER_LAUNCH_CHECK();
    if (use_tc(L, n_q, n_kv, head_dim))
        return attn_tc_fwd(q, k, v, ld, B, L, n_q, n_kv, mask_kind, tokens_per_item, am, act, sess, scale, a.vmean,
                           reinterpret_cast<uint8_t*>(workspace) + vmean_bytes(B, n_kv), o, ld_o, lse, drop, stream);
    GAMER_REQUIRE(drop == nullptr || !(drop->p > 0.f), "attention dropout needs the tcgen05 path (head_dim 64, GQA 2:1, L <= 4096)");
    dim3 grid((L + BQ - 1) / BQ, n_q, B);
    static bool cfg = false;
    if (!cfg) {
        if (set_smem(attn_fwd_kernel<0>, FWD_SMEM) || set_smem(attn_fwd_kernel<1>, FWD_SMEM) ||
            set_smem(attn_fwd_kernel<2>, FWD_SMEM) || set_smem(attn_fwd_kernel<3>, FWD_SMEM)) return -2;
        cfg = true;
    }
    switch (mask_kind) {
        case 0: attn_fwd_kernel<0><<<grid, 128, FWD_SMEM, stream>>>(a); break;
        case 1: attn_fwd_kernel<1><<<grid, 128, FWD_SMEM, stream>>>(a); break;
        case 2: attn_fwd_kernel<2><<<grid, 128, FWD_SMEM, stream>>>(a); break;
        default: attn_fwd_kernel<3><<<grid, 128, FWD_SMEM, stream>>>(a); break;
    }
    GAMER_LAUNCH_CHECK();
    return 0;
}

extern "C" long long gamer_attn_bwd_workspace_bytes(int B, int L, int n_q) {
    const int q_tiles = (L + BQ - 1) / BQ;
    const long long legacy = (long long)B * n_q * L * sizeof(float) + (long long)B * q_tiles * sizeof(int);
    const long long tc = attn_tc_bwd_ws_bytes(B, L, n_q);
    return legacy > tc ? legacy : tc;
}

extern "C" int gamer_attn_bwd(const void* q, const void* k, const void* v, long long ld, int B, int L, int n_q, int n_kv,
                              int head_dim, int mask_kind, int tokens_per_item, const int* am, const int* act,
                              const int* sess, float scale, const void* o, const void* d_o, long long ld_o,
                              const float* lse, void* workspace, void* dq, void* dk, void* dv, long long ld_d,
                              const gamer_dropout_t* drop, cudaStream_t stream) {
    if (int e = check_common(n_q, n_kv, head_dim, mask_kind, act, sess)) return e;
    if (B == 0 || L == 0) return 0;
    if (use_tc(L, n_q, n_kv, head_dim))
        return attn_tc_bwd(q, k, v, ld, B, L, n_q, n_kv, mask_kind, tokens_per_item, am, act, sess, scale, o, d_o, ld_o, lse,
                           workspace, dq, dk, dv, ld_d, drop, stream);
    GAMER_REQUIRE(drop == nullptr || !(drop->p > 0.f), "attention dropout needs the tcgen05 path (head_dim 64, GQA 2:1, L <= 4096)");
    const int q_tiles = (L + BQ - 1) / BQ;
    float* dsum = reinterpret_cast<float*>(workspace);
    int* uni = reinterpret_cast<int*>(dsum + (long long)B * n_q * L);
    GAMER_CHECK_CUDA(cudaMemsetAsync(uni, 0, (size_t)B * q_tiles * sizeof(int), stream));
    {
        const long long groups = (long long)B * L * n_q;
        const long long blocks = (groups * 8 + 255) / 256;
        attn_bwd_prep_kernel<<<(int)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, stream>>>(
            reinterpret_cast<const bf16*>(o), reinterpret_cast<const bf16*>(d_o), ld_o, B, L, n_q, lse, dsum, uni, q_tiles);
        GAMER_LAUNCH_CHECK();
    }
    AttnArgs a{};
    a.q = reinterpret_cast<const bf16*>(q); a.k = reinterpret_cast<const bf16*>(k); a.v = reinterpret_cast<const bf16*>(v);
    a.ld = ld; a.B = B; a.L = L; a.n_q = n_q; a.n_kv = n_kv; a.P = tokens_per_item;
    a.am = am; a.act = act; a.sess = sess; a.scale_log2 = scale * 1.4426950408889634f; a.scale = scale;
    a.ld_o = ld_o; a.lse = const_cast<float*>(lse); a.d_o = reinterpret_cast<const bf16*>(d_o);
    a.dsum = dsum; a.uni_flag = uni;
    a.dq = reinterpret_cast<bf16*>(dq); a.dk = reinterpret_cast<bf16*>(dk); a.dv = reinterpret_cast<bf16*>(dv);
    a.ld_d = ld_d;
    static bool cfg = false;
    if (!cfg) {
        if (set_smem(attn_bwd_dq_kernel<0>, DQ_SMEM) || set_smem(attn_bwd_dq_kernel<1>, DQ_SMEM) ||
            set_smem(attn_bwd_dq_kernel<2>, DQ_SMEM) || set_smem(attn_bwd_dq_kernel<3>, DQ_SMEM) ||
            set_smem(attn_bwd_dkv_kernel<0>, DKV_SMEM) || set_smem(attn_bwd_dkv_kernel<1>, DKV_SMEM) ||
            set_smem(attn_bwd_dkv_kernel<2>, DKV_SMEM) || set_smem(attn_bwd_dkv_kernel<3>, DKV_SMEM)) return -2;
        cfg = true;
    }
    dim3 gq(q_tiles, n_q, B), gk((L + BK - 1) / BK, n_kv, B);
    switch (mask_kind) {
        case 0: attn_bwd_dq_kernel<0><<<gq, 128, DQ_SMEM, stream>>>(a); attn_bwd_dkv_kernel<0><<<gk, 128, DKV_SMEM, stream>>>(a); break;
        case 1: attn_bwd_dq_kernel<1><<<gq, 128, DQ_SMEM, stream>>>(a); attn_bwd_dkv_kernel<1><<<gk, 128, DKV_SMEM, stream>>>(a); break;
        case 2: attn_bwd_dq_kernel<2><<<gq, 128, DQ_SMEM, stream>>>(a); attn_bwd_dkv_kernel<2><<<gk, 128, DKV_SMEM, stream>>>(a); break;
        default: attn_bwd_dq_kernel<3><<<gq, 128, DQ_SMEM, stream>>>(a); attn_bwd_dkv_kernel<3><<<gk, 128, DKV_SMEM, stream>>>(a); break;
    }
    GAMER_LAUNCH_CHECK();
    return 0;
}
