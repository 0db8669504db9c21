// bf16 GEMMs on the 5th-gen tensor cores (tcgen05 + TMEM), operands staged by TMA, fp32 accumulation.
//
//   gemm_tn  : C[r, n] (+)= alpha * sum_k A[r, k] * B[g(r)*N + n, k]      (forward projections and dgrad)
//              optional expert groups (row segments -> weight slab g), residual add, row scatter, fp32 out.
//   wgrad    : dW[g][i, j] += sum_r dY[r, i] * X[r, j]                     (both operands MN-major in smem)
//
// One persistent CTA per SM, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM owner),
// warps 2..5 = epilogue (TMEM -> registers -> global).  4-stage smem ring, 2 TMEM accumulator stages so the
// epilogue of tile t overlaps the MMAs of tile t+1.
#include <stdlib.h>

#include "sm100_ptx.cuh"

using namespace sm100;

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = one 128-byte swizzle atom
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int MAX_GROUPS = 8;

struct GemmParams {
    int rows;  // rows of A (upper bound in grouped mode)
    int N, K;
    int n_groups;
    const int* seg_off;  // device int32[n_groups+1], 128-aligned segment starts; nullptr => one group
    void* C;
    long long ldc;
    const bf16* resid;
    long long ldr;
    const int* row_map;  // out row for A-row r (-1: skip); nullptr => identity
    float alpha;
    DropParams drop;     // dropout of alpha * A B^T before the residual add (mask indexed by output row, column)
};

constexpr int TILE16K = 128 * 64 * 2;  // one [128 x 64] bf16 (or [128 x 32] fp32) SWIZZLE_128B box

// STAGED: the epilogue goes TMEM -> registers -> swizzled smem tile -> TMA store (fully coalesced; the residual tile is
// TMA-loaded into the same staging buffer beforehand and updated in place).  !STAGED: per-thread row stores, needed when
// output rows are scattered through `row_map`.
// BN = tile width: 128, or 256 for the wide projections (N a multiple of 256, >= 512): per 128 x 128 of output a
// 128 x 256 tile fetches 96 KB of operands instead of 128 KB, and operand fetch (TMA requests of 128-byte rows) is what
// paces these K = 256..1024 GEMMs.  The epilogue always works in 128-column halves through two staging buffers.
template <bool OUT_F32, bool STAGED, int BN>
struct GemmCfg {
    static constexpr int STAGES = STAGED ? (OUT_F32 ? 3 : (BN == 256 ? 3 : 5)) : (BN == 256 ? 4 : 6);
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_BYTES = BN * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int C_BUFS = STAGED ? 2 : 0;
    static constexpr int C_BYTES = BLOCK_M * 128 * (OUT_F32 ? 4 : 2);   // one 128-column half
    static constexpr int OFF_C = STAGES * STAGE_BYTES;
    static constexpr int OFF_BAR = OFF_C + C_BUFS * C_BYTES;
    static constexpr int TOTAL = OFF_BAR + 256 + 1024 /*align*/;
    // Epilogue warp sets.  With K = 256 a 128 x 256 tile is only 2 k clocks of MMA while four warps need ~3x that to move
    // its 64 KB through TMEM -> registers -> shared memory -> TMA store: the epilogue, not the tensor pipe, paces the wide
    // projections.  The 128 x 256 bf16 variant therefore runs TWO sets of four epilogue warps, one per 128-column half
    // (each half already has its own staging buffer).
    static constexpr int EPI_SETS = (STAGED && !OUT_F32 && BN == 256) ? 2 : 1;
    static constexpr int THREADS = 64 + 128 * EPI_SETS;
};

template <bool OUT_F32, bool STAGED, int BN>
__global__ void __launch_bounds__(GemmCfg<OUT_F32, STAGED, BN>::THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, GemmParams p) {
    using S = GemmCfg<OUT_F32, STAGED, BN>;
    constexpr int STAGES = S::STAGES;
    constexpr int TMEM_COLS = 2 * BN;
    constexpr int HALVES = BN / 128;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* c_free = tempty_bar + 2;   // [3] staging buffer's TMA store has been read out
    uint64_t* r_full = c_free + 3;       // [3] residual tile landed in the staging buffer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(r_full + 3);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const bool has_resid = STAGED && p.resid != nullptr;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        if (STAGED) {
            prefetch_tmap(&tmC);
            if (has_resid) prefetch_tmap(&tmR);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 4 * S::EPI_SETS);
        }
        for (int s = 0; s < 3; ++s) {
            mbar_init(&c_free[s], 1);
            mbar_init(&r_full[s], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // segment table -> registers
    int seg[MAX_GROUPS + 1];
    int row_end = p.rows;
    if (p.seg_off != nullptr) {
#pragma unroll
        for (int g = 0; g <= MAX_GROUPS; ++g) seg[g] = (g <= p.n_groups) ? p.seg_off[g] : 0x7fffffff;
        row_end = min(row_end, p.seg_off[p.n_groups]);
    } else {
        seg[0] = 0;
#pragma unroll
        for (int g = 1; g <= MAX_GROUPS; ++g) seg[g] = 0x7fffffff;
    }
    const int n_tiles = (p.N + BN - 1) / BN;
    const int m_tiles = (p.rows + BLOCK_M - 1) / BLOCK_M;
    const int total = n_tiles * m_tiles;
    const int k_blocks = (p.K + BLOCK_K - 1) / BLOCK_K;

    auto group_of = [&](int row0) {
        int g = 0;
#pragma unroll
        for (int i = 1; i < MAX_GROUPS; ++i) g += (row0 >= seg[i]) ? 1 : 0;
        return g;
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, tcount = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x, ++tcount) {
                const int m_blk = t / n_tiles, n_blk = t % n_tiles;
                const int row0 = m_blk * BLOCK_M;
                if (row0 >= row_end) break;
                const int g = group_of(row0);
                const int brow = g * p.N + n_blk * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * S::STAGE_BYTES;
                    mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
                    tma_load_2d(sa, &tmA, &full_bar[stage], kb * BLOCK_K, row0);
                    tma_load_2d(sa + S::A_BYTES, &tmB, &full_bar[stage], kb * BLOCK_K, brow);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (has_resid) {
                    constexpr int NB = S::C_BUFS > 0 ? S::C_BUFS : 1;
#pragma unroll
                    for (int half = 0; half < HALVES; ++half) {
                        const uint32_t hc = tcount * HALVES + half;
                        const int cbuf = hc % NB;
                        const uint32_t use = hc / NB;
                        mbar_wait(&c_free[cbuf], (use & 1) ^ 1);
                        uint8_t* sc = smem + S::OFF_C + cbuf * S::C_BYTES;
                        mbar_expect_tx(&r_full[cbuf], 2 * TILE16K);
                        tma_load_2d(sc, &tmR, &r_full[cbuf], n_blk * BN + half * 128, row0);
                        tma_load_2d(sc + TILE16K, &tmR, &r_full[cbuf], n_blk * BN + half * 128 + 64, row0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M, BN, 0, 0);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            const int row0 = (t / n_tiles) * BLOCK_M;
            if (row0 >= row_end) break;
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int kb = 0; kb < k_blocks; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
                    const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        const uint64_t ad = umma_desc_sw128(sa + k * UMMA_K * 2, 16, 1024);
                        const uint64_t bd = umma_desc_sw128(sb + k * UMMA_K * 2, 16, 1024);
                        umma_bf16(d_tmem, ad, bd, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (kb == k_blocks - 1) umma_commit(&tfull_bar[acc]);
                }
                __syncwarp();
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int quarter = warp & 3;  // TMEM lane quarter this warp may access
        const int trow = quarter * 32 + lane;
        const int eset = (warp - 2) >> 2;                 // epilogue warp set (0 unless EPI_SETS == 2)
        const bool leader = (warp == 2 + 4 * eset) && lane == 0;
        const DropParams drop = drop_resolve(p.drop);
        int acc = 0;
        uint32_t acc_phase = 0, tcount = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++tcount) {
            const int m_blk = t / n_tiles, n_blk = t % n_tiles;
            const int row0 = m_blk * BLOCK_M;
            if (row0 >= row_end) break;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN;
            if constexpr (STAGED) {
              constexpr int NB = S::C_BUFS > 0 ? S::C_BUFS : 1;
              // one set: it walks the halves; two sets: set e owns half e (and staging buffer e) of every tile
              const int half_lo = S::EPI_SETS == 2 ? eset : 0, half_hi = S::EPI_SETS == 2 ? eset + 1 : HALVES;
#pragma unroll 1
              for (int half = half_lo; half < half_hi; ++half) {
                const uint32_t hc = tcount * HALVES + half;
                const int cbuf = hc % NB;
                const uint32_t use = hc / NB;
                if (has_resid) mbar_wait(&r_full[cbuf], use & 1);
                else mbar_wait(&c_free[cbuf], (use & 1) ^ 1);
                if (half == half_lo) {
                    mbar_wait(&tfull_bar[acc], acc_phase);
                    tc_fence_after();
                }
                const uint32_t sc = smem_u32(smem + S::OFF_C + cbuf * S::C_BYTES) + trow * 128;
                const uint32_t sw = trow & 7;
                const int col_half = n_blk * BN + half * 128;
#pragma unroll
                for (int c2 = 0; c2 < 4; c2 += 2) {
                  // two 32-column chunks per TMEM round trip
                  uint32_t rr[2][32];
                  tmem_ld_32x32(taddr + half * 128 + c2 * 32, rr[0]);
                  tmem_ld_32x32(taddr + half * 128 + c2 * 32 + 32, rr[1]);
                  tmem_ld_wait();
#pragma unroll
                  for (int ci = 0; ci < 2; ++ci) {
                    const int c = c2 + ci;
                    const uint32_t* r = rr[ci];
                    if constexpr (OUT_F32) {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            sts128(sc + c * TILE16K + ((q ^ sw) << 4), __float_as_uint(__uint_as_float(r[4 * q]) * p.alpha),
                                   __float_as_uint(__uint_as_float(r[4 * q + 1]) * p.alpha),
                                   __float_as_uint(__uint_as_float(r[4 * q + 2]) * p.alpha),
                                   __float_as_uint(__uint_as_float(r[4 * q + 3]) * p.alpha));
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint32_t addr = sc + (c >> 1) * TILE16K + ((((c & 1) * 4 + q) ^ sw) << 4);
                            float v[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[8 * q + i]) * p.alpha;
                            if (drop.thresh)
                                drop_apply8(drop, (uint32_t)(row0 + trow), (uint32_t)((col_half + c * 32 + q * 8) >> 3), v);
                            if (has_resid) {
                                bf16x8 rv;
                                lds128(addr, rv.u[0], rv.u[1], rv.u[2], rv.u[3]);
                                float f[8];
                                bf16x8_to_float(rv, f);
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] += f[i];
                            }
                            const bf16x8 o = float_to_bf16x8(v);
                            sts128(addr, o.u[0], o.u[1], o.u[2], o.u[3]);
                        }
                    }
                  }
                }
                if (half == half_hi - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                }
                fence_proxy_async();
                named_bar_sync(1 + eset, 128);
                if (leader) {
                    constexpr int NBOX = OUT_F32 ? 4 : 2;
                    constexpr int BOX_COLS = OUT_F32 ? 32 : 64;
                    const uint8_t* sbuf = smem + S::OFF_C + cbuf * S::C_BYTES;
#pragma unroll
                    for (int bx = 0; bx < NBOX; ++bx)
                        if (col_half + bx * BOX_COLS < p.N)
                            tma_store_2d(&tmC, sbuf + bx * TILE16K, col_half + bx * BOX_COLS, row0);
                    bulk_commit();
                    if constexpr (S::EPI_SETS == 2) {
                        // this set owns the buffer: hand it on (to the producer's residual load or to the set's next tile)
                        // as soon as the store has read it out
                        bulk_wait_read0();
                        mbar_arrive(&c_free[cbuf]);
                    } else if (hc > 0) {  // every store but the one just issued has been read out: recycle its buffer
                        bulk_wait_read1();
                        mbar_arrive(&c_free[(hc - 1) % NB]);
                    }
                }
              }
            } else {
                const int row = row0 + trow;
                long long orow = -1;
                if (row < row_end) orow = (p.row_map != nullptr) ? (long long)p.row_map[row] : (long long)row;
                // The residual row does not depend on the accumulator: with a 128-wide tile its 256 bytes per thread are
                // fetched BEFORE the wait for the MMAs, so the scattered-row epilogue (expert down-projection) no longer
                // pays one memory round trip per 32-column chunk.
                constexpr bool PRE = (BN == 128) && !OUT_F32;
                bf16x8 pre[PRE ? 16 : 1];
                const bool pre_ok = PRE && p.resid != nullptr && orow >= 0 && (n_blk + 1) * BN <= p.N;
                if constexpr (PRE) {
                    if (pre_ok) {
                        const bf16* rp = p.resid + orow * p.ldr + n_blk * BN;
#pragma unroll
                        for (int q = 0; q < 16; ++q) pre[q] = *reinterpret_cast<const bf16x8*>(rp + 8 * q);
                    }
                }
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < BN; c += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32(taddr + c, r);
                    tmem_ld_wait();
                    const int col0 = n_blk * BN + c;
                    if (orow >= 0 && col0 < p.N) {
                        float v[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * p.alpha;
                        if (drop.thresh) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) drop_apply8(drop, (uint32_t)orow, (uint32_t)((col0 >> 3) + q), v + 8 * q);
                        }
                        const bool full = (col0 + 32 <= p.N);
                        if (p.resid != nullptr) {
                            const bf16* rp = p.resid + orow * p.ldr + col0;
                            if (PRE && pre_ok) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    float f[8];
                                    bf16x8_to_float(pre[PRE ? (c / 32) * 4 + q : 0], f);
#pragma unroll
                                    for (int i = 0; i < 8; ++i) v[8 * q + i] += f[i];
                                }
                            } else if (full) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    bf16x8 rv = *reinterpret_cast<const bf16x8*>(rp + 8 * q);
                                    float f[8];
                                    bf16x8_to_float(rv, f);
#pragma unroll
                                    for (int i = 0; i < 8; ++i) v[8 * q + i] += f[i];
                                }
                            } else {
                                for (int i = 0; i < 32 && col0 + i < p.N; ++i) v[i] += __bfloat162float(rp[i]);
                            }
                        }
                        if constexpr (OUT_F32) {
                            float* cp = reinterpret_cast<float*>(p.C) + orow * p.ldc + col0;
                            if (full) {
#pragma unroll
                                for (int q = 0; q < 8; ++q)
                                    *reinterpret_cast<float4*>(cp + 4 * q) =
                                        make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                            } else {
                                for (int i = 0; i < 32 && col0 + i < p.N; ++i) cp[i] = v[i];
                            }
                        } else {
                            bf16* cp = reinterpret_cast<bf16*>(p.C) + orow * p.ldc + col0;
                            if (full) {
#pragma unroll
                                for (int q = 0; q < 4; ++q)
                                    *reinterpret_cast<bf16x8*>(cp + 8 * q) = float_to_bf16x8(v + 8 * q);
                            } else {
                                for (int i = 0; i < 32 && col0 + i < p.N; ++i) cp[i] = __float2bfloat16(v[i]);
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            }
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
        if (STAGED && leader) bulk_wait0();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------
// wgrad: dW[g][i, j] += sum_r dY[r, i] * X[r, j].  The reduction runs over token rows, which is the slow dimension
// of both operands, so they are staged as MN-major SWIZZLE_128B tiles: TMA boxes of [64 rows x 64 cols] (128-byte
// rows), UMMA descriptors with LBO = 8192 (next 64-wide MN atom = next box) and SBO = 1024 (next 8 k-rows).
// Work item = (group, row chunk, i-block, j-block); partial sums are reduced with fp32 red.global.
// ---------------------------------------------------------------------------------------------------------------
constexpr int W_STAGES = 4;
constexpr int W_BOX_BYTES = 64 * 64 * 2;           // 8 KB
constexpr int W_A_BYTES = 2 * W_BOX_BYTES;         // 128 output rows (i)
constexpr int W_B_BYTES = 4 * W_BOX_BYTES;         // up to 256 output cols (j)
constexpr int W_STAGE_BYTES = W_A_BYTES + W_B_BYTES;
constexpr int W_OFF_STG = W_STAGES * W_STAGE_BYTES;  // 2 x [128 rows x 32 fp32] staging boxes for the reduce-add epilogue
constexpr int W_OFF_BAR = W_OFF_STG + 2 * TILE16K;
constexpr int W_SMEM_TOTAL = W_OFF_BAR + 256 + 1024;

struct WgradParams {
    int rows, N_out, K_in, n_groups;
    const int* seg_off;
    float* dW;
    int bj;        // j-block width (multiple of 64, <= 256)
    int n_i, n_j;  // output tile grid
    int chunk_rows;  // token rows per work item (multiple of 64): sized on the host so one wave of CTAs covers the work
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX,
             const __grid_constant__ CUtensorMap tmW, WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + W_OFF_BAR);
    uint64_t* empty_bar = full_bar + W_STAGES;
    uint64_t* tfull_bar = empty_bar + W_STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmY);
        prefetch_tmap(&tmX);
        for (int s = 0; s < W_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // chunk table: group g owns chunks [cstart[g], cstart[g+1])
    int gbeg[MAX_GROUPS], gend[MAX_GROUPS], cstart[MAX_GROUPS + 1];
    cstart[0] = 0;
#pragma unroll
    for (int g = 0; g < MAX_GROUPS; ++g) {
        int b = 0, e = 0;
        if (g < p.n_groups) {
            if (p.seg_off != nullptr) {
                b = p.seg_off[g];
                e = min(p.seg_off[g + 1], p.rows);
            } else {
                b = 0;
                e = p.rows;
            }
        }
        gbeg[g] = b;
        gend[g] = e;
        cstart[g + 1] = cstart[g] + (max(e - b, 0) + p.chunk_rows - 1) / p.chunk_rows;
    }
    const int n_ij = p.n_i * p.n_j;
    const int total = cstart[MAX_GROUPS] * n_ij;
    const int n_jbox = p.bj / 64;
    const uint32_t stage_tx = (uint32_t)(2 + n_jbox) * W_BOX_BYTES;

    auto decode = [&](int w, int& g, int& r0, int& r1, int& ib, int& jb) {
        const int chunk = w / n_ij, ij = w % n_ij;
        ib = ij / p.n_j;
        jb = ij % p.n_j;
        g = 0;
#pragma unroll
        for (int i = 1; i < MAX_GROUPS; ++i) g += (chunk >= cstart[i]) ? 1 : 0;
        r0 = 0;
        r1 = 0;
#pragma unroll
        for (int i = 0; i < MAX_GROUPS; ++i)
            if (i == g) {
                r0 = gbeg[i] + (chunk - cstart[i]) * p.chunk_rows;
                r1 = min(r0 + p.chunk_rows, gend[i]);
            }
    };

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = blockIdx.x; w < total; w += gridDim.x) {
                int g, r0, r1, ib, jb;
                decode(w, g, r0, r1, ib, jb);
                for (int r = r0; r < r1; r += 64) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * W_STAGE_BYTES;
                    mbar_expect_tx(&full_bar[stage], stage_tx);
                    tma_load_2d(sa, &tmY, &full_bar[stage], ib * 128, r);
                    tma_load_2d(sa + W_BOX_BYTES, &tmY, &full_bar[stage], ib * 128 + 64, r);
                    for (int c = 0; c < n_jbox; ++c)
                        tma_load_2d(sa + W_A_BYTES + c * W_BOX_BYTES, &tmX, &full_bar[stage], jb * p.bj + c * 64, r);
                    if (++stage == W_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = umma_idesc_bf16(128, p.bj, 1, 1);
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (int w = blockIdx.x; w < total; w += gridDim.x) {
            int g, r0, r1, ib, jb;
            decode(w, g, r0, r1, ib, jb);
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 256;
            for (int r = r0; r < r1; r += 64) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = smem_u32(smem + stage * W_STAGE_BYTES);
                    const uint32_t sb = sa + W_A_BYTES;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {  // 4 x 16 token rows per 64-row stage
                        const uint64_t ad = umma_desc_sw128(sa + k * 16 * 128, W_BOX_BYTES, 1024);
                        const uint64_t bd = umma_desc_sw128(sb + k * 16 * 128, W_BOX_BYTES, 1024);
                        umma_bf16(d_tmem, ad, bd, idesc, (r != r0 || k != 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (r + 64 >= r1) umma_commit(&tfull_bar[acc]);
                }
                __syncwarp();
                if (++stage == W_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    } else {
        const int quarter = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0, box_n = 0;
        for (int w = blockIdx.x; w < total; w += gridDim.x) {
            int g, r0, r1, ib, jb;
            decode(w, g, r0, r1, ib, jb);
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            // TMEM -> registers -> swizzled [128 x 32] fp32 staging box -> TMA reduce-add into dW (rows / columns past the
            // matrix edge are clipped by the tensor map).  Two staging boxes alternate.
            const int trow = quarter * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * 256;
            for (int c = 0; c < p.bj; c += 32, ++box_n) {
                uint32_t rr[32];
                tmem_ld_32x32(taddr + c, rr);
                tmem_ld_wait();
                uint8_t* stg = smem + W_OFF_STG + (box_n & 1) * TILE16K;
                if (warp == 2 && lane == 0 && box_n >= 2) bulk_wait_read1();  // the box issued two steps ago has been read
                named_bar_sync(1, 128);
                const uint32_t srow = smem_u32(stg) + trow * 128;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    sts128(srow + ((q ^ (trow & 7)) << 4), rr[4 * q], rr[4 * q + 1], rr[4 * q + 2], rr[4 * q + 3]);
                fence_proxy_async();
                named_bar_sync(1, 128);
                if (warp == 2 && lane == 0) {
                    if (jb * p.bj + c < p.K_in) tma_reduce_add_3d(&tmW, stg, jb * p.bj + c, ib * 128, g);
                    bulk_commit();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }
    if (warp == 2 && lane == 0) bulk_wait0();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------
// plain CUDA-core reference GEMMs: GPU-side checkers for the tests (never on the product path)
// ---------------------------------------------------------------------------------------------------------------
__global__ void ref_gemm_tn_kernel(const bf16* A, long long lda, const bf16* B, long long ldb, float* C, long long ldc,
                                   int rows, int N, int K) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (n >= N || r >= rows) return;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc += __bfloat162float(A[r * lda + k]) * __bfloat162float(B[n * ldb + k]);
    C[r * ldc + n] = acc;
}

// ---------------------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
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

// 2-D bf16 tensor [rows, cols] with row stride ld (elements); box = [box_rows, 64 cols], SWIZZLE_128B.
int make_tmap_bf16(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    GAMER_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
    GAMER_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
    GAMER_REQUIRE((ld * 2) % 16 == 0, "TMA row stride must be a multiple of 16 bytes (ld=%lld)", ld);
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GAMER_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (rows=%lld cols=%lld ld=%lld)", (int)r, rows,
                  cols, ld);
    return 0;
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    }
    return n;
}

// 2-D fp32 tensor [rows, cols] with row stride ld (elements); box = [128 rows, 32 cols] (128 bytes), SWIZZLE_128B.
int make_tmap_f32(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld) {
    EncodeTiledFn fn = get_encode_fn();
    GAMER_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
    GAMER_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 4) % 16 == 0,
                  "fp32 TMA output must be 16-byte aligned (ld=%lld)", ld);
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GAMER_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (fp32) failed with %d (rows=%lld cols=%lld ld=%lld)", (int)r, rows,
                  cols, ld);
    return 0;
}

template <bool OUT_F32, bool STAGED, int BN = 128>
int launch_gemm_tn(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR,
                   const GemmParams& p, cudaStream_t stream) {
    using S = GemmCfg<OUT_F32, STAGED, BN>;
    auto kern = gemm_tn_kernel<OUT_F32, STAGED, BN>;
    static PerDeviceOnce configured;
    if (configured.need())
        GAMER_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    const int tiles = ceil_div(p.rows, BLOCK_M) * ceil_div(p.N, BN);
    const int grid = tiles < num_sms() ? tiles : num_sms();
    kern<<<grid, S::THREADS, S::TOTAL, stream>>>(tmA, tmB, tmC, tmR, p);
    GAMER_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" int gamer_gemm_bf16_tn(const void* A, long long lda, int rows, const void* B, long long ldb, int n_groups,
                                  int N, int K, const int* seg_off, void* C, long long ldc, int c_is_f32,
                                  const void* resid, long long ldr, const int* row_map, float alpha,
                                  const gamer_dropout_t* drop, cudaStream_t stream) {
    if (rows <= 0) return 0;
    GAMER_REQUIRE(n_groups >= 1 && n_groups <= MAX_GROUPS, "n_groups=%d out of range", n_groups);
    GAMER_REQUIRE(n_groups == 1 || seg_off != nullptr, "grouped GEMM needs seg_off");
    GAMER_REQUIRE(ldc % (c_is_f32 ? 4 : 8) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0,
                  "C rows must be 16-byte aligned (ldc=%lld)", ldc);
    GAMER_REQUIRE(resid == nullptr || (ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(resid) & 15) == 0),
                  "residual rows must be 16-byte aligned (ldr=%lld)", ldr);
    CUtensorMap tmA, tmB, tmC, tmR;
    if (int e = make_tmap_bf16(&tmA, A, rows, K, lda, BLOCK_M)) return e;
    // staged (TMA-store) epilogue whenever output rows are the A rows; fp32 outputs carry no residual in this code base
    const bool staged = row_map == nullptr && !(c_is_f32 && resid != nullptr);
    // 128 x 256 tiles for the wide projections, and for the N = 256 dgrads with a long reduction and no residual tile
    // (measured per shape with tools/gemm_bench.py; the residual GEMMs of width 256 are faster with 128 x 128 tiles)
    const bool wide = staged && !c_is_f32 && N % 256 == 0 && (N >= 512 || (resid == nullptr && K >= 768));
    if (int e = make_tmap_bf16(&tmB, B, (long long)n_groups * N, K, ldb, wide ? 256 : 128)) return e;
    GemmParams p{rows, N, K, n_groups, seg_off, C, ldc, reinterpret_cast<const bf16*>(resid), ldr, row_map, alpha,
                 make_drop(drop, 16)};
    GAMER_REQUIRE(p.drop.thresh == 0 || !c_is_f32, "dropout epilogue is implemented for bf16 outputs");
    if (!staged) {
        tmC = tmA;
        tmR = tmA;
        return c_is_f32 ? launch_gemm_tn<true, false>(tmA, tmB, tmC, tmR, p, stream)
                        : launch_gemm_tn<false, false>(tmA, tmB, tmC, tmR, p, stream);
    }
    if (c_is_f32) {
        if (int e = make_tmap_f32(&tmC, C, rows, N, ldc)) return e;
        tmR = tmA;
        return launch_gemm_tn<true, true>(tmA, tmB, tmC, tmR, p, stream);
    }
    if (int e = make_tmap_bf16(&tmC, C, rows, N, ldc, BLOCK_M)) return e;
    tmR = tmA;
    if (resid != nullptr)
        if (int e = make_tmap_bf16(&tmR, resid, rows, N, ldr, BLOCK_M)) return e;
    if (wide) return launch_gemm_tn<false, true, 256>(tmA, tmB, tmC, tmR, p, stream);
    return launch_gemm_tn<false, true>(tmA, tmB, tmC, tmR, p, stream);
}

extern "C" int gamer_gemm_bf16_wgrad(const void* dY, long long ldy, const void* X, long long ldx, int rows, int N_out,
                                     int K_in, int n_groups, const int* seg_off, float* dW, cudaStream_t stream) {
    if (rows <= 0) return 0;
    GAMER_REQUIRE(n_groups >= 1 && n_groups <= MAX_GROUPS, "n_groups=%d out of range", n_groups);
    GAMER_REQUIRE(n_groups == 1 || seg_off != nullptr, "grouped wgrad needs seg_off");
    CUtensorMap tmY, tmX;
    if (int e = make_tmap_bf16(&tmY, dY, rows, N_out, ldy, 64)) return e;
    if (int e = make_tmap_bf16(&tmX, X, rows, K_in, ldx, 64)) return e;
    WgradParams p;
    p.rows = rows;
    p.N_out = N_out;
    p.K_in = K_in;
    p.n_groups = n_groups;
    p.seg_off = seg_off;
    p.dW = dW;
    p.n_i = ceil_div(N_out, 128);
    p.n_j = ceil_div(K_in, 256);
    p.bj = ceil_div(ceil_div(K_in, p.n_j), 64) * 64;
    // dW as [n_groups][N_out][K_in] fp32: box = [1][128 rows][32 cols]
    CUtensorMap tmW;
    {
        EncodeTiledFn fn = get_encode_fn();
        GAMER_REQUIRE((reinterpret_cast<uintptr_t>(dW) & 15) == 0 && (K_in * 4) % 16 == 0,
                      "wgrad output rows must be 16-byte aligned (K_in=%d)", K_in);
        cuuint64_t dims[3] = {(cuuint64_t)K_in, (cuuint64_t)N_out, (cuuint64_t)n_groups};
        cuuint64_t strides[2] = {(cuuint64_t)K_in * 4, (cuuint64_t)N_out * K_in * 4};
        cuuint32_t box[3] = {32, 128, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = fn(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dW, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        GAMER_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (dW) failed with %d (N_out=%d K_in=%d groups=%d)", (int)r, N_out,
                      K_in, n_groups);
    }
    static PerDeviceOnce configured;
    if (configured.need())
        GAMER_CHECK_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, W_SMEM_TOTAL));
    // chunk the token rows so that (groups x tiles x chunks) fills one wave of CTAs as evenly as possible
    const int tiles = n_groups * p.n_i * p.n_j;
    const int chunks_per_group = tiles >= num_sms() ? 1 : num_sms() / tiles;
    const int rows_per_group = ceil_div(rows, n_groups);
    p.chunk_rows = ceil_div(ceil_div(rows_per_group, chunks_per_group), 64) * 64;
    if (p.chunk_rows < 256) p.chunk_rows = 256;
    // upper bound on work items (grouped: every group may add one partial chunk)
    const long long items = (long long)(ceil_div(rows, p.chunk_rows) + n_groups) * p.n_i * p.n_j;
    const int grid = (int)(items < num_sms() ? items : num_sms());
    wgrad_kernel<<<grid, NUM_THREADS, W_SMEM_TOTAL, stream>>>(tmY, tmX, tmW, p);
    GAMER_LAUNCH_CHECK();
    return 0;
}

// GPU-side reference (tests only): C fp32 = A * B^T on CUDA cores
extern "C" int gamer_ref_gemm_tn(const void* A, long long lda, const void* B, long long ldb, float* C, long long ldc,
                                 int rows, int N, int K, cudaStream_t stream) {
    if (rows <= 0) return 0;
    dim3 grid(ceil_div(N, 128), rows);
    ref_gemm_tn_kernel<<<grid, 128, 0, stream>>>(reinterpret_cast<const bf16*>(A), lda, reinterpret_cast<const bf16*>(B),
                                                 ldb, C, ldc, rows, N, K);
    GAMER_LAUNCH_CHECK();
    return 0;
}
