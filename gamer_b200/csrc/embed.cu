// K1: fused semantic-ID embedding gather + router indices (forward), expert-routing permutation, and the sparse
// embedding gradient (backward) as a sort-once / segmented-reduce scatter-add.
//
// Reference semantics: Qwen3MultiDecoderRouter.forward (SeqRec/models/generative/Qwen3Multi/router.py:74-201) and
// nn.Embedding(padding_idx=4) (Qwen3Multi/model.py:263,779).
#include "common.cuh"

namespace {

constexpr int P_TOK = 5;       // tokens per item is a runtime value; 5 in every shipped config
constexpr int MAX_EXPERTS = 8;

// ------------------------------------------------------------------------------------------------------------
// forward: one warp per token.  x[m,:] = table[id]; position / behaviour / action indices.
// ------------------------------------------------------------------------------------------------------------
struct RouteArgs {
    const long long* ids;  // [B, S] tokens processed by this call
    const long long* ctx;  // [B, ctx_ld] whole sequence so far (== ids for a full forward)
    long long ctx_ld;
    int B, S, pos0;        // token s sits at absolute position pos0 + s
    int P;                 // tokens per item
    int pad, eos, vocab;
    const int* beh_lut;    // [vocab] mapped behaviour index (+1), raw id when unmapped (router.py:122-123)
    int n_beh;             // clamp for the embedding lookups
    int* err;              // nullable: bit 0 is set when a token id lies outside [0, vocab)
};

__device__ __forceinline__ void route_token(const RouteArgs& a, int b, int s, long long id, int& pos_i, int& beh_i,
                                            int& act_i) {
    const int t = a.pos0 + s;
    const bool special = (id == a.pad) || (id == a.eos);
    const int n_items = (a.pos0 + a.S - 1 + a.P - 1) / a.P;  // (max position + P - 1) // P
    const int item0 = (t / a.P) * a.P;
    int mapped = 0;
    if (item0 < n_items * a.P) {
        long long bt = a.ctx[(long long)b * a.ctx_ld + item0];
        mapped = (bt >= 0 && bt < a.vocab) ? a.beh_lut[bt] : (int)bt;
    }
    pos_i = special ? 0 : (t % a.P) + 1;
    act_i = special ? 0 : mapped;
    beh_i = (special || (t % a.P) == 0) ? 0 : mapped;
}

// A warp takes TOK consecutive tokens per iteration: lanes 0..TOK-1 load one id each (one coalesced request), run the
// router for their token and write the three index arrays; the ids are then broadcast by shuffle and the TOK table rows
// are copied with all their 16-byte loads in flight before the first store (the table is L2-resident: 533 KB; the
// kernel is bound by the 512 B/token row write).
constexpr int TOK = 8;

__global__ void __launch_bounds__(256)
embed_route_kernel(RouteArgs a, const bf16* __restrict__ table, int H, bf16* __restrict__ x,
                   int* __restrict__ pos_idx, int* __restrict__ beh_idx, int* __restrict__ act_idx) {
    const int warps_per_block = blockDim.x >> 5;
    const long long M = (long long)a.B * a.S;
    const int lane = threadIdx.x & 31;
    const long long n_groups = (M + TOK - 1) / TOK;
    for (long long grp = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); grp < n_groups;
         grp += (long long)gridDim.x * warps_per_block) {
        const long long m0 = grp * TOK;
        long long id = a.pad;
        if (lane < TOK && m0 + lane < M) {
            const long long m = m0 + lane;
            id = a.ids[m];
            if ((id < 0 || id >= a.vocab) && a.err != nullptr) atomicOr(a.err, 1);   // nn.Embedding raises here
            int p, be, ac;
            route_token(a, (int)(m / a.S), (int)(m % a.S), id, p, be, ac);
            pos_idx[m] = p;
            beh_idx[m] = min(max(be, 0), a.n_beh);
            act_idx[m] = min(max(ac, 0), a.n_beh);
        }
        if (x != nullptr) {
            long long rows[TOK];
#pragma unroll
            for (int u = 0; u < TOK; ++u) {
                const long long idu = __shfl_sync(0xffffffffu, id, u);
                rows[u] = (idu >= 0 && idu < a.vocab) ? idu : a.pad;
            }
            for (int c = lane; c < H / 8; c += 32) {  // H = 256: exactly one pass, 16 B per lane
                bf16x8 v[TOK];
#pragma unroll
                for (int u = 0; u < TOK; ++u) v[u] = *reinterpret_cast<const bf16x8*>(table + rows[u] * H + c * 8);
#pragma unroll
                for (int u = 0; u < TOK; ++u)
                    if (m0 + u < M) *reinterpret_cast<bf16x8*>(x + (m0 + u) * H + c * 8) = v[u];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// expert routing permutation (expert = position index).  Segment starts are 128-row aligned so a GEMM tile never
// straddles two experts; padding rows map to -1.
// ------------------------------------------------------------------------------------------------------------
__global__ void route_count_kernel(const int* __restrict__ pos_idx, int S, int n_exp, int* __restrict__ counts) {
    __shared__ int sc[MAX_EXPERTS];
    if (threadIdx.x < MAX_EXPERTS) sc[threadIdx.x] = 0;
    __syncthreads();
    const int b = blockIdx.x;
    int local[MAX_EXPERTS] = {0};
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const int e = pos_idx[(long long)b * S + s];
#pragma unroll
        for (int k = 0; k < MAX_EXPERTS; ++k) local[k] += (e == k);
    }
#pragma unroll
    for (int k = 0; k < MAX_EXPERTS; ++k) {
        int v = local[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sc[k], v);
    }
    __syncthreads();
    if (threadIdx.x < n_exp) counts[b * MAX_EXPERTS + threadIdx.x] = sc[threadIdx.x];
}

// single block: per-expert exclusive scan over sequences + aligned segment offsets
__global__ void route_scan_kernel(const int* __restrict__ counts, int B, int n_exp, int* __restrict__ base,
                                  int* __restrict__ seg_off) {
    __shared__ int totals[MAX_EXPERTS];
    __shared__ int warp_part[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int e = 0; e < n_exp; ++e) {
        int running = 0;
        for (int b0 = 0; b0 < B; b0 += blockDim.x) {
            const int b = b0 + threadIdx.x;
            const int v = (b < B) ? counts[b * MAX_EXPERTS + e] : 0;
            int incl = v;
            for (int o = 1; o < 32; o <<= 1) {
                int n = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += n;
            }
            if (lane == 31) warp_part[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                int w = (lane < (blockDim.x >> 5)) ? warp_part[lane] : 0;
                int wi = w;
                for (int o = 1; o < 32; o <<= 1) {
                    int n = __shfl_up_sync(0xffffffffu, wi, o);
                    if (lane >= o) wi += n;
                }
                warp_part[lane] = wi - w;  // exclusive
                if (lane == 31) totals[e] = wi;
            }
            __syncthreads();
            if (b < B) base[b * MAX_EXPERTS + e] = running + warp_part[warp] + incl - v;
            running += totals[e];
            __syncthreads();
        }
        if (threadIdx.x == 0) totals[e] = running;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int off = 0;
        for (int e = 0; e < n_exp; ++e) {
            seg_off[e] = off;
            off += (totals[e] + 127) / 128 * 128;
        }
        seg_off[n_exp] = off;
    }
}

__global__ void route_fill_kernel(const int* __restrict__ pos_idx, int B, int S, int n_exp, const int* __restrict__ base,
                                  const int* __restrict__ seg_off, int* __restrict__ perm, int* __restrict__ rows) {
    // one warp per sequence; in-order ranks via ballot
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    int cursor[MAX_EXPERTS];
#pragma unroll
    for (int e = 0; e < MAX_EXPERTS; ++e) cursor[e] = (e < n_exp) ? seg_off[e] + base[b * MAX_EXPERTS + e] : 0;
    for (int s0 = 0; s0 < S; s0 += 32) {
        const int s = s0 + lane;
        const int e = (s < S) ? pos_idx[(long long)b * S + s] : -1;
        int dst = -1;
#pragma unroll
        for (int k = 0; k < MAX_EXPERTS; ++k) {
            const unsigned mask = __ballot_sync(0xffffffffu, e == k);
            if (e == k) dst = cursor[k] + __popc(mask & ((1u << lane) - 1));
            cursor[k] += __popc(mask);
        }
        if (s < S) {
            perm[(long long)b * S + s] = dst;
            rows[dst] = (int)((long long)b * S + s);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// backward: counting sort of token rows by vocabulary id (done once per batch, at forward time), then a segmented
// reduction: one warp sums <= CHUNK gradient rows of one id in registers and issues one fp32 atomic per column.
// Row `pad` (padding_idx) receives no embedding-path gradient (Q10).
// ------------------------------------------------------------------------------------------------------------
constexpr int EMB_CHUNK = 64;

__global__ void emb_hist_kernel(const long long* __restrict__ ids, long long M, int vocab, int* __restrict__ hist) {
    extern __shared__ int sh[];
    for (int v = threadIdx.x; v < vocab; v += blockDim.x) sh[v] = 0;
    __syncthreads();
    for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
        const long long id = ids[m];
        const bool ok = id >= 0 && id < vocab;
        // warp-aggregated: lanes with the same id elect one leader that adds the group's population
        const unsigned act = __ballot_sync(__activemask(), ok);
        if (ok) {
            const unsigned peers = __match_any_sync(act, (int)id);
            if ((__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&sh[id], __popc(peers));
        }
    }
    __syncthreads();
    for (int v = threadIdx.x; v < vocab; v += blockDim.x)
        if (sh[v]) atomicAdd(&hist[v], sh[v]);
}

// single block: bin_start[v] (exclusive scan of hist) and chunk_start[v] (exclusive scan of ceil(hist/CHUNK), pad bin
// contributes no chunks).  Also resets cursor[] for the scatter pass.
__global__ void emb_scan_kernel(const int* __restrict__ hist, int vocab, int pad, int* __restrict__ bin_start,
                                int* __restrict__ chunk_start, int* __restrict__ cursor) {
    __shared__ int carry[2];
    __shared__ int wsum[2][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry[0] = carry[1] = 0;
    __syncthreads();
    for (int v0 = 0; v0 < vocab; v0 += blockDim.x) {
        const int v = v0 + threadIdx.x;
        const int h = (v < vocab) ? hist[v] : 0;
        const int c = (v < vocab && v != pad) ? (h + EMB_CHUNK - 1) / EMB_CHUNK : 0;
        int ih = h, ic = c;
        for (int o = 1; o < 32; o <<= 1) {
            int a = __shfl_up_sync(0xffffffffu, ih, o), b = __shfl_up_sync(0xffffffffu, ic, o);
            if (lane >= o) {
                ih += a;
                ic += b;
            }
        }
        if (lane == 31) {
            wsum[0][warp] = ih;
            wsum[1][warp] = ic;
        }
        __syncthreads();
        if (warp == 0) {
            int a = (lane < (blockDim.x >> 5)) ? wsum[0][lane] : 0, b = (lane < (blockDim.x >> 5)) ? wsum[1][lane] : 0;
            int ia = a, ib = b;
            for (int o = 1; o < 32; o <<= 1) {
                int x = __shfl_up_sync(0xffffffffu, ia, o), y = __shfl_up_sync(0xffffffffu, ib, o);
                if (lane >= o) {
                    ia += x;
                    ib += y;
                }
            }
            wsum[0][lane] = ia - a;
            wsum[1][lane] = ib - b;
        }
        __syncthreads();
        if (v < vocab) {
            bin_start[v] = carry[0] + wsum[0][warp] + ih - h;
            chunk_start[v] = carry[1] + wsum[1][warp] + ic - c;
            cursor[v] = 0;
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) {
            carry[0] += wsum[0][warp] + ih;
            carry[1] += wsum[1][warp] + ic;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        bin_start[vocab] = carry[0];
        chunk_start[vocab] = carry[1];
    }
}

__global__ void emb_scatter_kernel(const long long* __restrict__ ids, long long M, int vocab,
                                   const int* __restrict__ bin_start, int* __restrict__ cursor,
                                   int* __restrict__ sorted_rows) {
    for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
        const long long id = ids[m];
        const bool ok = id >= 0 && id < vocab;
        const unsigned act = __ballot_sync(__activemask(), ok);
        if (ok) {
            const unsigned peers = __match_any_sync(act, (int)id);
            const int leader = __ffs(peers) - 1;
            int basep = 0;
            if ((int)(threadIdx.x & 31) == leader) basep = atomicAdd(&cursor[id], __popc(peers));
            basep = __shfl_sync(peers, basep, leader);
            const int rank = __popc(peers & ((1u << (threadIdx.x & 31)) - 1));
            sorted_rows[bin_start[id] + basep + rank] = (int)m;
        }
    }
}

__global__ void emb_reduce_kernel(const bf16* __restrict__ dx, int H, const int* __restrict__ bin_start,
                                  const int* __restrict__ chunk_start, const int* __restrict__ sorted_rows, int vocab,
                                  float* __restrict__ dtable) {
    const int lane = threadIdx.x & 31;
    const int n_chunks = chunk_start[vocab];
    for (int ch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); ch < n_chunks;
         ch += gridDim.x * (blockDim.x >> 5)) {
        // largest v with chunk_start[v] <= ch
        int lo = 0, hi = vocab;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (chunk_start[mid] <= ch) lo = mid; else hi = mid;
        }
        const int v = lo;
        const int beg = bin_start[v] + (ch - chunk_start[v]) * EMB_CHUNK;
        const int end = min(beg + EMB_CHUNK, bin_start[v + 1]);
        // the chunk's (<= 64) row indices: two coalesced loads, broadcast by shuffle below
        const int n = end - beg;
        const int idx0 = (lane < n) ? sorted_rows[beg + lane] : 0;
        const int idx1 = (32 + lane < n) ? sorted_rows[beg + 32 + lane] : 0;
        for (int c0 = 0; c0 < H / 8; c0 += 32) {  // H = 256: exactly one pass, 16 B per lane
            const int c = c0 + lane;
            const bool col_ok = c < H / 8;
            float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const int cc = col_ok ? c : 0;
            for (int i = 0; i < n; i += 8) {      // 8 rows (4 KB per warp) in flight
                // every load is unconditional (past the end of the chunk it re-reads the chunk's last row, a cache hit, and
                // the sum skips it): predicated loads made the compiler reuse destination registers, which serialised them
                bf16x8 r[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int k = min(i + u, n - 1);
                    const int row = __shfl_sync(0xffffffffu, k < 32 ? idx0 : idx1, k & 31);
                    const uint4 t = __ldcs(reinterpret_cast<const uint4*>(dx + (long long)row * H + cc * 8));
                    r[u].u[0] = t.x; r[u].u[1] = t.y; r[u].u[2] = t.z; r[u].u[3] = t.w;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    float f[8];
                    bf16x8_to_float(r[u], f);
                    const float keep = (i + u < n) ? 1.f : 0.f;
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[k] = fmaf(f[k], keep, acc[k]);
                }
            }
            // behaviour tokens are a fifth of all rows, so thousands of chunks land on the same table row at the same
            // time: two 16-byte vector reductions per lane instead of eight scalar ones (same-address serialisation in L2
            // was what held this kernel at 30 % of HBM peak)
            float* out = dtable + (long long)v * H + c * 8;
            if (col_ok) {
                if ((reinterpret_cast<unsigned long long>(out) & 15ull) == 0) {
                    atomicAdd(reinterpret_cast<float4*>(out), make_float4(acc[0], acc[1], acc[2], acc[3]));
                    atomicAdd(reinterpret_cast<float4*>(out) + 1, make_float4(acc[4], acc[5], acc[6], acc[7]));
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) atomicAdd(out + k, acc[k]);
                }
            }
        }
    }
}

}  // namespace

extern "C" int gamer_embed_route_fwd(const long long* ids, const long long* ctx, long long ctx_ld, int B, int S, int pos0,
                                     int tokens_per_item, int pad, int eos, int vocab, const int* beh_lut, int n_beh,
                                     const void* table_bf16, int H, void* x_bf16, int* pos_idx, int* beh_idx,
                                     int* act_idx, int* err, cudaStream_t stream) {
    GAMER_REQUIRE(H % 8 == 0, "hidden size must be a multiple of 8");
    GAMER_REQUIRE(tokens_per_item >= 1, "tokens_per_item must be >= 1");
    const long long M = (long long)B * S;
    if (M == 0) return 0;
    RouteArgs a{ids, ctx ? ctx : ids, ctx ? ctx_ld : (long long)S, B, S, pos0, tokens_per_item, pad, eos, vocab,
                beh_lut, n_beh, err};
    const int threads = 256, wpb = threads / 32;
    const long long groups = (M + TOK - 1) / TOK;
    const int grid = (int)((groups + wpb - 1) / wpb < 148 * 8 ? (groups + wpb - 1) / wpb : 148 * 8);
    embed_route_kernel<<<grid, threads, 0, stream>>>(a, reinterpret_cast<const bf16*>(table_bf16), H,
                                                     reinterpret_cast<bf16*>(x_bf16), pos_idx, beh_idx, act_idx);
    GAMER_LAUNCH_CHECK();
    (void)P_TOK;
    return 0;
}

// workspace (int32): counts[B*8] | base[B*8]
extern "C" long long gamer_route_perm_workspace_bytes(int B) { return (long long)B * MAX_EXPERTS * 2 * sizeof(int); }

extern "C" int gamer_route_perm_build(const int* pos_idx, int B, int S, int n_experts, void* workspace, int* perm,
                                      int* rows, long long rows_capacity, int* seg_off, cudaStream_t stream) {
    GAMER_REQUIRE(n_experts >= 1 && n_experts <= MAX_EXPERTS, "n_experts=%d out of range", n_experts);
    GAMER_REQUIRE(rows_capacity >= (long long)B * S + 128LL * n_experts, "rows buffer too small");
    int* counts = reinterpret_cast<int*>(workspace);
    int* base = counts + (long long)B * MAX_EXPERTS;
    GAMER_CHECK_CUDA(cudaMemsetAsync(rows, 0xFF, rows_capacity * sizeof(int), stream));
    route_count_kernel<<<B, 128, 0, stream>>>(pos_idx, S, n_experts, counts);
    GAMER_LAUNCH_CHECK();
    route_scan_kernel<<<1, 1024, 0, stream>>>(counts, B, n_experts, base, seg_off);
    GAMER_LAUNCH_CHECK();
    route_fill_kernel<<<ceil_div(B, 4), 128, 0, stream>>>(pos_idx, B, S, n_experts, base, seg_off, perm, rows);
    GAMER_LAUNCH_CHECK();
    return 0;
}

// sort metadata (int32): hist[V] | cursor[V] | bin_start[V+1] | chunk_start[V+1] | sorted_rows[M]
extern "C" long long gamer_embed_sort_bytes(long long M, int vocab) {
    return (long long)(4LL * vocab + 2 + M) * sizeof(int);
}

extern "C" int gamer_embed_sort_build(const long long* ids, long long M, int vocab, int pad, void* sort_buf,
                                      cudaStream_t stream) {
    GAMER_REQUIRE(vocab * sizeof(int) <= 48 * 1024, "vocab too large for the shared-memory histogram");
    int* hist = reinterpret_cast<int*>(sort_buf);
    int* cursor = hist + vocab;
    int* bin_start = cursor + vocab;
    int* chunk_start = bin_start + vocab + 1;
    int* sorted_rows = chunk_start + vocab + 1;
    GAMER_CHECK_CUDA(cudaMemsetAsync(hist, 0, vocab * sizeof(int), stream));
    const int grid = (int)((M + 255) / 256 < 148 * 4 ? (M + 255) / 256 : 148 * 4);
    emb_hist_kernel<<<grid, 256, vocab * sizeof(int), stream>>>(ids, M, vocab, hist);
    GAMER_LAUNCH_CHECK();
    emb_scan_kernel<<<1, 1024, 0, stream>>>(hist, vocab, pad, bin_start, chunk_start, cursor);
    GAMER_LAUNCH_CHECK();
    emb_scatter_kernel<<<grid, 256, 0, stream>>>(ids, M, vocab, bin_start, cursor, sorted_rows);
    GAMER_LAUNCH_CHECK();
    return 0;
}

extern "C" int gamer_embed_bwd(const void* dx_bf16, long long M, int H, int vocab, const void* sort_buf, float* dtable,
                               cudaStream_t stream) {
    GAMER_REQUIRE(H % 8 == 0, "hidden size must be a multiple of 8");
    const int* hist = reinterpret_cast<const int*>(sort_buf);
    const int* bin_start = hist + 2 * vocab;
    const int* chunk_start = bin_start + vocab + 1;
    const int* sorted_rows = chunk_start + vocab + 1;
    const long long max_chunks = M / EMB_CHUNK + vocab + 1;
    const int wpb = 8;
    const int grid = (int)((max_chunks + wpb - 1) / wpb < 148 * 8 ? (max_chunks + wpb - 1) / wpb : 148 * 8);
    emb_reduce_kernel<<<grid, wpb * 32, 0, stream>>>(reinterpret_cast<const bf16*>(dx_bf16), H, bin_start, chunk_start,
                                                     sorted_rows, vocab, dtable);
    GAMER_LAUNCH_CHECK();
    return 0;
}
