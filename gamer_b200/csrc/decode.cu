// K9: constrained beam-search decode.
//   * attn_decode   — one new token per beam row against the user's prompt K/V (stored once per user, shared by all
//                     beams) plus the beam's own generated K/V reached through an ancestry table (no cache reorder): the
//                     tcgen05 kernel attn_decode_kernel of attention_tc.cu;
//   * trie_init     — walk the flat (CSR) candidate trie from the suffix after the last item-terminating token
//                     (SeqRec/generation/trie.py:92-104);
//   * beam_step     — fused full-vocab log-softmax + trie-child mask + running-score add + per-user top-K in shared
//                     memory with warp shuffles (HF _beam_search + PrefixConstrainedLogitsProcessor as driven by
//                     SeqRec/tasks/test_SMB_decoder.py:159-177; the mask is applied AFTER normalisation, quirk Q6).
#include "attention_tc.cuh"
#include "common.cuh"

namespace {

constexpr int D = 64;

// ------------------------------------------------------------------------------------------------------------
// flat trie: node n has children [child_start[n], child_start[n+1]) with tokens child_tok[] (ascending) and node
// ids child_node[]
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int trie_child(const int* child_start, const int* child_tok, const int* child_node, int node,
                                          int tok) {
    int lo = child_start[node], hi = child_start[node + 1];
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const int t = child_tok[mid];
        if (t == tok) return child_node[mid];
        if (t < tok) lo = mid + 1; else hi = mid;
    }
    return -1;
}

// one thread per user: suffix after the last token in `last_set` (bitmap [vocab]); walk from the root.
__global__ void trie_init_kernel(const long long* __restrict__ ids, int B, int L, int vocab,
                                 const unsigned char* __restrict__ last_set, const int* __restrict__ child_start,
                                 const int* __restrict__ child_tok, const int* __restrict__ child_node,
                                 int* __restrict__ node_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const long long* row = ids + (long long)b * L;
    int i = L - 1;
    while (i >= 0) {
        const long long t = row[i];
        if (t >= 0 && t < vocab && last_set[t]) break;
        --i;
    }
    int node = 0;
    for (int j = i + 1; j < L && node >= 0; ++j) node = trie_child(child_start, child_tok, child_node, node, (int)row[j]);
    node_out[b] = node;
}

// one CTA per user.  Candidates = children of every live beam's trie node; score = running + logit - lse(beam).
// Selects the top `beams` by (score desc, flat index asc).  err[0] is set when fewer than `beams` finite candidates.
constexpr int MAX_BEAMS = 32;

__global__ void __launch_bounds__(256) beam_step_kernel(const float* __restrict__ logits, long long ld, int vocab, int beams,
                                                        const float* __restrict__ run_score, const int* __restrict__ node,
                                                        const int* __restrict__ child_start,
                                                        const int* __restrict__ child_tok,
                                                        const int* __restrict__ child_node, float* __restrict__ new_score,
                                                        int* __restrict__ new_parent, int* __restrict__ new_tok,
                                                        int* __restrict__ new_node, int* __restrict__ err,
                                                        int MAX_CAND) {
    __shared__ float s_lse[MAX_BEAMS];
    __shared__ int s_off[MAX_BEAMS + 1];
    extern __shared__ __align__(16) unsigned char dyn[];
    float* c_score = reinterpret_cast<float*>(dyn);                       // [MAX_CAND]
    int* c_edge = reinterpret_cast<int*>(c_score + MAX_CAND);            // index into child_tok/child_node
    short* c_beam = reinterpret_cast<short*>(c_edge + MAX_CAND);
    __shared__ float r_val[8];
    __shared__ int r_idx[8];
    const int u = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // full-vocabulary log-sum-exp per beam (fp32, as HF: logits.float().log_softmax())
    for (int b = warp; b < beams; b += 8) {
        const float* lp = logits + (long long)(u * beams + b) * ld;
        float mx = -INFINITY;
        for (int c = lane; c < vocab; c += 32) mx = fmaxf(mx, lp[c]);
        mx = warp_max(mx);
        float se = 0.f;
        for (int c = lane; c < vocab; c += 32) se += expf(lp[c] - mx);
        se = warp_sum(se);
        if (lane == 0) s_lse[b] = mx + logf(se);
    }
    if (threadIdx.x == 0) {
        int off = 0;
        for (int b = 0; b < beams; ++b) {
            s_off[b] = off;
            const int n = node[u * beams + b];
            off += (n >= 0) ? child_start[n + 1] - child_start[n] : 0;
        }
        s_off[beams] = off;
        if (off > MAX_CAND) err[0] = 2;
    }
    __syncthreads();
    const int n_cand = min(s_off[beams], MAX_CAND);
    for (int b = 0; b < beams; ++b) {
        const int n = node[u * beams + b];
        if (n < 0) continue;
        const int e0 = child_start[n], cnt = child_start[n + 1] - e0;
        const float base = run_score[u * beams + b] - s_lse[b];
        const float* lp = logits + (long long)(u * beams + b) * ld;
        for (int c = threadIdx.x; c < cnt; c += blockDim.x) {
            const int slot = s_off[b] + c;
            if (slot < MAX_CAND) {
                c_score[slot] = base + lp[child_tok[e0 + c]];
                c_edge[slot] = e0 + c;
                c_beam[slot] = (short)b;
            }
        }
    }
    __syncthreads();
    for (int k = 0; k < beams; ++k) {
        float best = -INFINITY;
        int bi = 0x7fffffff;
        for (int c = threadIdx.x; c < n_cand; c += blockDim.x) {
            const float v = c_score[c];
            if (v > best) {            // strided ascending scan keeps the lowest index among equal scores per thread
                best = v;
                bi = c;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) {
                best = ov;
                bi = oi;
            }
        }
        if (lane == 0) {
            r_val[warp] = best;
            r_idx[warp] = bi;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w)
                if (r_val[w] > best || (r_val[w] == best && r_idx[w] < bi)) {
                    best = r_val[w];
                    bi = r_idx[w];
                }
            const int out = u * beams + k;
            if (bi == 0x7fffffff || best == -INFINITY) {
                err[0] = 1;
                new_score[out] = -INFINITY;
                new_parent[out] = 0;
                new_tok[out] = 0;
                new_node[out] = -1;
            } else {
                new_score[out] = best;
                new_parent[out] = c_beam[bi];
                new_tok[out] = child_tok[c_edge[bi]];
                new_node[out] = child_node[c_edge[bi]];
                c_score[bi] = -INFINITY;
            }
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int gamer_attn_decode(const void* qcur, const void* pk, const void* pv, long long ld_p, const void* gen_k,
                                 const void* gen_v, long long gen_step_stride, long long ld_g, const int* anc, int B,
                                 int beams, int L0, int n_gen, int n_q, int n_kv, int head_dim, int S_max,
                                 const int* am, const int* act, const int* sess, int mask_kind, const float* vmean,
                                 float scale, void* o, long long ld_o, cudaStream_t stream) {
    GAMER_REQUIRE(head_dim == D, "decode attention is specialised for head_dim 64");
    GAMER_REQUIRE(n_kv > 0 && n_q == 2 * n_kv, "decode attention is specialised for two query heads per kv head (got %d / %d)",
                  n_q, n_kv);
    GAMER_REQUIRE(beams >= 1 && beams <= 64, "num_beams = %d exceeds the 64 beam rows one CTA handles", beams);
    GAMER_REQUIRE(L0 >= 1 && L0 <= 4096, "prompt length %d out of range (1..4096)", L0);
    const bool cross = mask_kind == MASK_MULTI_CROSS || mask_kind == MASK_SESSION_CROSS;
    GAMER_REQUIRE(!cross || (act != nullptr && vmean != nullptr), "cross decode attention needs actions and vmean");
    GAMER_REQUIRE(mask_kind != MASK_SESSION_CROSS || sess != nullptr, "SESSION_CROSS needs session ids");
    GAMER_REQUIRE(vmean != nullptr || !cross, "decode attention needs the prompt value mean for rows without an allowed key");
    if (B == 0) return 0;
    return attn_tc_decode(qcur, pk, pv, ld_p, gen_k, gen_v, gen_step_stride, ld_g, anc, B, beams, L0, n_gen, n_q, n_kv,
                          S_max, am, act, sess, mask_kind, vmean, scale, o, ld_o, stream);
}

extern "C" int gamer_trie_init(const long long* ids, int B, int L, int vocab, const unsigned char* last_set,
                               const int* child_start, const int* child_tok, const int* child_node, int* node_out,
                               cudaStream_t stream) {
    if (B == 0) return 0;
    trie_init_kernel<<<ceil_div(B, 128), 128, 0, stream>>>(ids, B, L, vocab, last_set, child_start, child_tok, child_node,
                                                           node_out);
    GAMER_LAUNCH_CHECK();
    return 0;
}

extern "C" int gamer_beam_step(const float* logits, long long ld, int vocab, int n_users, int beams,
                               const float* run_score, const int* node, const int* child_start, const int* child_tok,
                               const int* child_node, int max_children, float* new_score, int* new_parent,
                               int* new_tok, int* new_node, int* err, cudaStream_t stream) {
    GAMER_REQUIRE(beams >= 1 && beams <= MAX_BEAMS, "num_beams=%d out of range (1..%d)", beams, MAX_BEAMS);
    if (n_users == 0) return 0;
    const int cap = beams * max_children;                      // every beam's node has <= max_children children
    const size_t smem = (size_t)cap * (sizeof(float) + sizeof(int) + sizeof(short)) + 16;
    GAMER_REQUIRE(smem <= 200 * 1024, "beams * max_children = %d candidates do not fit in shared memory", cap);
    static PerDeviceOnce configured;
    if (configured.need(smem))
        GAMER_CHECK_CUDA(cudaFuncSetAttribute(beam_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    beam_step_kernel<<<n_users, 256, smem, stream>>>(logits, ld, vocab, beams, run_score, node, child_start, child_tok,
                                                     child_node, new_score, new_parent, new_tok, new_node, err, cap);
    GAMER_LAUNCH_CHECK();
    return 0;
}
