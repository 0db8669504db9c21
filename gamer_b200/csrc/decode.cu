// K9: constrained beam-search decode.
//   * attn_decode   — one new token per beam row against the user's prompt K/V (stored once per user, shared by all
//                     beams) plus the beam's own generated K/V reached through an ancestry table (no cache reorder);
//   * trie_init     — walk the flat (CSR) candidate trie from the suffix after the last item-terminating token
//                     (SeqRec/generation/trie.py:92-104);
//   * beam_step     — fused full-vocab log-softmax + trie-child mask + running-score add + per-user top-K in shared
//                     memory with warp shuffles (HF _beam_search + PrefixConstrainedLogitsProcessor as driven by
//                     SeqRec/tasks/test_SMB_decoder.py:159-177; the mask is applied AFTER normalisation, quirk Q6).
#include "common.cuh"

namespace {

constexpr int D = 64;
constexpr int KROW = 66;   // padded smem row (bf16) -> conflict-free column access

struct DecodeArgs {
    const bf16* qcur;        // [R, ld_g], q head h at column h*64
    const bf16* pk;          // prompt keys   [B*L0, ld_p] (+ kvh*64)
    const bf16* pv;          // prompt values
    long long ld_p;
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

__global__ void __launch_bounds__(256) attn_decode_kernel(DecodeArgs a) {
    __shared__ bf16 sK[64 * KROW];
    __shared__ bf16 sV[64 * KROW];
    __shared__ int sOk[64];
    extern __shared__ float sQ[];   // [n_queries][64] fp32

    const int u = blockIdx.x, kvh = blockIdx.y;
    const int group = a.n_q / a.n_kv;
    const int nqv = a.beams * group;          // query vectors handled by this CTA
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool cross = (a.kind == MASK_MULTI_CROSS || a.kind == MASK_SESSION_CROSS);

    for (int x = threadIdx.x; x < nqv * D; x += blockDim.x) {
        const int qi = x / D, d = x % D;
        const int r = u * a.beams + qi / group, h = kvh * group + qi % group;
        sQ[x] = __bfloat162float(a.qcur[(long long)r * a.ld_g + h * D + d]);
    }
    int act_last = 0, sess_last = 0;
    if (cross) {
        act_last = a.act[(long long)u * a.L0 + a.L0 - 1];
        if (a.sess) sess_last = a.sess[(long long)u * a.L0 + a.L0 - 1];
    }
    constexpr int QPW = 8;                       // queries per warp (8 warps x 8 >= 40 for 20 beams x 2 heads)
    float m[QPW], l[QPW], acc0[QPW], acc1[QPW];
#pragma unroll
    for (int i = 0; i < QPW; ++i) {
        m[i] = -INFINITY;
        l[i] = 0.f;
        acc0[i] = acc1[i] = 0.f;
    }
    const int n_tiles = (a.L0 + 63) / 64;
    for (int t = 0; t < n_tiles; ++t) {
        __syncthreads();
        const int j0 = t * 64;
        for (int x = threadIdx.x; x < 64 * 8; x += blockDim.x) {
            const int row = x >> 3, ch = x & 7;
            const int j = j0 + row;
            uint4 kk = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
            if (j < a.L0) {
                kk = *reinterpret_cast<const uint4*>(a.pk + ((long long)u * a.L0 + j) * a.ld_p + kvh * D + ch * 8);
                vv = *reinterpret_cast<const uint4*>(a.pv + ((long long)u * a.L0 + j) * a.ld_p + kvh * D + ch * 8);
            }
            uint32_t* dk = reinterpret_cast<uint32_t*>(sK + row * KROW + ch * 8);
            uint32_t* dv = reinterpret_cast<uint32_t*>(sV + row * KROW + ch * 8);
            dk[0] = kk.x; dk[1] = kk.y; dk[2] = kk.z; dk[3] = kk.w;
            dv[0] = vv.x; dv[1] = vv.y; dv[2] = vv.z; dv[3] = vv.w;
        }
        if (threadIdx.x < 64) {
            const int j = j0 + threadIdx.x;
            int ok = 0;
            if (j < a.L0) {
                const long long idx = (long long)u * a.L0 + j;
                ok = a.am[idx];
                if (cross) {
                    ok = ok && (a.act[idx] < act_last);
                    if (a.kind == MASK_SESSION_CROSS) ok = ok && (a.sess[idx] < sess_last);
                }
            }
            sOk[threadIdx.x] = ok;
        }
        __syncthreads();
#pragma unroll
        for (int qq = 0; qq < QPW; ++qq) {
            const int qi = warp + qq * 8;
            if (qi >= nqv) break;
            const float* qv = sQ + qi * D;
            float s0 = 0.f, s1 = 0.f;
            const uint32_t* k0 = reinterpret_cast<const uint32_t*>(sK + lane * KROW);
            const uint32_t* k1 = reinterpret_cast<const uint32_t*>(sK + (lane + 32) * KROW);
#pragma unroll 8
            for (int w = 0; w < 32; ++w) {
                const float2 a0 = unpack_bf16(k0[w]), a1 = unpack_bf16(k1[w]);
                s0 += qv[2 * w] * a0.x + qv[2 * w + 1] * a0.y;
                s1 += qv[2 * w] * a1.x + qv[2 * w + 1] * a1.y;
            }
            s0 = sOk[lane] ? s0 * a.scale_log2 : -INFINITY;
            s1 = sOk[lane + 32] ? s1 * a.scale_log2 : -INFINITY;
            const float mx = warp_max(fmaxf(s0, s1));
            const float mnew = fmaxf(m[qq], mx);
            if (mnew == -INFINITY) continue;           // warp-uniform
            const float alpha = exp2f(m[qq] - mnew);
            const float p0 = exp2f(s0 - mnew), p1 = exp2f(s1 - mnew);
            l[qq] = l[qq] * alpha + warp_sum(p0 + p1);
            float o0 = acc0[qq] * alpha, o1 = acc1[qq] * alpha;
            for (int j = 0; j < 32; ++j) {
                const float pj = __shfl_sync(0xffffffffu, p0, j);
                const float2 v = unpack_bf16(*reinterpret_cast<const uint32_t*>(sV + j * KROW + 2 * lane));
                o0 += pj * v.x;
                o1 += pj * v.y;
            }
            for (int j = 0; j < 32; ++j) {
                const float pj = __shfl_sync(0xffffffffu, p1, j);
                const float2 v = unpack_bf16(*reinterpret_cast<const uint32_t*>(sV + (j + 32) * KROW + 2 * lane));
                o0 += pj * v.x;
                o1 += pj * v.y;
            }
            m[qq] = mnew;
            acc0[qq] = o0;
            acc1[qq] = o1;
        }
    }
    // generated keys (the beam's own ancestry, current token included) and the epilogue
#pragma unroll
    for (int qq = 0; qq < QPW; ++qq) {
        const int qi = warp + qq * 8;
        if (qi >= nqv) break;
        const int r = u * a.beams + qi / group, h = kvh * group + qi % group;
        const float* qv = sQ + qi * D;
        float gsum0 = 0.f, gsum1 = 0.f;     // unmasked sum of generated values (uniform-row fallback)
        for (int s = 0; s < a.n_gen; ++s) {
            const int slot = a.anc[(long long)r * a.S_max + s];
            const bf16* kp = a.gen_k + s * a.gen_step_stride + (long long)slot * a.ld_g + kvh * D;
            const bf16* vp = a.gen_v + s * a.gen_step_stride + (long long)slot * a.ld_g + kvh * D;
            const float2 kv2 = unpack_bf16(*reinterpret_cast<const uint32_t*>(kp + 2 * lane));
            const float2 vv2 = unpack_bf16(*reinterpret_cast<const uint32_t*>(vp + 2 * lane));
            gsum0 += vv2.x;
            gsum1 += vv2.y;
            if (cross) continue;            // generated columns are masked for the cross rows (model.py:605-617)
            float sc = warp_sum(qv[2 * lane] * kv2.x + qv[2 * lane + 1] * kv2.y) * a.scale_log2;
            const float mnew = fmaxf(m[qq], sc);
            const float alpha = exp2f(m[qq] - mnew), p = exp2f(sc - mnew);
            l[qq] = l[qq] * alpha + p;
            acc0[qq] = acc0[qq] * alpha + p * vv2.x;
            acc1[qq] = acc1[qq] * alpha + p * vv2.y;
            m[qq] = mnew;
        }
        float o0, o1;
        if (l[qq] > 0.f) {
            o0 = acc0[qq] / l[qq];
            o1 = acc1[qq] / l[qq];
        } else {
            // no allowed key: uniform over ALL cached keys (quirk Q1) = (L0 * mean(prompt V) + sum gen V) / (L0 + n_gen)
            const float* vm = a.vmean + ((long long)u * a.n_kv + kvh) * D;
            const float inv = 1.0f / (float)(a.L0 + a.n_gen);
            o0 = (vm[2 * lane] * a.L0 + gsum0) * inv;
            o1 = (vm[2 * lane + 1] * a.L0 + gsum1) * inv;
        }
        *reinterpret_cast<uint32_t*>(a.o + (long long)r * a.ld_o + h * D + 2 * lane) = pack_bf16(o0, o1);
    }
}

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
    GAMER_REQUIRE(n_kv > 0 && n_q % n_kv == 0, "n_q must be a multiple of n_kv");
    const int nqv = beams * (n_q / n_kv);
    GAMER_REQUIRE(nqv <= 64, "beams * (n_q / n_kv) = %d exceeds the 64 query vectors one CTA handles", nqv);
    const bool cross = mask_kind == MASK_MULTI_CROSS || mask_kind == MASK_SESSION_CROSS;
    GAMER_REQUIRE(!cross || (act != nullptr && vmean != nullptr), "cross decode attention needs actions and vmean");
    GAMER_REQUIRE(mask_kind != MASK_SESSION_CROSS || sess != nullptr, "SESSION_CROSS needs session ids");
    if (B == 0) return 0;
    DecodeArgs a;
    a.qcur = reinterpret_cast<const bf16*>(qcur); a.pk = reinterpret_cast<const bf16*>(pk);
    a.pv = reinterpret_cast<const bf16*>(pv); a.ld_p = ld_p;
    a.gen_k = reinterpret_cast<const bf16*>(gen_k); a.gen_v = reinterpret_cast<const bf16*>(gen_v);
    a.gen_step_stride = gen_step_stride; a.ld_g = ld_g; a.anc = anc;
    a.B = B; a.beams = beams; a.L0 = L0; a.n_gen = n_gen; a.n_q = n_q; a.n_kv = n_kv; a.S_max = S_max;
    a.am = am; a.act = act; a.sess = sess; a.kind = mask_kind; a.vmean = vmean;
    a.scale_log2 = scale * 1.4426950408889634f;
    a.o = reinterpret_cast<bf16*>(o); a.ld_o = ld_o;
    dim3 grid(B, n_kv);
    attn_decode_kernel<<<grid, 256, nqv * D * sizeof(float), stream>>>(a);
    GAMER_LAUNCH_CHECK();
    return 0;
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
