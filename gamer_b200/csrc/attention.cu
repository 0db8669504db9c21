// K6: session / behaviour-masked attention, forward and backward, GQA 2:1, head_dim 64 — the C-ABI entry points.
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
// The kernels are the tcgen05 ones of attention_tc.cu; there is no second backend: shapes outside their specialisation
// (head_dim 64, two query heads per kv head, L <= 4096) are rejected.
#include "attention_tc.cuh"
#include "common.cuh"

namespace {

constexpr int D = 64;

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

int check_common(int L, int n_q, int n_kv, int head_dim, int kind, const int* act, const int* sess) {
    GAMER_REQUIRE(kind >= 0 && kind <= 3, "unknown mask kind %d", kind);
    GAMER_REQUIRE(kind == MASK_CAUSAL || kind == MASK_SESSION || act != nullptr, "mask kind %d needs `actions`", kind);
    GAMER_REQUIRE(kind == MASK_CAUSAL || kind == MASK_MULTI_CROSS || sess != nullptr, "mask kind %d needs `session_ids`", kind);
    GAMER_REQUIRE(attn_tc_supported(L, n_q, n_kv, head_dim),
                  "attention kernels are specialised for head_dim 64, n_q = 2 n_kv and L <= 4096 (got head_dim %d, n_q %d, "
                  "n_kv %d, L %d)", head_dim, n_q, n_kv, L);
    return 0;
}

long long vmean_bytes(int B, int n_kv) { return ((long long)B * n_kv * D * sizeof(float) + 255) / 256 * 256; }

}  // namespace

extern "C" long long gamer_attn_workspace_bytes(int B, int L, int n_q, int n_kv) {
    // vmean [B, n_kv, 64] fp32 first (callers read it back), then the key codes
    return vmean_bytes(B, n_kv) + attn_tc_fwd_ws_bytes(B, L);
}

extern "C" long long gamer_attn_keep_bytes(int B, int L, int n_q) { return attn_tc_keep_bytes(B, L, n_q); }

extern "C" int gamer_attn_fwd(const void* q, const void* k, const void* v, long long ld, int B, int L, int n_q, int n_kv,
                              int head_dim, int mask_kind, int tokens_per_item, const int* am, const int* act,
                              const int* sess, float scale, void* workspace, void* o, long long ld_o, float* lse,
                              const gamer_dropout_t* drop, void* keep, cudaStream_t stream) {
    if (B == 0 || L == 0) return 0;
    if (int e = check_common(L, n_q, n_kv, head_dim, mask_kind, act, sess)) return e;
    float* vmean = reinterpret_cast<float*>(workspace);
    v_colmean_kernel<<<B * n_kv, 256, 0, stream>>>(reinterpret_cast<const bf16*>(v), ld, L, n_kv, vmean);
    GAMER_LAUNCH_CHECK();
    return attn_tc_fwd(q, k, v, ld, B, L, n_q, n_kv, mask_kind, tokens_per_item, am, act, sess, scale, vmean,
                       reinterpret_cast<uint8_t*>(workspace) + vmean_bytes(B, n_kv), o, ld_o, lse, drop, keep, stream);
}

extern "C" long long gamer_attn_bwd_workspace_bytes(int B, int L, int n_q) { return attn_tc_bwd_ws_bytes(B, L, n_q); }

extern "C" int gamer_attn_bwd(const void* q, const void* k, const void* v, long long ld, int B, int L, int n_q, int n_kv,
                              int head_dim, int mask_kind, int tokens_per_item, const int* am, const int* act,
                              const int* sess, float scale, const void* o, const void* d_o, long long ld_o,
                              const float* lse, void* workspace, void* dq, void* dk, void* dv, long long ld_d,
                              const gamer_dropout_t* drop, const void* keep, cudaStream_t stream) {
    if (B == 0 || L == 0) return 0;
    if (int e = check_common(L, n_q, n_kv, head_dim, mask_kind, act, sess)) return e;
    return attn_tc_bwd(q, k, v, ld, B, L, n_q, n_kv, mask_kind, tokens_per_item, am, act, sess, scale, o, d_o, ld_o, lse,
                       workspace, dq, dk, dv, ld_d, drop, keep, stream);
}
