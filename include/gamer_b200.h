/* gamer_b200 — C-ABI of the B200-native GAMER decoder hot path.
 *
 * The reference (wzf2000/GAMER) is pure Python/PyTorch and has no FFI of its own; the boundary its hot path sits
 * behind is the HF-model Python surface.  This header is the native boundary underneath that
 * surface: every entry point replaces the stock PyTorch/library kernels executed by the reference call site cited
 * next to it (paths relative to the reference's SeqRec/ package).  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - plain C: raw pointers, sizes and a cudaStream_t; no torch types.  All data pointers are DEVICE pointers owned
 *     by the caller; nothing is allocated inside (workspace sizes come from the *_bytes helpers).
 *   - every call is asynchronous on `stream`; return 0 = ok, negative = error, message via gamer_last_error()
 *     (thread-local).  One host thread per device.
 *   - activations are bf16 (uint16 storage), statistics / gradients of parameters fp32, indices int32 unless noted.
 *   - "ld" arguments are row strides in ELEMENTS.
 */
#ifndef GAMER_B200_H
#define GAMER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* gamer_stream_t; /* == cudaStream_t */

#define GAMER_MASK_CAUSAL 0        /* Qwen3Multi self:   j<=i & am[j]                     Qwen3Multi/model.py:691-741 */
#define GAMER_MASK_MULTI_CROSS 1   /* Qwen3Multi cross:  j<=i & act[j]<act[i] & am[j]      Qwen3Multi/model.py:573-630 */
#define GAMER_MASK_SESSION 2       /* Session* self:     (item(j)==item(i)&j<=i | s[j]<s[i]) & am[j]
                                                                                        Qwen3SessionMoe/model.py:416-468 */
#define GAMER_MASK_SESSION_CROSS 3 /* SessionMulti cross: s[j]<s[i] & act[j]<act[i] & am[j]
                                                                                     Qwen3SessionMulti/model.py:556-613 */

const char* gamer_last_error(void);

/* Dropout of one call site (nn.Dropout(config.dropout_rate): Qwen3Multi/model.py:177,217,235,241, Qwen3Moe/FFN.py:23-26;
 * SDPA dropout_p = config.attention_dropout: Qwen3Multi/model.py:139).  HOST struct; NULL or p <= 0 = no dropout
 * (eval mode).  Masks are Philox4x32-7 bits indexed by (seed, offset, site, row, column) — the backward entry points
 * take the same struct and regenerate them.  Hidden-state sites drop with probability floor(p*65536)/65536, attention
 * probabilities with floor(p*256)/256; kept elements are scaled by the reciprocal of the realised keep probability. */
typedef struct {
    unsigned long long seed; /* generator seed (per rank) */
    unsigned int offset;     /* advances once per forward pass (micro-batch) */
    unsigned int site;       /* dropout call within the pass: layer * 8 + {0 self P, 1 self out, 2 cross P, 3 cross out,
                                4 expert inner, 5 FFN out} */
    float p;
    const unsigned int* offset_dev; /* optional DEVICE word XOR-ed into the key at kernel run time: lets a captured CUDA
                                       graph draw fresh masks on every replay (NULL = none) */
} gamer_dropout_t;

/* ---- K1: embedding gather + router indices ------------------------------------------------------------------
 * replaces embed_tokens(input_ids) (Qwen3Multi/model.py:779) and Qwen3MultiDecoderRouter.forward
 * (Qwen3Multi/router.py:74-201; Qwen3Moe/router.py:74-154).  ids/ctx are int64 as the reference passes them.
 * Token s of row b sits at absolute position pos0+s; ctx (may be NULL = ids) is the whole sequence so far.
 * A token id outside [0, vocab) gathers the pad row and sets bit 0 of *err (device int32, may be NULL): nn.Embedding
 * raises IndexError there, the caller reads the word when it next synchronises. */
int gamer_embed_route_fwd(const long long* ids, const long long* ctx, long long ctx_ld, int B, int S, int pos0,
                          int tokens_per_item, int pad, int eos, int vocab, const int* beh_lut, int n_beh,
                          const void* table_bf16, int H, void* x_bf16, int* pos_idx, int* beh_idx, int* act_idx,
                          int* err, gamer_stream_t stream);
/* expert routing permutation for MyQwen3SparseMLP (Qwen3Moe/FFN.py:53-72): expert = position index. */
long long gamer_route_perm_workspace_bytes(int B);
int gamer_route_perm_build(const int* pos_idx, int B, int S, int n_experts, void* workspace, int* perm, int* rows,
                           long long rows_capacity, int* seg_off, gamer_stream_t stream);
/* sparse embedding gradient (embedding_dense_backward of nn.Embedding(padding_idx=4), Qwen3Multi/model.py:263). */
long long gamer_embed_sort_bytes(long long M, int vocab);
int gamer_embed_sort_build(const long long* ids, long long M, int vocab, int pad, void* sort_buf, gamer_stream_t stream);
int gamer_embed_bwd(const void* dx_bf16, long long M, int H, int vocab, const void* sort_buf, float* dtable,
                    gamer_stream_t stream);

/* ---- K2/K3: RMSNorm, head norm + behaviour embedding + RoPE ---------------------------------------------------
 * Qwen3RMSNorm (Qwen3Multi/model.py:165-176,205,222,239,284,869); q/k norm, behaviour embeddings and
 * apply_rotary_pos_emb (Qwen3Multi/model.py:88-101); FFN behaviour-embedding concat (Qwen3Moe/FFN.py:60-62).
 * cos_tab / sin_tab: fp32 [n_pos, head_dim/2]; a position (pos_ids[m], or m % L + pos0) outside [0, n_pos) is clamped
 * to the nearest table row instead of read out of bounds — size the tables for the largest position the caller uses. */
int gamer_rmsnorm_fwd(const void* x, const float* w, float eps, long long M, int H, void* out, long long ld_out,
                      const int* row_map, const void* cat_table, const int* cat_idx, int cat_dim, float* rstd,
                      gamer_stream_t stream);
int gamer_rmsnorm_bwd(const void* x, const float* w, const float* rstd, float eps, long long M, int H, const void* dh,
                      long long ld_dh, const int* row_map, const void* dres, void* dx, float* dw, const int* cat_idx,
                      int cat_dim, int cat_rows, float* dcat, gamer_stream_t stream);
int gamer_qk_norm_rope_fwd(const void* raw, long long ld_raw, void* out, long long ld_out, long long M, int L, int n_q,
                           int n_kv, int head_dim, const int* pos_ids, int pos0, int n_pos, const float* cos_tab,
                           const float* sin_tab, const float* qn_w, const float* kn_w, const void* q_emb,
                           const void* k_emb, const void* v_emb, const int* act_idx, float eps, gamer_stream_t stream);
int gamer_qk_norm_rope_bwd(const void* raw, long long ld_raw, const void* dout, long long ld_dout, void* draw,
                           long long ld_draw, long long M, int L, int n_q, int n_kv, int head_dim, const int* pos_ids,
                           int pos0, int n_pos, const float* cos_tab, const float* sin_tab, const float* qn_w,
                           const float* kn_w, const void* q_emb, const void* k_emb, const void* v_emb, const int* act_idx, int emb_rows,
                           float eps, float* d_qn_w, float* d_kn_w, float* d_q_emb, float* d_k_emb, float* d_v_emb,
                           gamer_stream_t stream);

/* ---- K4/K7: tcgen05 GEMMs ------------------------------------------------------------------------------------
 * q/k/v/o/gating nn.Linear (Qwen3Multi/model.py:38-49,66,93-99,147-149), expert gate/up/down projections
 * (Qwen3Moe/FFN.py:19-27,64-68) and lm_head (Qwen3Multi/model.py:1001).
 *   C[r, n] = alpha * sum_k A[r,k] * B[g(r)*N + n, k]  (+ resid[out_row, n]);  out_row = row_map ? row_map[r] : r
 * grouped mode: seg_off[n_groups+1] (device) gives 128-aligned row segments, segment g uses weight slab g. */
int gamer_gemm_bf16_tn(const void* A, long long lda, int rows, const void* B, long long ldb, int n_groups, int N, int K,
                       const int* seg_off, void* C, long long ldc, int c_is_f32, const void* resid, long long ldr,
                       const int* row_map, float alpha, const gamer_dropout_t* drop, gamer_stream_t stream);
/* dW[g][i, j] += sum_r dY[r, i] * X[r, j]   (fp32 accumulate into dW [n_groups, N_out, K_in]) */
int gamer_gemm_bf16_wgrad(const void* dY, long long ldy, const void* X, long long ldx, int rows, int N_out, int K_in,
                          int n_groups, const int* seg_off, float* dW, gamer_stream_t stream);
/* CUDA-core checker used by the GPU tests only */
int gamer_ref_gemm_tn(const void* A, long long lda, const void* B, long long ldb, float* C, long long ldc, int rows,
                      int N, int K, gamer_stream_t stream);

/* ---- K6: masked attention (replaces mask materialisation + SDPA, Qwen3Multi/model.py:123-143,573-741) ---------- */
long long gamer_attn_workspace_bytes(int B, int L, int n_q, int n_kv);
/* Attention-probability dropout (SDPA dropout_p, Qwen3Multi/model.py:139): the forward writes one keep word per (query,
 * 32 keys) into `keep` (gamer_attn_keep_bytes bytes; may be NULL when drop is NULL or drop->p == 0) and the backward of the
 * same call site reads it back, instead of regenerating the random stream. */
long long gamer_attn_keep_bytes(int B, int L, int n_q);
/* debug hook (tools/attn_trace.py): record an in-kernel timeline of CTA 0 of the next attention launches; buf = NULL disables */
int gamer_attn_set_trace(void* buf, int cap);
int gamer_attn_fwd(const void* q, const void* k, const void* v, long long ld, int B, int L, int n_q, int n_kv,
                   int head_dim, int mask_kind, int tokens_per_item, const int* am, const int* act, const int* sess,
                   float scale, void* workspace, void* o, long long ld_o, float* lse, const gamer_dropout_t* drop,
                   void* keep, gamer_stream_t stream);
long long gamer_attn_bwd_workspace_bytes(int B, int L, int n_q);
int gamer_attn_bwd(const void* q, const void* k, const void* v, long long ld, int B, int L, int n_q, int n_kv,
                   int head_dim, int mask_kind, int tokens_per_item, const int* am, const int* act, const int* sess,
                   float scale, const void* o, const void* d_o, long long ld_o, const float* lse, void* workspace,
                   void* dq, void* dk, void* dv, long long ld_d, const gamer_dropout_t* drop, const void* keep,
                   gamer_stream_t stream);

/* ---- elementwise pieces --------------------------------------------------------------------------------------- */
/* act = dropout(silu(gate) * up) (Qwen3Moe/FFN.py:26).  row_ids (may be NULL = identity) maps a row of the expert-
 * permuted space to its token row (-1 = padding row): the dropout mask is indexed by token row. */
int gamer_swiglu_fwd(const void* gu, long long ld_gu, void* act, long long ld_act, long long R, int I,
                     const int* row_ids, const gamer_dropout_t* drop, gamer_stream_t stream);
int gamer_swiglu_bwd(const void* gu, long long ld_gu, const void* dact, long long ld_dact, void* dgu, long long ld_dgu,
                     long long R, int I, const int* row_ids, const gamer_dropout_t* drop, gamer_stream_t stream);
/* out = x + dropout(y * silu(g))   (cross attention: o_proj(a) * silu(gating(h)), Qwen3Multi/model.py:146-147,235) */
int gamer_gate_residual_fwd(const void* x, const void* y, const void* g, long long ld_g, void* out, long long R, int W,
                            const gamer_dropout_t* drop, gamer_stream_t stream);
int gamer_gate_residual_bwd(const void* dout, const void* y, const void* g, long long ld_g, void* dy, void* dg,
                            long long ld_dg, long long R, int W, const gamer_dropout_t* drop, gamer_stream_t stream);
/* dst[r] = dropout_mask[rows[r]] * src[rows[r]]  (0 where rows[r] < 0) */
int gamer_gather_rows(const void* src, long long ld_src, const int* rows, const int* n_rows_dev, long long n_rows_max,
                      void* dst, long long ld_dst, int W, const gamer_dropout_t* drop, gamer_stream_t stream);
/* dst[r, 0..W) = 0 where rows[r] < 0: the padding rows of the expert-permuted token space (the rows the reference never
 * materialises: Qwen3Moe/FFN.py:64-68 indexes each expert's tokens), so that the grouped GEMMs see finite operands there */
int gamer_zero_unmapped_rows(void* dst, long long ld_dst, const int* rows, long long n_rows, int W, gamer_stream_t stream);
/* out = dropout(in) with the mask of `drop` (backward of a residual-branch dropout fused into a GEMM epilogue) */
int gamer_dropout_apply(const void* in, void* out, long long R, int W, const gamer_dropout_t* drop,
                        gamer_stream_t stream);

/* ---- K8: fused softmax cross-entropy (ForCausalLMLoss, Qwen3Multi/model.py:904-922) ---------------------------- */
int gamer_ce_fwd_bwd(const float* logits, long long ld_l, const long long* labels, long long R, int V, int ignore_index,
                     const float* inv_norm, float grad_scale, float* loss_row, void* dlogits, long long ld_d,
                     gamer_stream_t stream);

/* ---- K9: constrained beam-search decode ------------------------------------------------------------------------
 * replaces, per generated token, HF GenerationMixin._beam_search + PrefixConstrainedLogitsProcessor + the Python Trie
 * walk (SeqRec/tasks/test_SMB_decoder.py:159-177, SeqRec/generation/trie.py:5-104) and the cached attention of
 * Qwen3MultiAttention.forward with DynamicCache (Qwen3Multi/model.py:118-143, masks :605-617,:717-728).
 * The flat trie is CSR: node n has children [child_start[n], child_start[n+1]) with ascending child_tok[]. */
int gamer_attn_decode(const void* qcur, const void* pk, const void* pv, long long ld_p, const void* gen_k,
                      const void* gen_v, long long gen_step_stride, long long ld_g, const int* anc, int B, int beams,
                      int L0, int n_gen, int n_q, int n_kv, int head_dim, int S_max, const int* am, const int* act,
                      const int* sess, int mask_kind, const float* vmean, float scale, void* o, long long ld_o,
                      gamer_stream_t stream);
int gamer_trie_init(const long long* ids, int B, int L, int vocab, const unsigned char* last_set, const int* child_start,
                    const int* child_tok, const int* child_node, int* node_out, gamer_stream_t stream);
int gamer_beam_step(const float* logits, long long ld, int vocab, int n_users, int beams, const float* run_score,
                    const int* node, const int* child_start, const int* child_tok, const int* child_node,
                    int max_children, float* new_score, int* new_parent, int* new_tok, int* new_node, int* err,
                    gamer_stream_t stream);

/* ---- input pipeline ---------------------------------------------------------------------
 * One launch builds the collators' batch tensors from the pre-tokenised interaction store: replaces
 * DecoderOnlyCollator / DecoderOnlyTestCollator (SeqRec/datasets/collator.py:47-107,149-207), the per-sample
 * session / extended-session / action arrays (SeqRec/datasets/SMB_dataset.py:194-234) and the target-behaviour column
 * of SeqRec/tasks/test_SMB_decoder.py:105-117.  Store: item_tokens int32 [T,4], behavior int16 [T], session int32 [T],
 * offsets int64 [N+1] (user u owns rows offsets[u]..offsets[u+1]-1, oldest first).  Row r = the last <= n_max items of
 * users[r], 5 tokens per item, `width` items per row, right-padded (left_pad = 0, training) or left-padded (1,
 * evaluation); target_behavior >= 0 appends one column (that behaviour's token, session max+1, extended max+1, its
 * level).  Outputs int64 [n_users, 5*width (+1)]; labels may be NULL. */
int gamer_collate_sessions(const int* item_tokens, const short* behavior, const int* session, const long long* offsets,
                           const long long* users, int n_users, int n_max, int width, int left_pad,
                           const long long* beh_tokens, const long long* beh_level, int n_beh, long long pad,
                           int target_behavior, long long* input_ids, long long* attention_mask, long long* labels,
                           long long* session_ids, long long* extended_session_ids, long long* actions,
                           gamer_stream_t stream);

/* ---- optimizer over the flat fused parameter buffer ------------------------------------
 * AdamW as HF Trainer configures it (SeqRec/tasks/train_SMB_decoder.py:396-428: adamw_torch, weight decay on non-norm
 * weights, clip_grad_norm_(max_grad_norm)).  hp (device floats): lr, 1-beta1^t, 1-beta2^t. */
int gamer_sumsq_accumulate(const float* g, long long n, float* out, gamer_stream_t stream);
int gamer_adamw_step(float* p, const float* g, float* m, float* v, const unsigned char* decay_mask, void* p_bf16,
                     long long n, const float* hp, float beta1, float beta2, float eps, float weight_decay,
                     const float* gnorm_sq, float max_grad_norm, float grad_scale, gamer_stream_t stream);
int gamer_cast_f32_bf16(const float* src, void* dst, long long n, gamer_stream_t stream);
/* The transposed (dgrad) copies of every weight, rewritten in ONE launch after an optimizer step (what autograd of
 * nn.Linear gets for free from the same storage: SeqRec/models/generative/Qwen3Multi/model.py:38-49,66,147-149).
 * desc (device): n_mats x {src, dst, rows, cols, ld_src, ld_dst} as int64, bf16 matrices, dst[c][r] = src[r][c];
 * tile_start (device int32[n_mats + 1]): prefix sum of ceil(rows/32) * ceil(cols/32); total_tiles = tile_start[n_mats]. */
int gamer_transpose_bf16_batch(const long long* desc, const int* tile_start, int n_mats, int total_tiles,
                               gamer_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GAMER_B200_H */
