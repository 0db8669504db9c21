// tcgen05 attention (attention_tc.cu): internal entry points used by the C-ABI wrappers in attention.cu.
#pragma once
#include "common.cuh"

bool attn_tc_supported(int L, int n_q, int n_kv, int head_dim);
long long attn_tc_fwd_ws_bytes(int B, int L);
int attn_tc_fwd(const void* q, const void* k, const void* v, long long ld, int B, int L, int n_q, int n_kv, int kind, int P,
                const int* am, const int* act, const int* sess, float scale, const float* vmean, void* ws, void* o,
                long long ld_o, float* lse, const gamer_dropout_t* drop, void* keep, cudaStream_t stream);
long long attn_tc_keep_bytes(int B, int L, int n_q);
long long attn_tc_bwd_ws_bytes(int B, int L, int n_q);
int attn_tc_bwd(const void* q, const void* k, const void* v, long long ld, int B, int L, int n_q, int n_kv, int kind, int P,
                const int* am, const int* act, const int* sess, float scale, const void* o, const void* d_o, long long ld_o,
                const float* lse, void* ws, void* dq, void* dk, void* dv, long long ld_d, const gamer_dropout_t* drop,
                const void* keep, cudaStream_t stream);
// decode attention (one new token per beam row): see attn_decode_kernel in attention_tc.cu
int attn_tc_decode(const void* qcur, const void* pk, const void* pv, long long ld_p, const void* gen_k, const void* gen_v,
                   long long gen_step_stride, long long ld_g, const int* anc, int B, int beams, int L0, int n_gen, int n_q,
                   int n_kv, int S_max, const int* am, const int* act, const int* sess, int kind, const float* vmean,
                   float scale, void* o, long long ld_o, cudaStream_t stream);
