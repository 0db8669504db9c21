// Input pipeline on the device: one launch turns a set of users of the pre-tokenised interaction
// store into the six [B, L] tensors the reference's collators emit
// (SeqRec/datasets/collator.py:47-107 DecoderOnlyCollator, :149-207 DecoderOnlyTestCollator;
//  SeqRec/datasets/SMB_dataset.py:194-234 session / extended-session / action arrays;
//  SeqRec/tasks/test_SMB_decoder.py:105-117 target-behaviour column).
// A warp owns a row: lanes walk its items 32 at a time (window position, session-change rank by ballot prefix), then the
// 160 tokens of the 32 items are written as contiguous int64 runs (lane = token, item fields fetched by shuffle).
#include "common.cuh"

namespace {

constexpr int TPI = 5;  // tokens per item: behaviour token + 4 semantic-code tokens

struct CollateArgs {
    const int* item_tokens;      // [T, 4]
    const short* behavior;       // [T]
    const int* session;          // [T]
    const long long* offsets;    // [N + 1]
    const long long* users;      // [B]
    int n_users, n_max, width, left_pad;
    const long long* beh_tokens; // [n_beh]
    const long long* beh_level;  // [n_beh]
    int n_beh;
    long long pad;
    int target_behavior;         // >= 0: append the target-behaviour column (evaluation)
    long long *ids, *mask, *labels, *sess, *ext, *act;
};

__global__ void __launch_bounds__(128) collate_kernel(CollateArgs a) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= a.n_users) return;
    const long long u = a.users[row];
    const long long start = a.offsets[u], end = a.offsets[u + 1];
    const int have = (int)(end - start);
    const int n = have < a.n_max ? have : a.n_max;       // items kept (the last n)
    const int L = TPI * a.width + (a.target_behavior >= 0 ? 1 : 0);
    const long long base = (long long)row * L;
    int rank_base = 0;            // session rank of the previous chunk's last valid item
    int prev_sess = 0;            // ... and its session id
    bool prev_valid = false;
    long long sess_max = 0;
    for (int c0 = 0; c0 < a.width; c0 += 32) {
        const int t = c0 + lane;
        const bool in_row = t < a.width;
        const bool valid = in_row && (a.left_pad ? t >= a.width - n : t < n);
        const long long idx = a.left_pad ? end - a.width + t : end - n + t;
        int4 tok = make_int4(0, 0, 0, 0);
        int beh = 0, se = 0;
        if (valid) {
            tok = *reinterpret_cast<const int4*>(a.item_tokens + 4 * idx);
            beh = a.behavior[idx];
            se = a.session[idx];
        }
        // session rank inside the window: +1 at every change of session id between two valid neighbours
        int left_se = __shfl_up_sync(0xffffffffu, se, 1);
        int left_ok = __shfl_up_sync(0xffffffffu, (int)valid, 1);
        if (lane == 0) {
            left_se = prev_sess;
            left_ok = prev_valid;
        }
        const bool change = valid && left_ok && se != left_se;
        const unsigned cm = __ballot_sync(0xffffffffu, change);
        const int rank = rank_base + __popc(cm & (0xffffffffu >> (31 - lane)));
        const long long beh_tok = valid ? a.beh_tokens[beh] : a.pad;
        const long long level = valid ? a.beh_level[beh] : 100;
        if (valid && se > sess_max) sess_max = se;
        rank_base += __popc(cm);
        prev_sess = __shfl_sync(0xffffffffu, se, 31);
        prev_valid = __shfl_sync(0xffffffffu, (int)valid, 31) != 0;
        // 32 items = 160 tokens: five passes of 32 consecutive tokens
        const int n_tok = (min(a.width - c0, 32)) * TPI;
#pragma unroll
        for (int p = 0; p < TPI; ++p) {
            const int j = p * 32 + lane;
            const int it = j / TPI, slot = j - it * TPI;
            const int v = __shfl_sync(0xffffffffu, (int)valid, it);
            const long long bt = __shfl_sync(0xffffffffu, beh_tok, it);
            const int t0 = __shfl_sync(0xffffffffu, tok.x, it), t1 = __shfl_sync(0xffffffffu, tok.y, it);
            const int t2 = __shfl_sync(0xffffffffu, tok.z, it), t3 = __shfl_sync(0xffffffffu, tok.w, it);
            const int s = __shfl_sync(0xffffffffu, se, it);
            const int r = __shfl_sync(0xffffffffu, rank, it);
            const long long lv = __shfl_sync(0xffffffffu, level, it);
            if (j < n_tok) {
                const long long o = base + (long long)c0 * TPI + j;
                const long long id = slot == 0 ? bt : (long long)(slot == 1 ? t0 : (slot == 2 ? t1 : (slot == 3 ? t2 : t3)));
                a.ids[o] = v ? id : a.pad;
                a.mask[o] = v ? 1 : 0;
                a.sess[o] = v ? (long long)s : 0;
                a.ext[o] = v ? (long long)r * TPI + slot : 0;
                a.act[o] = v ? lv : 100;
                if (a.labels) {            // collator.py:68-78: pad and behaviour tokens are not predicted
                    const long long x = v ? id : a.pad;
                    bool ignore = x == a.pad;
                    for (int k = 0; k < a.n_beh; ++k) ignore |= x == a.beh_tokens[k];
                    a.labels[o] = ignore ? -100 : x;
                }
            }
        }
    }
    if (a.target_behavior >= 0) {
        // one more column: the target behaviour token in a session of its own (max + 1 of what the row holds)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const long long other = __shfl_xor_sync(0xffffffffu, sess_max, o);
            sess_max = other > sess_max ? other : sess_max;
        }
        if (lane == 0) {
            const long long o = base + (long long)TPI * a.width;
            const long long ext_max = n > 0 ? (long long)rank_base * TPI + (TPI - 1) : 0;
            a.ids[o] = a.beh_tokens[a.target_behavior];
            a.mask[o] = 1;
            a.sess[o] = sess_max + 1;
            a.ext[o] = ext_max + 1;
            a.act[o] = a.beh_level[a.target_behavior];
            if (a.labels) a.labels[o] = -100;
        }
    }
}

}  // namespace

extern "C" int gamer_collate_sessions(const int* item_tokens, const short* behavior, const int* session,
                                      const long long* offsets, const long long* users, int n_users, int n_max, int width,
                                      int left_pad, const long long* beh_tokens, const long long* beh_level, int n_beh,
                                      long long pad, int target_behavior, long long* input_ids, long long* attention_mask,
                                      long long* labels, long long* session_ids, long long* extended_session_ids,
                                      long long* actions, cudaStream_t stream) {
    if (n_users == 0) return 0;
    GAMER_REQUIRE(n_users > 0 && width > 0 && n_max > 0, "bad collate shape (users %d, width %d, n_max %d)", n_users, width, n_max);
    GAMER_REQUIRE(n_beh > 0 && target_behavior < n_beh, "target behaviour %d out of range (%d behaviours)", target_behavior, n_beh);
    GAMER_REQUIRE((reinterpret_cast<uintptr_t>(item_tokens) & 15) == 0, "item_tokens must be 16-byte aligned");
    CollateArgs a{item_tokens, behavior, session, offsets, users, n_users, n_max, width, left_pad, beh_tokens, beh_level,
                  n_beh, pad, target_behavior, input_ids, attention_mask, labels, session_ids, extended_session_ids, actions};
    const int wpb = 4;
    collate_kernel<<<(n_users + wpb - 1) / wpb, wpb * 32, 0, stream>>>(a);
    GAMER_LAUNCH_CHECK();
    return 0;
}
