"""Thin functional wrappers: torch tensors in, C-ABI calls (include/gamer_b200.h) on the current CUDA stream, torch
tensors out.  PyTorch is used for device memory and streams only; all arithmetic happens in libgamer_b200.so.
"""
from __future__ import annotations

import ctypes

import torch

from ._cabi import call, lib, ptr

BF16 = torch.bfloat16
MASK_CAUSAL, MASK_MULTI_CROSS, MASK_SESSION, MASK_SESSION_CROSS = 0, 1, 2, 3


def _stream():
    return torch.cuda.current_stream().cuda_stream


class Dropout(ctypes.Structure):
    """gamer_dropout_t (include/gamer_b200.h): one dropout call site of a forward pass; the backward passes the same
    struct and the kernels regenerate the Philox mask."""
    _fields_ = [("seed", ctypes.c_ulonglong), ("offset", ctypes.c_uint), ("site", ctypes.c_uint), ("p", ctypes.c_float),
                ("offset_dev", ctypes.c_void_p)]


def _drop(d):
    return None if d is None else ctypes.addressof(d)


def _req(t: torch.Tensor, dtype=None):
    if not t.is_cuda:
        raise RuntimeError("gamer_b200 kernels need CUDA tensors (no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    return t


# ------------------------------------------------------------------------------------------------ K1
_ID_ERR: dict = {}


def _dev_key(device):
    d = torch.device(device)
    return d.index if d.index is not None else torch.cuda.current_device()


def id_error_word(device):
    """Per-device int32 word the gather kernel flags out-of-range token ids in (persistent: captured graphs hold it)."""
    key = _dev_key(device)
    w = _ID_ERR.get(key)
    if w is None:
        w = torch.zeros(1, dtype=torch.int32, device=device)
        if torch.cuda.is_current_stream_capturing():
            return w
        _ID_ERR[key] = w
    return w


def check_token_ids(device):
    """Raise what nn.Embedding raises if any token id since the last check was outside the vocabulary (one host sync)."""
    w = _ID_ERR.get(_dev_key(device))
    if w is not None and int(w.item()) != 0:
        w.zero_()
        raise IndexError("index out of range in self: a token id passed to the model is outside [0, vocab_size)")


def embed_route(ids, table_bf16, beh_lut, n_beh, tokens_per_item, pad, eos, ctx=None, pos0=0, want_x=True):
    """ids int64 [B,S] -> (x bf16 [B*S,H] | None, pos_idx, beh_idx, act_idx int32 [B*S])."""
    _req(ids, torch.int64)
    ids = ids.contiguous()
    B, S = ids.shape
    V, H = table_bf16.shape
    dev = ids.device
    x = torch.empty(B * S, H, dtype=BF16, device=dev) if want_x else None
    pos = torch.empty(B * S, dtype=torch.int32, device=dev)
    beh = torch.empty_like(pos)
    act = torch.empty_like(pos)
    if ctx is not None:
        ctx = ctx.contiguous()
    call("gamer_embed_route_fwd", ptr(ids), ptr(ctx), 0 if ctx is None else ctx.shape[1], B, S, pos0, tokens_per_item,
         pad, eos, V, ptr(beh_lut), n_beh, ptr(table_bf16), H, ptr(x), ptr(pos), ptr(beh), ptr(act),
         ptr(id_error_word(dev)), _stream(), work=(0, B * S * (8 + 2 * H + 12)))        # id + bf16 row + 3 int32 indices per token (§8d)
    return x, pos, beh, act


def route_perm(pos_idx, B, S, n_experts):
    """-> perm [M] (token row -> permuted row), rows [M + 128*E] (permuted row -> token row | -1), seg_off [E+1]."""
    dev = pos_idx.device
    M = B * S
    cap = M + 128 * n_experts
    ws = torch.empty(lib().gamer_route_perm_workspace_bytes(B), dtype=torch.uint8, device=dev)
    perm = torch.empty(M, dtype=torch.int32, device=dev)
    rows = torch.empty(cap, dtype=torch.int32, device=dev)
    seg = torch.empty(n_experts + 1, dtype=torch.int32, device=dev)
    call("gamer_route_perm_build", ptr(pos_idx), B, S, n_experts, ptr(ws), ptr(perm), ptr(rows), cap, ptr(seg), _stream())
    return perm, rows, seg


def embed_sort(ids, vocab, pad):
    ids = ids.contiguous()
    M = ids.numel()
    buf = torch.empty(lib().gamer_embed_sort_bytes(M, vocab), dtype=torch.uint8, device=ids.device)
    call("gamer_embed_sort_build", ptr(ids), M, vocab, pad, ptr(buf), _stream())
    return buf


def embed_bwd(dx, vocab, sort_buf, dtable=None):
    M, H = dx.shape
    if dtable is None:
        dtable = torch.zeros(vocab, H, dtype=torch.float32, device=dx.device)
    call("gamer_embed_bwd", ptr(dx), M, H, vocab, ptr(sort_buf), ptr(dtable), _stream(),
         work=(0, M * (4 + 2 * H) + vocab * H * 4))
    return dtable


# ------------------------------------------------------------------------------------------------ K2/K3
def rmsnorm_fwd(x, w, eps, out=None, ld_out=None, row_map=None, cat_table=None, cat_idx=None, out_rows=None):
    M, H = x.shape
    cat_dim = 0 if cat_table is None else cat_table.shape[1]
    if out is None:
        out = torch.empty(out_rows if out_rows is not None else M, H + cat_dim, dtype=BF16, device=x.device)
    rstd = torch.empty(M, dtype=torch.float32, device=x.device)
    call("gamer_rmsnorm_fwd", ptr(x), ptr(w), eps, M, H, ptr(out), out.stride(0) if ld_out is None else ld_out,
         ptr(row_map), ptr(cat_table), ptr(cat_idx), cat_dim, ptr(rstd), _stream())
    return out, rstd


def rmsnorm_bwd(x, w, rstd, eps, dh, dw, row_map=None, dres=None, cat_idx=None, cat_dim=0, cat_rows=0, dcat=None):
    M, H = x.shape
    dx = torch.empty_like(x)
    call("gamer_rmsnorm_bwd", ptr(x), ptr(w), ptr(rstd), eps, M, H, ptr(dh), dh.stride(0), ptr(row_map), ptr(dres),
         ptr(dx), ptr(dw), ptr(cat_idx), cat_dim, cat_rows, ptr(dcat), _stream())
    return dx


def qk_norm_rope_fwd(raw, L, n_q, n_kv, hd, cos_tab, sin_tab, qn_w, kn_w, eps, pos_ids=None, pos0=0, q_emb=None,
                     k_emb=None, v_emb=None, act_idx=None, width=None, out=None):
    M = raw.shape[0]
    width = width if width is not None else (n_q + 2 * n_kv) * hd
    if out is None:
        out = torch.empty(M, width, dtype=BF16, device=raw.device)
    call("gamer_qk_norm_rope_fwd", ptr(raw), raw.stride(0), ptr(out), out.stride(0), M, L, n_q, n_kv, hd, ptr(pos_ids),
         pos0, cos_tab.shape[0], ptr(cos_tab), ptr(sin_tab), ptr(qn_w), ptr(kn_w), ptr(q_emb), ptr(k_emb), ptr(v_emb),
         ptr(act_idx), eps, _stream())
    return out


def qk_norm_rope_bwd(raw, dout, draw, L, n_q, n_kv, hd, cos_tab, sin_tab, qn_w, kn_w, eps, d_qn_w, d_kn_w, pos_ids=None,
                     pos0=0, q_emb=None, k_emb=None, v_emb=None, act_idx=None, emb_rows=0, d_q_emb=None, d_k_emb=None,
                     d_v_emb=None):
    M = raw.shape[0]
    call("gamer_qk_norm_rope_bwd", ptr(raw), raw.stride(0), ptr(dout), dout.stride(0), ptr(draw), draw.stride(0), M, L,
         n_q, n_kv, hd, ptr(pos_ids), pos0, cos_tab.shape[0], ptr(cos_tab), ptr(sin_tab), ptr(qn_w), ptr(kn_w), ptr(q_emb),
         ptr(k_emb), ptr(v_emb), ptr(act_idx), emb_rows, eps, ptr(d_qn_w), ptr(d_kn_w), ptr(d_q_emb), ptr(d_k_emb), ptr(d_v_emb),
         _stream())
    return draw


# ------------------------------------------------------------------------------------------------ K4/K7
def gemm_tn(a, b, N, K=None, rows=None, n_groups=1, seg_off=None, out=None, out_f32=False, resid=None, row_map=None,
            alpha=1.0, out_rows=None, drop=None):
    """C = dropout(alpha * A @ B_g^T) (+ resid).  a: bf16 [rows, >=K] (row stride a.stride(0)); b: bf16 [n_groups*N, >=K]."""
    rows = a.shape[0] if rows is None else rows
    K = a.shape[1] if K is None else K
    if out is None:
        out = torch.empty(out_rows if out_rows is not None else rows, N, dtype=torch.float32 if out_f32 else BF16,
                          device=a.device)
    call("gamer_gemm_bf16_tn", ptr(a), a.stride(0), rows, ptr(b), b.stride(0), n_groups, N, K, ptr(seg_off), ptr(out),
         out.stride(0), 1 if out.dtype == torch.float32 else 0, ptr(resid), 0 if resid is None else resid.stride(0),
         ptr(row_map), float(alpha), _drop(drop), _stream(),
         work=(2 * rows * N * K, rows * K * 2 + n_groups * N * K * 2 + rows * N * out.element_size()))
    return out


def gemm_wgrad(dy, x, N_out, K_in, dw, rows=None, n_groups=1, seg_off=None):
    """dw[g] += dy_g^T @ x_g   (dw fp32 [n_groups, N_out, K_in], contiguous)."""
    rows = dy.shape[0] if rows is None else rows
    assert dw.dtype == torch.float32 and dw.is_contiguous()
    call("gamer_gemm_bf16_wgrad", ptr(dy), dy.stride(0), ptr(x), x.stride(0), rows, N_out, K_in, n_groups, ptr(seg_off),
         ptr(dw), _stream(), work=(2 * rows * N_out * K_in, rows * (N_out + K_in) * 2 + n_groups * N_out * K_in * 4))
    return dw


def ref_gemm_tn(a, b, N, K):
    out = torch.empty(a.shape[0], N, dtype=torch.float32, device=a.device)
    call("gamer_ref_gemm_tn", ptr(a), a.stride(0), ptr(b), b.stride(0), ptr(out), out.stride(0), a.shape[0], N, K, _stream())
    return out


# ------------------------------------------------------------------------------------------------ K6
def attn_fwd(qkv, B, L, n_q, n_kv, hd, kind, P, am, act, sess, scale, drop=None):
    """qkv bf16 [B*L, ld]: q cols [0, n_q*hd), k next n_kv*hd, v next n_kv*hd.
    -> (o [B*L, n_q*hd], lse [B,n_q,L], vmean fp32 [B, n_kv, hd] = mean of all L value rows, keep words | None).
    `keep` (dropout on) holds the keep flags the backward of this call site reads back."""
    dev = qkv.device
    ws = torch.empty(lib().gamer_attn_workspace_bytes(B, L, n_q, n_kv), dtype=torch.uint8, device=dev)
    o = torch.empty(B * L, n_q * hd, dtype=BF16, device=dev)
    lse = torch.empty(B, n_q, L, dtype=torch.float32, device=dev)
    keep = None
    if drop is not None and drop.p > 0:
        keep = torch.empty(lib().gamer_attn_keep_bytes(B, L, n_q) // 4, dtype=torch.int32, device=dev)
    esz = 2
    q = qkv.data_ptr()
    k = q + n_q * hd * esz
    v = k + n_kv * hd * esz
    call("gamer_attn_fwd", q, k, v, qkv.stride(0), B, L, n_q, n_kv, hd, kind, P, ptr(am), ptr(act), ptr(sess),
         float(scale), ptr(ws), ptr(o), o.stride(0), ptr(lse), _drop(drop), ptr(keep), _stream(),
         work=(4 * hd * n_q * B * L * (L + 1) // 2, B * L * (2 * n_q + 2 * n_kv) * hd * 2))   # causal pair count (§8d)
    return o, lse, ws[: B * n_kv * hd * 4].view(torch.float32).view(B, n_kv, hd), keep


def attn_bwd(qkv, o, d_o, lse, B, L, n_q, n_kv, hd, kind, P, am, act, sess, scale, dqkv, drop=None, keep=None):
    dev = qkv.device
    ws = torch.empty(lib().gamer_attn_bwd_workspace_bytes(B, L, n_q), dtype=torch.uint8, device=dev)
    esz = 2
    q = qkv.data_ptr()
    k = q + n_q * hd * esz
    v = k + n_kv * hd * esz
    dq = dqkv.data_ptr()
    dk = dq + n_q * hd * esz
    dv = dk + n_kv * hd * esz
    call("gamer_attn_bwd", q, k, v, qkv.stride(0), B, L, n_q, n_kv, hd, kind, P, ptr(am), ptr(act), ptr(sess),
         float(scale), ptr(o), ptr(d_o), o.stride(0), ptr(lse), ptr(ws), dq, dk, dv, dqkv.stride(0), _drop(drop),
         ptr(keep), _stream(),
         work=(10 * hd * n_q * B * L * (L + 1) // 2, B * L * (4 * n_q + 4 * n_kv) * hd * 2))
    return dqkv


# ------------------------------------------------------------------------------------------------ elementwise
def swiglu_fwd(gu, I, rows=None, row_ids=None, drop=None):
    """act = dropout(silu(gate) * up); row_ids maps rows of the expert-permuted space to token rows (mask index)."""
    R = gu.shape[0] if rows is None else rows
    act = torch.empty(gu.shape[0], I, dtype=BF16, device=gu.device)
    call("gamer_swiglu_fwd", ptr(gu), gu.stride(0), ptr(act), act.stride(0), R, I, ptr(row_ids), _drop(drop), _stream())
    return act


def swiglu_bwd(gu, dact, I, row_ids=None, drop=None):
    dgu = torch.empty_like(gu)
    call("gamer_swiglu_bwd", ptr(gu), gu.stride(0), ptr(dact), dact.stride(0), ptr(dgu), dgu.stride(0), gu.shape[0], I,
         ptr(row_ids), _drop(drop), _stream())
    return dgu


def gate_residual_fwd(x, y, g_view, drop=None):
    """out = x + dropout(y * silu(g))."""
    out = torch.empty_like(x)
    call("gamer_gate_residual_fwd", ptr(x), ptr(y), ptr(g_view), g_view.stride(0), ptr(out), x.shape[0], x.shape[1],
         _drop(drop), _stream())
    return out


def gate_residual_bwd(dout, y, g_view, dg_view, drop=None):
    dy = torch.empty_like(dout)
    call("gamer_gate_residual_bwd", ptr(dout), ptr(y), ptr(g_view), g_view.stride(0), ptr(dy), ptr(dg_view),
         dg_view.stride(0), dout.shape[0], dout.shape[1], _drop(drop), _stream())
    return dy


def gather_rows(src, rows, n_rows_max, width=None, n_rows_dev=None, drop=None):
    W = src.shape[1] if width is None else width
    dst = torch.empty(n_rows_max, W, dtype=BF16, device=src.device)
    call("gamer_gather_rows", ptr(src), src.stride(0), ptr(rows), ptr(n_rows_dev), n_rows_max, ptr(dst), dst.stride(0),
         W, _drop(drop), _stream())
    return dst


def zero_unmapped_rows(dst, rows):
    """dst[r] = 0 for the rows r with rows[r] < 0 (padding rows of the permuted token space)."""
    call("gamer_zero_unmapped_rows", ptr(dst), dst.stride(0), ptr(rows), dst.shape[0], dst.shape[1], _stream())
    return dst


def dropout_apply(x, drop):
    """dropout(x) with the mask of `drop` (x contiguous [R, W]); the backward of a dropout fused into a GEMM epilogue."""
    out = torch.empty_like(x)
    call("gamer_dropout_apply", ptr(x), ptr(out), x.shape[0], x.shape[1], _drop(drop), _stream())
    return out


def ce_fwd_bwd(logits, labels, V, inv_norm, grad_scale, dlogits=None, ignore_index=-100):
    R = logits.shape[0]
    loss_row = torch.empty(R, dtype=torch.float32, device=logits.device)
    call("gamer_ce_fwd_bwd", ptr(logits), logits.stride(0), ptr(labels), R, V, ignore_index, ptr(inv_norm),
         float(grad_scale), ptr(loss_row), ptr(dlogits), 0 if dlogits is None else dlogits.stride(0), _stream())
    return loss_row


# ------------------------------------------------------------------------------------------------ K9 decode
def attn_decode(qcur, prompt_rot, gen, step_stride, anc, B, beams, L0, n_gen, n_q, n_kv, hd, S_max, am, act, sess, kind,
                vmean, scale):
    """qcur [R, ld_g] (this step's rotated q|k|v), prompt_rot [B*L0, ld_p], gen = base tensor of the per-step buffers."""
    R = B * beams
    o = torch.empty(R, n_q * hd, dtype=BF16, device=qcur.device)
    esz = 2
    koff, voff = n_q * hd * esz, (n_q + n_kv) * hd * esz
    call("gamer_attn_decode", ptr(qcur), prompt_rot.data_ptr() + koff, prompt_rot.data_ptr() + voff, prompt_rot.stride(0),
         gen.data_ptr() + koff, gen.data_ptr() + voff, step_stride, qcur.stride(0), ptr(anc), B, beams, L0, n_gen, n_q,
         n_kv, hd, S_max, ptr(am), ptr(act), ptr(sess), kind, ptr(vmean), float(scale), ptr(o), o.stride(0), _stream(),
         work=(4 * hd * n_q * R * (L0 + n_gen),                                  # QK^T + PV over every cached key
               B * L0 * 2 * n_kv * hd * esz + R * (n_q + 2 * n_kv * n_gen + n_q) * hd * esz))   # prompt K/V once per user
    return o


def trie_init(ids, vocab, last_set, flat):
    B, L = ids.shape
    node = torch.empty(B, dtype=torch.int32, device=ids.device)
    call("gamer_trie_init", ptr(ids), B, L, vocab, ptr(last_set), ptr(flat.child_start), ptr(flat.child_tok),
         ptr(flat.child_node), ptr(node), _stream())
    return node


def beam_step(logits, vocab, n_users, beams, run_score, node, flat, err):
    dev = logits.device
    new_score = torch.empty(n_users, beams, dtype=torch.float32, device=dev)
    new_parent = torch.empty(n_users, beams, dtype=torch.int32, device=dev)
    new_tok = torch.empty_like(new_parent)
    new_node = torch.empty_like(new_parent)
    call("gamer_beam_step", ptr(logits), logits.stride(0), vocab, n_users, beams, ptr(run_score), ptr(node),
         ptr(flat.child_start), ptr(flat.child_tok), ptr(flat.child_node), flat.max_children, ptr(new_score),
         ptr(new_parent), ptr(new_tok), ptr(new_node), ptr(err), _stream())
    return new_score, new_parent, new_tok, new_node


def transpose_batch(desc, tile_start, n_mats, total_tiles):
    """dst = src.t() for every (src, dst) bf16 matrix pair of the descriptor table, one launch."""
    call("gamer_transpose_bf16_batch", ptr(desc), ptr(tile_start), int(n_mats), int(total_tiles), _stream())


def collate_sessions(store, users, n_max, width, left_pad, beh_tokens, beh_level, pad, target_behavior, with_labels):
    """gamer_collate_sessions: the collators' batch tensors from the packed store (all on the device), one launch."""
    B = users.numel()
    L = 5 * width + (1 if target_behavior >= 0 else 0)
    dev = users.device
    names = ["input_ids", "attention_mask", "session_ids", "extended_session_ids", "actions"] + (["labels"] if with_labels else [])
    buf = torch.empty(len(names), B, L, dtype=torch.int64, device=dev)
    out = {n: buf[i] for i, n in enumerate(names)}
    toks = store.item_tokens
    assert toks.dtype == torch.int32 and toks.is_contiguous() and store.behavior.dtype == torch.int16
    assert store.session.dtype == torch.int32 and store.offsets.dtype == torch.int64 and users.dtype == torch.int64
    call("gamer_collate_sessions", ptr(toks), ptr(store.behavior), ptr(store.session), ptr(store.offsets), ptr(users), B,
         int(n_max), int(width), 1 if left_pad else 0, ptr(beh_tokens), ptr(beh_level), beh_tokens.numel(), int(pad),
         int(target_behavior), ptr(out["input_ids"]), ptr(out["attention_mask"]), ptr(out.get("labels")),
         ptr(out["session_ids"]), ptr(out["extended_session_ids"]), ptr(out["actions"]), _stream())
    return out

