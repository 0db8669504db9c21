"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of the GAMER decoder hot path (the parity oracle).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
this module; the product (`gamer_b200/`) never does.

Parity status: PINNED.  The reference ships no tests or golden vectors, so the pins are
outputs of the reference itself executed in the build container through `oracle/ref_shim.py`:
`oracle/make_golden.py` freezes them under `tests/golden/`, and `tests/test_oracle_vs_golden.py` /
`tests/test_oracle_vs_reference.py` check this restatement against them.

This is a functional restatement (weights dict in, tensors out), not a copy of the reference's module tree.
Each function cites the reference lines it follows (paths relative to /root/reference/SeqRec/).

Third-party arithmetic restated here (transformers==4.51.0, un-vendored; formulas read from the installed
5.5.0 copy, site-packages/transformers/models/qwen3_moe/modeling_qwen3_moe.py): RMSNorm (:290-308),
half-split RoPE (:56-86, :386-449), SDPA with an additive finfo.min mask, ForCausalLMLoss
(loss/loss_utils.py:28-68), and HF `_beam_search` + PrefixConstrainedLogitsProcessor (oracle_decode.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch
import torch.nn.functional as F

MASK_CAUSAL = 0          # Qwen3Multi self:            j<=i & am[j]                       (Qwen3Multi/model.py:691-741)
MASK_MULTI_CROSS = 1     # Qwen3Multi cross:           j<=i & act[j]<act[i] & am[j]        (Qwen3Multi/model.py:573-630)
MASK_SESSION = 2         # Qwen3SessionMoe/-Multi self: (item(j)==item(i)&j<=i | s[j]<s[i]) & am[j]
#                                                       (Qwen3SessionMoe/model.py:402-468)
MASK_SESSION_CROSS = 3   # Qwen3SessionMulti cross:    s[j]<s[i] & act[j]<act[i] & am[j]   (Qwen3SessionMulti/model.py:556-613)


@dataclass
class Spec:
    """Architecture constants (config/s2s-models/*/config.json + train_SMB_decoder.py:321-360)."""
    variant: str = "Qwen3Multi"            # Qwen3Multi | Qwen3SessionMoe | Qwen3SessionMulti | Qwen3Moe
    vocab_size: int = 1041
    hidden: int = 256
    n_q: int = 6
    n_kv: int = 3
    head_dim: int = 64
    inter: int = 512
    n_layers: int = 8
    beh_dim: int = 64
    n_behavior: int = 3
    n_positions: int = 5                   # tokens per item (behaviour token + 4 codes)
    n_experts: int = 6
    sparse_layers: tuple = (0, 1, 2, 3, 4, 5, 6, 7)
    inject_layers: tuple = (0, 1, 2, 3)    # FFN behaviour-embedding concat
    cross_layers: tuple = (4, 5, 6, 7)     # gated behaviour "cross" attention
    behavior_maps: dict = field(default_factory=lambda: {526: 0, 527: 1, 528: 2})
    pad: int = 4
    eos: int = 8
    eps: float = 1e-6
    rope_theta: float = 1e6
    temperature: float = 1.0

    @staticmethod
    def from_hf_config(cfg, variant: str, temperature: float = 1.0) -> "Spec":
        cross = tuple(getattr(cfg, "cross_attention_decoder", ()) or ()) \
            if variant not in ("Qwen3SessionMoe", "Qwen3Moe") else ()
        return Spec(
            variant=variant, vocab_size=cfg.vocab_size, hidden=cfg.hidden_size, n_q=cfg.num_attention_heads,
            n_kv=cfg.num_key_value_heads, head_dim=cfg.head_dim, inter=cfg.intermediate_size,
            n_layers=cfg.num_hidden_layers, beh_dim=cfg.behavior_embedding_dim, n_behavior=cfg.num_behavior,
            n_positions=cfg.num_positions, n_experts=cfg.num_experts,
            sparse_layers=tuple(cfg.sparse_layers_decoder), inject_layers=tuple(cfg.behavior_injection_decoder),
            cross_layers=cross, behavior_maps={int(k): int(v) for k, v in cfg.behavior_maps.items()},
            pad=cfg.pad_token_id, eos=cfg.eos_token_id, eps=cfg.rms_norm_eps, rope_theta=float(cfg.rope_theta)
            if hasattr(cfg, "rope_theta") else float(cfg.rope_parameters["rope_theta"]), temperature=temperature)


# --------------------------------------------------------------------------------------------------------------
# A1 router  (Qwen3Multi/router.py:74-201; two-output twin Qwen3Moe/router.py:74-154)
# --------------------------------------------------------------------------------------------------------------
def route(spec: Spec, ids: torch.Tensor, positions: torch.Tensor, context_ids: torch.Tensor | None = None):
    """ids [B,S] are the tokens being processed, at absolute `positions` [S]; `context_ids` [B,>=max(pos)+1] is
    the whole sequence so far (== ids for a full forward).  Returns (position_index, behavior_index,
    action_index), int64 [B,S].

    position_index = (t mod P)+1, 0 at pad/eos (router.py:50-58,104).
    The item's behaviour token sits at P*(t//P); its mapped id is behavior_maps[tok]+1, unmapped tokens keep
    their raw id (router.py:122-123).  n_items = (max(pos)+P-1)//P (router.py:113-118); slot t == P*n_items is
    the "EOS" zero (router.py:127-137).  behavior_index is additionally zero on the behaviour slot itself
    (router.py:139); both are zero at pad/eos (router.py:144-147, 190-193).
    """
    P = spec.n_positions
    ctx = ids if context_ids is None else context_ids
    B, S = ids.shape
    positions = positions.to(torch.long)
    pos_index = (positions % P + 1).unsqueeze(0).repeat(B, 1)
    special = (ids == spec.pad) | (ids == spec.eos)
    pos_index = pos_index.masked_fill(special, 0)
    n_items = (int(positions.max()) + P - 1) // P
    item_start = (positions // P) * P                                   # [S]
    in_range = item_start < n_items * P
    beh_tok = ctx[:, item_start.clamp(max=ctx.shape[1] - 1)]             # [B,S]
    mapped = beh_tok.clone()
    for tok, idx in spec.behavior_maps.items():                          # sequential replacement, as the reference
        mapped = torch.where(mapped == tok, torch.full_like(mapped, idx + 1), mapped)
    mapped = torch.where(in_range.unsqueeze(0), mapped, torch.zeros_like(mapped))
    action = mapped.masked_fill(special, 0)
    behavior = mapped.masked_fill((positions % P == 0).unsqueeze(0), 0).masked_fill(special, 0)
    return pos_index, behavior, action


# --------------------------------------------------------------------------------------------------------------
# A4/A5 mask predicates, evaluated densely here (the CUDA kernels evaluate them per (i,j) on the fly)
# --------------------------------------------------------------------------------------------------------------
def allow_matrix(kind: int, am: torch.Tensor, actions: torch.Tensor | None, sessions: torch.Tensor | None,
                 P: int = 5) -> torch.Tensor:
    """bool [B,L,L]: True where query i may attend key j (full forward / prefill)."""
    B, L = am.shape
    i = torch.arange(L, device=am.device).view(1, L, 1)
    j = torch.arange(L, device=am.device).view(1, 1, L)
    key_ok = am.bool().view(B, 1, L)
    if kind == MASK_CAUSAL:
        allow = (j <= i).expand(B, L, L)
    elif kind == MASK_MULTI_CROSS:
        allow = (j <= i) & (actions.view(B, 1, L) < actions.view(B, L, 1))
    elif kind == MASK_SESSION:
        allow = ((j <= i) & (j // P == i // P)) | (sessions.view(B, 1, L) < sessions.view(B, L, 1))
    elif kind == MASK_SESSION_CROSS:
        allow = (sessions.view(B, 1, L) < sessions.view(B, L, 1)) & (actions.view(B, 1, L) < actions.view(B, L, 1))
    else:
        raise ValueError(kind)
    return allow & key_ok


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """Qwen3RMSNorm: w * (x * rsqrt(mean(x^2)+eps)), statistics in fp32."""
    xf = x.float()
    return w * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).to(x.dtype)


def rope_cos_sin(spec: Spec, position_ids: torch.Tensor):
    """Qwen3RotaryEmbedding.forward: inv_freq_i = theta^(-2i/d); emb = cat(f, f). position_ids [B or 1, S]."""
    d = spec.head_dim
    inv = 1.0 / (spec.rope_theta ** (torch.arange(0, d, 2, dtype=torch.float32, device=position_ids.device) / d))
    f = position_ids.float().unsqueeze(-1) * inv                        # [b,S,d/2]
    emb = torch.cat([f, f], dim=-1)
    return emb.cos(), emb.sin()


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x [B,H,S,d]; half-split convention: rotate_half(x) = cat(-x2, x1)."""
    d = x.shape[-1]
    rot = torch.cat([-x[..., d // 2:], x[..., : d // 2]], dim=-1)
    return x * cos.unsqueeze(1) + rot * sin.unsqueeze(1)


USE_SDPA = False     # bench.py's gpu_baseline switches the attention core to F.scaled_dot_product_attention


def masked_attention(q, k, v, allow, scale, zp=None):
    """q [B,Hq,S,d], k/v [B,Hkv,T,d], allow bool [B,S,T].  Additive finfo(fp32).min mask then softmax, exactly
    what SDPA computes for the reference: a row with no allowed key is *uniform over all T keys* (quirk Q1).
    `zp` [B,Hq,S,T] (training): multiplicative dropout mask on the probabilities (SDPA dropout_p,
    Qwen3Multi/model.py:139).  Documented deviation of the CUDA path, restated here: uniform rows are not dropped
    (they keep their expectation, the mean of V)."""
    g = q.shape[1] // k.shape[1]
    k = k.repeat_interleave(g, dim=1)
    v = v.repeat_interleave(g, dim=1)
    bias = torch.where(allow, 0.0, torch.finfo(torch.float32).min).unsqueeze(1)
    if USE_SDPA and zp is None:
        # the reference's default branch (sdpa_attention_forward with the materialised additive mask,
        # Qwen3Multi/model.py:123-143): used by bench.py's GPU baseline, which times stock PyTorch on the same device
        return F.scaled_dot_product_attention(q, k, v, attn_mask=bias.to(q.dtype).expand(-1, q.shape[1], -1, -1),
                                              dropout_p=0.0, scale=scale)
    s = torch.matmul(q, k.transpose(-1, -2)) * scale + bias
    p = torch.softmax(s, dim=-1)
    if zp is not None:
        p = torch.where(allow.any(-1).view(allow.shape[0], 1, -1, 1), p * zp, p)
    return torch.matmul(p, v)


# --------------------------------------------------------------------------------------------------------------
# A6 attention block, A7 routed FFN, A8 layer
# --------------------------------------------------------------------------------------------------------------
def attention_block(spec, W, pre, h, cos, sin, allow, action_index, is_cross, cache=None, zp=None):
    """Qwen3MultiAttention.forward (Qwen3Multi/model.py:75-150); with is_cross=False and no behaviour terms it is
    also the third-party Qwen3MoeAttention used by Qwen3SessionMoe.  `cache` = dict(k=,v=) of earlier keys."""
    B, S, _ = h.shape
    d = spec.head_dim
    q = F.linear(h, W[pre + "q_proj.weight"]).view(B, S, spec.n_q, d)
    k = F.linear(h, W[pre + "k_proj.weight"]).view(B, S, spec.n_kv, d)
    v = F.linear(h, W[pre + "v_proj.weight"]).view(B, S, spec.n_kv, d)
    if is_cross:
        q = q + F.embedding(action_index, W[pre + "q_behavior_embedding.weight"]).view(B, S, spec.n_q, d)
        k = k + F.embedding(action_index, W[pre + "k_behavior_embedding.weight"]).view(B, S, spec.n_kv, d)
        v = v + F.embedding(action_index, W[pre + "v_behavior_embedding.weight"]).view(B, S, spec.n_kv, d)
    q = rmsnorm(q, W[pre + "q_norm.weight"], spec.eps).transpose(1, 2)
    k = rmsnorm(k, W[pre + "k_norm.weight"], spec.eps).transpose(1, 2)
    v = v.transpose(1, 2)
    q = apply_rope(q, cos, sin)
    k = apply_rope(k, cos, sin)
    if cache is not None:
        if "k" in cache:
            k = torch.cat([cache["k"], k], dim=2)
            v = torch.cat([cache["v"], v], dim=2)
        cache["k"], cache["v"] = k, v
    a = masked_attention(q, k, v, allow, d ** -0.5, zp)
    a = a.transpose(1, 2).reshape(B, S, spec.n_q * d)
    out = F.linear(a, W[pre + "o_proj.weight"])
    if is_cross:
        out = out * F.silu(F.linear(h, W[pre + "gating.weight"]))
    return out


def routed_ffn(spec, W, pre, h, position_index, behavior_index, inject, sparse, zi=None):
    """MyQwen3SparseMLP.forward (Qwen3Moe/FFN.py:53-72): expert = position_index (hard routing);
    expert(x) = down(dropout(silu(gate x) * up x)) (FFN.py:25-27); `zi` [B,S,I] = that dropout's mask (training)."""
    x = h
    if inject:
        x = torch.cat([x, F.embedding(behavior_index, W[pre + "behavior_embedding.weight"])], dim=-1)

    def expert(p, t, z):
        a = F.silu(F.linear(t, W[p + "gate_proj.weight"])) * F.linear(t, W[p + "up_proj.weight"])
        if z is not None:
            a = a * z
        return F.linear(a, W[p + "down_proj.weight"])

    if not sparse:
        return expert(pre + "mlp.", x, zi)
    out = torch.zeros_like(h)
    for e in range(spec.n_experts):
        sel = position_index == e
        if sel.any():
            out[sel] = expert(f"{pre}experts.expert_{e}.", x[sel], None if zi is None else zi[sel]).to(out.dtype)
    return out


def backbone(spec: Spec, W: dict, ids, am, positions, rope_pos, self_allow, cross_allow, context_ids=None,
             caches=None, drop=None):
    """Embedding -> router -> layers -> final norm.  (Qwen3MultiModel.forward, Qwen3Multi/model.py:744-880;
    Qwen3SessionMoeModel.forward, Qwen3SessionMoe/model.py:471-587.)  `drop` (oracle/dropout_masks.OracleDropout):
    training-mode dropout masks — nn.Dropout on each residual branch (Qwen3Multi/model.py:217,235,241), inside the
    experts (Qwen3Moe/FFN.py:26) and on the attention probabilities (:139)."""
    from oracle import dropout_masks as dm
    B, S = ids.shape

    def zh(l, kind, width):
        return None if drop is None else drop.hidden(l, kind, B, S, width)

    def zp(l, kind):
        return None if drop is None else drop.attn(l, kind, B, spec.n_q, S)

    def dropped(branch, z):
        return branch if z is None else branch * z

    x = F.embedding(ids, W["model.embed_tokens.weight"], padding_idx=spec.pad)   # padding_idx=4: Q10 (model.py:263)
    pos_idx, beh_idx, act_idx = route(spec, ids, positions, context_ids)
    cos, sin = rope_cos_sin(spec, rope_pos)
    for l in range(spec.n_layers):
        p = f"model.layers.{l}."
        c = caches[l] if caches is not None else None
        h = rmsnorm(x, W[p + "input_layernorm.weight"], spec.eps)
        x = x + dropped(attention_block(spec, W, p + "self_attn.", h, cos, sin, self_allow, None, False,
                                        None if c is None else c["self"], zp(l, dm.SITE_SELF_P)),
                        zh(l, dm.SITE_SELF_OUT, spec.hidden))
        if l in spec.cross_layers:
            h = rmsnorm(x, W[p + "post_self_attention_layernorm.weight"], spec.eps)
            x = x + dropped(attention_block(spec, W, p + "cross_attn.", h, cos, sin, cross_allow, act_idx, True,
                                            None if c is None else c["cross"], zp(l, dm.SITE_CROSS_P)),
                            zh(l, dm.SITE_CROSS_OUT, spec.hidden))
        post = "post_attention_layernorm.weight" if spec.variant in ("Qwen3SessionMoe", "Qwen3Moe") else \
            "post_cross_attention_layernorm.weight"
        h = rmsnorm(x, W[p + post], spec.eps)
        x = x + dropped(routed_ffn(spec, W, p + "mlp.", h, pos_idx, beh_idx, l in spec.inject_layers,
                                   l in spec.sparse_layers, zh(l, dm.SITE_FFN_INNER, spec.inter)),
                        zh(l, dm.SITE_FFN_OUT, spec.hidden))
    return rmsnorm(x, W["model.norm.weight"], spec.eps), (pos_idx, beh_idx, act_idx)


def mask_kinds(spec: Spec):
    if spec.variant == "Qwen3Multi":
        return MASK_CAUSAL, MASK_MULTI_CROSS
    if spec.variant == "Qwen3SessionMoe":
        return MASK_SESSION, None
    if spec.variant == "Qwen3Moe":            # train_MB_decoder backbone: HF causal + padding mask (Qwen3Moe/model.py:306-461)
        return MASK_CAUSAL, None
    if spec.variant == "Qwen3SessionMulti":
        return MASK_SESSION, MASK_SESSION_CROSS
    raise ValueError(spec.variant)


def forward(spec: Spec, W: dict, input_ids, attention_mask, labels=None, session_ids=None,
            extended_session_ids=None, actions=None, num_items_in_batch=None, return_hidden=False, drop=None):
    """Full (uncached) forward = Qwen3MultiWithTemperature.forward (Qwen3Multi/model.py:928-1013) /
    Qwen3SessionMoeWithTemperature.forward (Qwen3SessionMoe/model.py:633-735).

    Returns dict(logits, loss, hidden, route).  With labels, logits are the temperature-scaled ones (the
    reference divides in place, quirk Q7) and loss follows ForCausalLMLoss: shift by one, fp32 CE with
    ignore_index=-100, mean — or sum / num_items_in_batch when given.
    """
    B, L = input_ids.shape
    positions = torch.arange(L, device=input_ids.device)
    k_self, k_cross = mask_kinds(spec)
    self_allow = allow_matrix(k_self, attention_mask, actions, session_ids, spec.n_positions)
    cross_allow = allow_matrix(k_cross, attention_mask, actions, session_ids, spec.n_positions) \
        if (k_cross is not None and spec.cross_layers) else None
    if spec.variant in ("Qwen3SessionMoe", "Qwen3SessionMulti") and extended_session_ids is not None:
        rope_pos = extended_session_ids                                   # Qwen3SessionMoe/model.py:688-703
    else:
        rope_pos = positions.unsqueeze(0)                                  # Qwen3Multi/model.py:787-794
    hidden, routes = backbone(spec, W, input_ids, attention_mask, positions, rope_pos, self_allow, cross_allow,
                              drop=drop)
    logits = F.linear(hidden, W["lm_head.weight"])
    out = {"hidden": hidden, "route": routes, "loss": None}
    if labels is not None:
        logits = logits / spec.temperature
        out["loss"] = causal_lm_loss(logits, labels, num_items_in_batch)
    out["logits"] = logits
    return out


def causal_lm_loss(logits, labels, num_items_in_batch=None):
    """ForCausalLMLoss: logits.float(); labels padded with -100 and shifted left by one; CE(ignore=-100)."""
    V = logits.shape[-1]
    shift = F.pad(labels, (0, 1), value=-100)[..., 1:].contiguous()
    if num_items_in_batch is None:
        return F.cross_entropy(logits.float().view(-1, V), shift.view(-1), ignore_index=-100, reduction="mean")
    return F.cross_entropy(logits.float().view(-1, V), shift.view(-1), ignore_index=-100,
                           reduction="sum") / num_items_in_batch


# --------------------------------------------------------------------------------------------------------------
# cached decode: prefill + single-token steps (the path HF generate drives; SURVEY.md §8 A4/A5/A6 decode notes)
# --------------------------------------------------------------------------------------------------------------
class DecodeState:
    """Per-beam-row state.  Everything that the reference keeps on `self` (cross cache, last cross-mask row,
    router id cache — quirk Q3) lives here and IS reordered with the beams (Q3 waived)."""

    def __init__(self, spec: Spec, n_layers: int):
        self.caches = [{"self": {}, "cross": {}} for _ in range(n_layers)]
        self.am = None            # [R, T] attention mask incl. generated columns
        self.cross_row = None     # [R, T] bool: allowed keys of the last prompt row; generated columns False
        self.seq = None           # [R, T] all token ids so far
        self.rope_next = None     # [R] next RoPE position for session variants

    def reorder(self, idx: torch.Tensor):
        for c in self.caches:
            for kind in ("self", "cross"):
                for name in ("k", "v"):
                    if name in c[kind]:
                        c[kind][name] = c[kind][name].index_select(0, idx)
        self.am = self.am.index_select(0, idx)
        self.cross_row = self.cross_row.index_select(0, idx)
        self.seq = self.seq.index_select(0, idx)
        if self.rope_next is not None:
            self.rope_next = self.rope_next.index_select(0, idx)


def prefill(spec: Spec, W: dict, input_ids, attention_mask, session_ids=None, extended_session_ids=None,
            actions=None):
    """Prompt pass with caches; returns (last-position logits [R,V] fp32, DecodeState)."""
    R, L = input_ids.shape
    st = DecodeState(spec, spec.n_layers)
    positions = torch.arange(L)
    k_self, k_cross = mask_kinds(spec)
    self_allow = allow_matrix(k_self, attention_mask, actions, session_ids, spec.n_positions)
    cross_allow = allow_matrix(k_cross, attention_mask, actions, session_ids, spec.n_positions) \
        if (k_cross is not None and spec.cross_layers) else None
    session_rope = spec.variant in ("Qwen3SessionMoe", "Qwen3SessionMulti") and extended_session_ids is not None
    rope_pos = extended_session_ids if session_rope else positions.unsqueeze(0)
    hidden, _ = backbone(spec, W, input_ids, attention_mask, positions, rope_pos, self_allow, cross_allow,
                         caches=st.caches)
    st.am = attention_mask.clone()
    st.seq = input_ids.clone()
    st.cross_row = cross_allow[:, -1, :].clone() if cross_allow is not None else torch.zeros(R, L, dtype=torch.bool)
    if session_rope:
        st.rope_next = extended_session_ids.max(dim=-1)[0] + 1            # Qwen3SessionMoe/model.py:688-701
    return F.linear(hidden[:, -1], W["lm_head.weight"]).float(), st


def decode_step(spec: Spec, W: dict, st: DecodeState, new_ids: torch.Tensor):
    """One cached step on tokens new_ids [R] appended at position T.  Self attention sees every cached key with
    am[j] (Qwen3Multi/model.py:717-728); cross attention sees the last prompt row's keys only — generated
    columns are masked (:605-617) — and, when that set is empty, averages all T+1 cached values (Q1)."""
    R = new_ids.shape[0]
    T = st.seq.shape[1]
    st.seq = torch.cat([st.seq, new_ids.view(R, 1)], dim=1)
    st.am = torch.cat([st.am, torch.ones(R, 1, dtype=st.am.dtype)], dim=1)
    st.cross_row = torch.cat([st.cross_row, torch.zeros(R, 1, dtype=torch.bool)], dim=1)
    positions = torch.tensor([T])
    self_allow = st.am.bool().view(R, 1, T + 1)
    cross_allow = (st.cross_row & st.am.bool()).view(R, 1, T + 1)
    if st.rope_next is not None:
        rope_pos = st.rope_next.view(R, 1)
        st.rope_next = st.rope_next + 1
    else:
        rope_pos = positions.unsqueeze(0)
    hidden, _ = backbone(spec, W, new_ids.view(R, 1), None, positions, rope_pos, self_allow, cross_allow,
                         context_ids=st.seq, caches=st.caches)
    return F.linear(hidden[:, -1], W["lm_head.weight"]).float()
