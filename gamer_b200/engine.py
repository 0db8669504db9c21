"""Explicit forward/backward schedule of the GAMER decoder stack over the C-ABI kernels.

One `torch.autograd.Function` (`DecoderLossFunction`) spans embedding -> layers -> final norm -> lm_head -> loss, so the
backward is a hand-ordered kernel sequence (reverse layer order, weight gradients produced layer by layer — the order
the gradient buckets are all-reduced in, see gamer_b200/distributed.py) instead of an autograd tape of small ops.

Data layout (HBM): tokens are flattened to M = B*L rows; the residual stream and every activation are bf16 row-major
[M, width]; q|k|v(|gate) of one attention live in one fused projection buffer [M, 768 or 1024]; the routed FFN works in
an expert-permuted row space [Mp, .] whose segments are 128-row aligned (Mp <= M + 128*E).  Statistics (rstd, lse),
parameter gradients and the loss are fp32.

Reference call sites: Qwen3MultiModel.forward (SeqRec/models/generative/Qwen3Multi/model.py:744-880),
Qwen3MultiDecoderLayer.forward (:186-247), Qwen3SessionMoeDecoderLayer.forward (Qwen3SessionMoe/model.py:51-93),
MyQwen3SparseMLP.forward (Qwen3Moe/FFN.py:53-72), Qwen3MultiWithTemperature.forward/loss_function (:904-1013).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch

from . import kernels as K

BF16 = torch.bfloat16
# backbones without the gated behaviour "cross" attention (one self-attention + routed FFN per layer,
# `post_attention_layernorm`): Qwen3SessionMoe (session mask, session RoPE) and Qwen3Moe (train_MB_decoder: plain causal)
NO_CROSS_VARIANTS = ("Qwen3SessionMoe", "Qwen3Moe")
CE_CHUNK_ROWS = 32768
V_ALIGN = 64


@dataclass
class Arch:
    variant: str
    vocab: int
    hidden: int
    n_q: int
    n_kv: int
    head_dim: int
    inter: int
    n_layers: int
    beh_dim: int
    n_beh: int
    P: int
    n_exp: int
    sparse: tuple
    inject: tuple
    cross: tuple
    pad: int
    eos: int
    eps: float
    theta: float
    behavior_maps: dict = field(default_factory=dict)

    @staticmethod
    def from_config(cfg, variant: str) -> "Arch":
        cross = tuple(getattr(cfg, "cross_attention_decoder", ()) or ()) if variant not in NO_CROSS_VARIANTS else ()
        theta = getattr(cfg, "rope_theta", None)
        if theta is None:
            theta = (getattr(cfg, "rope_parameters", None) or {}).get("rope_theta", 1e6)
        return Arch(variant=variant, vocab=cfg.vocab_size, hidden=cfg.hidden_size, n_q=cfg.num_attention_heads,
                    n_kv=cfg.num_key_value_heads, head_dim=getattr(cfg, "head_dim", 64), inter=cfg.intermediate_size,
                    n_layers=cfg.num_hidden_layers, beh_dim=cfg.behavior_embedding_dim, n_beh=cfg.num_behavior,
                    P=cfg.num_positions, n_exp=cfg.num_experts, sparse=tuple(cfg.sparse_layers_decoder),
                    inject=tuple(cfg.behavior_injection_decoder), cross=cross, pad=cfg.pad_token_id,
                    eos=cfg.eos_token_id, eps=cfg.rms_norm_eps, theta=float(theta),
                    behavior_maps={int(k): int(v) for k, v in cfg.behavior_maps.items()})

    @property
    def qkv_w(self):
        return (self.n_q + 2 * self.n_kv) * self.head_dim

    @property
    def q_w(self):
        return self.n_q * self.head_dim

    @property
    def kv_w(self):
        return self.n_kv * self.head_dim

    @property
    def v_ld(self):
        return (self.vocab + V_ALIGN - 1) // V_ALIGN * V_ALIGN

    def mask_kinds(self):
        if self.variant == "Qwen3Multi":
            return K.MASK_CAUSAL, K.MASK_MULTI_CROSS
        if self.variant == "Qwen3SessionMoe":
            return K.MASK_SESSION, None
        if self.variant == "Qwen3Moe":              # HF causal + padding mask (Qwen3Moe/model.py:306-461)
            return K.MASK_CAUSAL, None
        if self.variant == "Qwen3SessionMulti":
            return K.MASK_SESSION, K.MASK_SESSION_CROSS
        raise ValueError(self.variant)

    def session_rope(self):
        return self.variant in ("Qwen3SessionMoe", "Qwen3SessionMulti")


def param_names(arch: Arch):
    """State-dict keys in the order the engine consumes them (SURVEY.md §8(b) checkpoint contract)."""
    names = ["model.embed_tokens.weight"]
    post = "post_attention_layernorm" if arch.variant in NO_CROSS_VARIANTS else "post_cross_attention_layernorm"
    for l in range(arch.n_layers):
        p = f"model.layers.{l}."
        names += [p + "input_layernorm.weight"]
        names += [p + f"self_attn.{n}.weight" for n in ("q_proj", "k_proj", "v_proj", "o_proj", "q_norm", "k_norm")]
        if l in arch.cross:
            names += [p + "post_self_attention_layernorm.weight"]
            names += [p + f"cross_attn.{n}.weight" for n in
                      ("q_proj", "k_proj", "v_proj", "o_proj", "q_norm", "k_norm", "q_behavior_embedding",
                       "k_behavior_embedding", "v_behavior_embedding", "gating")]
        names += [p + post + ".weight"]
        if l in arch.sparse:
            for e in range(arch.n_exp):
                names += [p + f"mlp.experts.expert_{e}.{n}.weight" for n in ("gate_proj", "up_proj", "down_proj")]
        else:
            names += [p + f"mlp.mlp.{n}.weight" for n in ("gate_proj", "up_proj", "down_proj")]
        if l in arch.inject:
            names += [p + "mlp.behavior_embedding.weight"]
    names += ["model.norm.weight"]
    return names


class Pack:
    """bf16 operand copies of the fp32 master weights, fused/stacked/transposed the way the GEMM kernels read them.
    Rebuilt once per optimizer step (24.5 M parameters: a few tens of microseconds of HBM traffic)."""

    def __init__(self, arch: Arch, W: dict):
        a = arch
        bf = lambda t: t.detach().to(BF16)
        self.emb = bf(W["model.embed_tokens.weight"]).contiguous()                     # [V, H]  (also lm_head, tied)
        emb_t = torch.zeros(a.hidden, a.v_ld, dtype=BF16, device=self.emb.device)      # [H, v_ld] for the lm_head dgrad
        emb_t[:, :a.vocab] = self.emb.t()
        self.emb_t = emb_t
        self.norm = W["model.norm.weight"].detach().float().contiguous()
        self.layers = []
        post = "post_attention_layernorm" if a.variant in NO_CROSS_VARIANTS else "post_cross_attention_layernorm"
        for l in range(a.n_layers):
            p = f"model.layers.{l}."
            d = {}
            d["in_norm"] = W[p + "input_layernorm.weight"].detach().float().contiguous()
            sa = p + "self_attn."
            d["w_qkv"] = torch.cat([bf(W[sa + "q_proj.weight"]), bf(W[sa + "k_proj.weight"]), bf(W[sa + "v_proj.weight"])], 0).contiguous()
            d["w_qkv_t"] = d["w_qkv"].t().contiguous()
            d["w_o"] = bf(W[sa + "o_proj.weight"]).contiguous()
            d["w_o_t"] = d["w_o"].t().contiguous()
            d["qn"] = W[sa + "q_norm.weight"].detach().float().contiguous()
            d["kn"] = W[sa + "k_norm.weight"].detach().float().contiguous()
            if l in a.cross:
                ca = p + "cross_attn."
                d["ps_norm"] = W[p + "post_self_attention_layernorm.weight"].detach().float().contiguous()
                d["c_w_qkvg"] = torch.cat([bf(W[ca + "q_proj.weight"]), bf(W[ca + "k_proj.weight"]),
                                           bf(W[ca + "v_proj.weight"]), bf(W[ca + "gating.weight"])], 0).contiguous()
                d["c_w_qkvg_t"] = d["c_w_qkvg"].t().contiguous()
                d["c_w_o"] = bf(W[ca + "o_proj.weight"]).contiguous()
                d["c_w_o_t"] = d["c_w_o"].t().contiguous()
                d["c_qn"] = W[ca + "q_norm.weight"].detach().float().contiguous()
                d["c_kn"] = W[ca + "k_norm.weight"].detach().float().contiguous()
                d["c_qe"] = bf(W[ca + "q_behavior_embedding.weight"]).contiguous()
                d["c_ke"] = bf(W[ca + "k_behavior_embedding.weight"]).contiguous()
                d["c_ve"] = bf(W[ca + "v_behavior_embedding.weight"]).contiguous()
            d["post_norm"] = W[p + post + ".weight"].detach().float().contiguous()
            if l in a.sparse:
                prefixes = [p + f"mlp.experts.expert_{e}." for e in range(a.n_exp)]
            else:
                prefixes = [p + "mlp.mlp."]
            gu = [torch.cat([bf(W[q + "gate_proj.weight"]), bf(W[q + "up_proj.weight"])], 0) for q in prefixes]
            d["w_gu"] = torch.cat(gu, 0).contiguous()                                   # [E*2I, Kf]
            d["w_gu_t"] = torch.cat([g.t() for g in gu], 0).contiguous()               # [E*Kf, 2I]
            dn = [bf(W[q + "down_proj.weight"]) for q in prefixes]
            d["w_d"] = torch.cat(dn, 0).contiguous()                                    # [E*H, I]
            d["w_d_t"] = torch.cat([t.t() for t in dn], 0).contiguous()                # [E*I, H]
            if l in a.inject:
                d["beh_emb"] = bf(W[p + "mlp.behavior_embedding.weight"]).contiguous()
            self.layers.append(d)


class FlatPack(Pack):
    """Pack built from the flat bf16 copy of the fused master buffer (written by the AdamW kernel): the forward
    operands are plain views, only the transposed (dgrad) copies are materialised."""

    def __init__(self, arch: Arch, flat_bf16: torch.Tensor, flat_f32: torch.Tensor):
        a = arch
        Wb = flat_views(a, flat_bf16)
        Wf = flat_views(a, flat_f32)
        self.emb = Wb["model.embed_tokens.weight"]
        emb_t = torch.zeros(a.hidden, a.v_ld, dtype=BF16, device=self.emb.device)
        emb_t[:, :a.vocab] = self.emb.t()
        self.emb_t = emb_t
        self.norm = Wf["model.norm.weight"]
        self.layers = []
        for l in range(a.n_layers):
            p = f"L{l}."
            d = {"in_norm": Wf[p + "in_norm"], "w_qkv": Wb[p + "w_qkv"], "w_o": Wb[p + "w_o"], "qn": Wf[p + "qn"],
                 "kn": Wf[p + "kn"], "post_norm": Wf[p + "post_norm"]}
            d["w_qkv_t"] = d["w_qkv"].t().contiguous()
            d["w_o_t"] = d["w_o"].t().contiguous()
            if l in a.cross:
                d.update(ps_norm=Wf[p + "ps_norm"], c_w_qkvg=Wb[p + "c_w_qkvg"], c_w_o=Wb[p + "c_w_o"],
                         c_qn=Wf[p + "c_qn"], c_kn=Wf[p + "c_kn"], c_qe=Wb[p + "c_qe"], c_ke=Wb[p + "c_ke"],
                         c_ve=Wb[p + "c_ve"])
                d["c_w_qkvg_t"] = d["c_w_qkvg"].t().contiguous()
                d["c_w_o_t"] = d["c_w_o"].t().contiguous()
            gu, dn = Wb[p + "w_gu"], Wb[p + "w_d"]                       # [E, 2I, Kf], [E, H, I]
            E_, Kf = gu.shape[0], gu.shape[2]
            d["w_gu"] = gu.view(E_ * 2 * a.inter, Kf)
            d["w_gu_t"] = gu.transpose(1, 2).contiguous().view(E_ * Kf, 2 * a.inter)
            d["w_d"] = dn.view(E_ * a.hidden, a.inter)
            d["w_d_t"] = dn.transpose(1, 2).contiguous().view(E_ * a.inter, a.hidden)
            if l in a.inject:
                d["beh_emb"] = Wb[p + "beh_emb"]
            self.layers.append(d)
        self._tr_desc = None
        self._build_transpose_table()          # built here, not lazily: refresh() also runs under CUDA-graph capture

    def _transposes(self):
        """(src, dst) pairs of 2-D bf16 views with dst == src.t(): every dgrad copy of the pack, experts one by one."""
        pairs = [(self.emb, self.emb_t[:, :self.emb.shape[0]])]
        for d in self.layers:
            pairs += [(d["w_qkv"], d["w_qkv_t"]), (d["w_o"], d["w_o_t"])]
            if "c_w_qkvg" in d:
                pairs += [(d["c_w_qkvg"], d["c_w_qkvg_t"]), (d["c_w_o"], d["c_w_o_t"])]
            for w, wt in ((d["w_gu"], d["w_gu_t"]), (d["w_d"], d["w_d_t"])):
                E_ = wt.shape[0] // w.shape[1]
                src, dst = w.view(E_, -1, w.shape[1]), wt.view(E_, w.shape[1], -1)
                pairs += [(src[e], dst[e]) for e in range(E_)]
        return pairs

    def refresh(self):
        """Re-derive the transposed (dgrad) copies IN PLACE after the optimizer rewrote the flat bf16 buffer: every
        operand keeps its address, so CUDA graphs captured over this pack stay valid.  One batched-transpose launch over a
        descriptor table built once (the ~40 per-matrix `copy_` kernels it replaces were 0.26 ms of every step, 2 % of an
        8-GPU step)."""
        if getattr(self, "_tr_desc", None) is None:
            self._build_transpose_table()
        K.transpose_batch(self._tr_desc, self._tr_start, self._tr_n, self._tr_tiles)

    def _build_transpose_table(self):
        rows, starts = [], [0]
        for src, dst in self._transposes():
            assert src.dim() == 2 and dst.shape == (src.shape[1], src.shape[0]) and src.stride(1) == 1 and dst.stride(1) == 1
            rows.append([src.data_ptr(), dst.data_ptr(), src.shape[0], src.shape[1], src.stride(0), dst.stride(0)])
            starts.append(starts[-1] + ((src.shape[0] + 31) // 32) * ((src.shape[1] + 31) // 32))
        dev = self.emb.device
        self._tr_desc = torch.tensor(rows, dtype=torch.int64).to(dev)
        self._tr_start = torch.tensor(starts, dtype=torch.int32).to(dev)
        self._tr_n, self._tr_tiles = len(rows), starts[-1]


_ROPE_CACHE: dict = {}


def rope_tables(arch: Arch, n_pos: int, device):
    """cos/sin [n_pos, head_dim/2] fp32, computed as Qwen3RotaryEmbedding does (fp32 outer product, then cos/sin).
    Cached per (head_dim, theta, n_pos, device): the tables are constants, and a cached tensor keeps its address across
    CUDA-graph replays."""
    key = (arch.head_dim, arch.theta, n_pos, str(device))
    hit = _ROPE_CACHE.get(key)
    if hit is not None:
        return hit
    d = arch.head_dim
    inv = 1.0 / (arch.theta ** (torch.arange(0, d, 2, dtype=torch.float32, device=device) / d))
    f = torch.arange(n_pos, dtype=torch.float32, device=device).unsqueeze(-1) * inv
    tabs = (f.cos().contiguous(), f.sin().contiguous())
    if not torch.cuda.is_current_stream_capturing():
        _ROPE_CACHE[key] = tabs      # never evicted: captured graphs may hold the address (≈ 130 KB per sequence length)
    return tabs


def behaviour_lut(arch: Arch, device):
    lut = torch.arange(arch.vocab, dtype=torch.int32)
    for tok, idx in arch.behavior_maps.items():       # same sequential replacement as the reference router
        lut = torch.where(lut == tok, torch.full_like(lut, idx + 1), lut)
    return lut.to(device)


# dropout call sites within one layer (gamer_dropout_t.site = layer * 8 + kind)
SITE_SELF_P, SITE_SELF_OUT, SITE_CROSS_P, SITE_CROSS_OUT, SITE_FFN_INNER, SITE_FFN_OUT = range(6)


@dataclass
class DropCtx:
    """Training-mode dropout of one forward pass (Qwen3Multi/model.py:139,177,217,235,241; Qwen3Moe/FFN.py:23-26):
    `p_hidden` = config.dropout_rate on the three residual branches and inside the experts, `p_attn` =
    config.attention_dropout on the attention probabilities.  (seed, offset) key the Philox masks; the backward is
    handed the same context and regenerates them."""
    seed: int
    offset: int
    p_hidden: float
    p_attn: float
    offset_dev: torch.Tensor | None = None     # device int32[1] XOR-ed into the key at run time (CUDA-graph replays)

    def site(self, layer: int, kind: int):
        p = self.p_attn if kind in (SITE_SELF_P, SITE_CROSS_P) else self.p_hidden
        if p <= 0.0:
            return None
        return K.Dropout(self.seed & 0xFFFFFFFFFFFFFFFF, self.offset & 0xFFFFFFFF, layer * 8 + kind, float(p),
                         None if self.offset_dev is None else self.offset_dev.data_ptr())


def _site(drop, layer, kind):
    return None if drop is None else drop.site(layer, kind)


@dataclass
class BatchMeta:
    """Per-batch integer side inputs, converted once to the int32 arrays the kernels read."""
    B: int
    L: int
    am: torch.Tensor
    act: torch.Tensor | None
    sess: torch.Tensor | None
    rope_pos: torch.Tensor | None     # [M] int32 (session variants) or None (= token position)
    n_pos: int


def make_meta(arch: Arch, input_ids, attention_mask, actions, session_ids, extended_session_ids) -> BatchMeta:
    B, L = input_ids.shape
    i32 = lambda t: None if t is None else t.to(torch.int32).contiguous()
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids)
    k_self, k_cross = arch.mask_kinds()
    need_act = k_cross is not None and len(arch.cross) > 0
    need_sess = k_self == K.MASK_SESSION
    if need_act and actions is None:
        raise ValueError("`actions` is required by this backbone's behaviour-level attention mask")
    if need_sess and session_ids is None:
        raise AssertionError("Session IDs must be provided to generate session-wise causal mask.")
    # RoPE table rows: token positions 0..L-1 plus the few positions a decode call appends; extended session ids are
    # < P * n_sessions <= L + P, so the same bound covers the session variants
    rope_pos, n_pos = None, L + 2 * arch.P + 8
    if arch.session_rope() and extended_session_ids is not None:
        rope_pos = i32(extended_session_ids).view(-1)
    return BatchMeta(B, L, i32(attention_mask), i32(actions) if need_act else None,
                     i32(session_ids) if need_sess else None, rope_pos, n_pos)


# ======================================================================================================================
# forward
# ======================================================================================================================
def _attention_fwd(arch, meta, x, norm_w, w_qkv, qn, kn, w_o, kind, tabs, act_idx=None, embs=None, gated=False,
                   drop_p=None, drop_out=None):
    """pre-norm attention sub-block: x + dropout(attention(norm(x))).  Returns (x_out, saved)."""
    M = x.shape[0]
    h, rstd = K.rmsnorm_fwd(x, norm_w, arch.eps)
    n_proj = w_qkv.shape[0]
    raw = K.gemm_tn(h, w_qkv, n_proj)
    qe, ke, ve = embs if embs is not None else (None, None, None)
    rot = K.qk_norm_rope_fwd(raw, meta.L, arch.n_q, arch.n_kv, arch.head_dim, tabs[0], tabs[1], qn, kn, arch.eps,
                             pos_ids=meta.rope_pos, q_emb=qe, k_emb=ke, v_emb=ve, act_idx=act_idx)
    o, lse, vmean, keep = K.attn_fwd(rot, meta.B, meta.L, arch.n_q, arch.n_kv, arch.head_dim, kind, arch.P, meta.am,
                                     meta.act, meta.sess, arch.head_dim ** -0.5, drop=drop_p)
    if gated:
        y = K.gemm_tn(o, w_o, arch.hidden)
        x_out = K.gate_residual_fwd(x, y, raw[:, arch.qkv_w:], drop=drop_out)
    else:
        y = None
        x_out = K.gemm_tn(o, w_o, arch.hidden, resid=x, drop=drop_out)
    return x_out, dict(x=x, rstd=rstd, h=h, raw=raw, rot=rot, o=o, lse=lse, y=y, vmean=vmean, keep=keep)


def forward_stack(arch: Arch, pack: Pack, input_ids, meta: BatchMeta, lut, save: bool, kv_sink: list | None = None,
                  drop: DropCtx | None = None):
    """-> (final normed hidden [M,H] bf16, ctx).  ctx holds what the backward needs when save=True.  `kv_sink` (decode
    prefill) receives per layer {"self": (rot, vmean), "cross": (rot, vmean)} — the rotated q|k|v buffers double as the
    prompt K/V cache."""
    B, L = meta.B, meta.L
    dev = input_ids.device
    x, pos_idx, beh_idx, act_idx = K.embed_route(input_ids, pack.emb, lut, arch.n_beh, arch.P, arch.pad, arch.eos)
    any_sparse = len(arch.sparse) > 0
    perm = rows = seg = None
    Mp = B * L
    if any_sparse:
        perm, rows, seg = K.route_perm(pos_idx, B, L, arch.n_exp)
        Mp = rows.shape[0]
    tabs = rope_tables(arch, meta.n_pos, dev)
    k_self, k_cross = arch.mask_kinds()
    ctx = dict(layers=[], pos_idx=pos_idx, beh_idx=beh_idx, act_idx=act_idx, perm=perm, rows=rows, seg=seg, tabs=tabs)
    for l in range(arch.n_layers):
        d = pack.layers[l]
        saved = {}
        x, s_self = _attention_fwd(arch, meta, x, d["in_norm"], d["w_qkv"], d["qn"], d["kn"], d["w_o"], k_self, tabs,
                                   drop_p=_site(drop, l, SITE_SELF_P), drop_out=_site(drop, l, SITE_SELF_OUT))
        saved["self"] = s_self
        if l in arch.cross:
            x, s_cross = _attention_fwd(arch, meta, x, d["ps_norm"], d["c_w_qkvg"], d["c_qn"], d["c_kn"], d["c_w_o"],
                                        k_cross, tabs, act_idx=act_idx, embs=(d["c_qe"], d["c_ke"], d["c_ve"]),
                                        gated=True, drop_p=_site(drop, l, SITE_CROSS_P),
                                        drop_out=_site(drop, l, SITE_CROSS_OUT))
            saved["cross"] = s_cross
        # routed FFN
        sparse = l in arch.sparse
        inject = l in arch.inject
        Kf = arch.hidden + (arch.beh_dim if inject else 0)
        if sparse:
            # padding rows of the permuted space must hold zeros (finite operands for the grouped GEMMs and wgrads);
            # the norm kernel writes every mapped row, so only the < 128 padding rows per expert are cleared
            hp = K.zero_unmapped_rows(torch.empty(Mp, Kf, dtype=BF16, device=dev), rows)
            _, rstd = K.rmsnorm_fwd(x, d["post_norm"], arch.eps, out=hp, row_map=perm,
                                    cat_table=d.get("beh_emb"), cat_idx=beh_idx if inject else None)
            gu = K.gemm_tn(hp, d["w_gu"], 2 * arch.inter, rows=Mp, n_groups=arch.n_exp, seg_off=seg)
            a = K.swiglu_fwd(gu, arch.inter, row_ids=rows, drop=_site(drop, l, SITE_FFN_INNER))
            x_out = torch.empty_like(x)
            K.gemm_tn(a, d["w_d"], arch.hidden, rows=Mp, n_groups=arch.n_exp, seg_off=seg, out=x_out, resid=x,
                      row_map=rows, drop=_site(drop, l, SITE_FFN_OUT))
        else:
            hp, rstd = K.rmsnorm_fwd(x, d["post_norm"], arch.eps, cat_table=d.get("beh_emb"),
                                     cat_idx=beh_idx if inject else None)
            gu = K.gemm_tn(hp, d["w_gu"], 2 * arch.inter)
            a = K.swiglu_fwd(gu, arch.inter, drop=_site(drop, l, SITE_FFN_INNER))
            x_out = K.gemm_tn(a, d["w_d"], arch.hidden, resid=x, drop=_site(drop, l, SITE_FFN_OUT))
        saved["ffn"] = dict(x=x, rstd=rstd, hp=hp, gu=gu, a=a)
        x = x_out
        if save:
            ctx["layers"].append(saved)
        if kv_sink is not None:
            kv_sink.append({k: (saved[k]["rot"], saved[k]["vmean"]) for k in ("self", "cross") if k in saved})
    hidden, rstd = K.rmsnorm_fwd(x, pack.norm, arch.eps)
    if save:
        ctx["final"] = dict(x=x, rstd=rstd)
    return hidden, ctx


def lm_head_logits(arch: Arch, pack: Pack, hidden, alpha=1.0):
    """fp32 logits [M, V] (a view of a [M, v_ld] buffer so rows stay 16-byte aligned)."""
    M = hidden.shape[0]
    buf = torch.empty(M, arch.v_ld, dtype=torch.float32, device=hidden.device)
    K.gemm_tn(hidden, pack.emb, arch.vocab, out=buf, alpha=alpha)
    return buf[:, :arch.vocab]


def shift_labels(labels):
    """ForCausalLMLoss: pad with -100 on the right, drop the first column."""
    return torch.nn.functional.pad(labels, (0, 1), value=-100)[..., 1:].contiguous().view(-1)


def lm_head_loss(arch: Arch, pack: Pack, hidden, shifted, inv_norm, temperature):
    """Chunked lm_head GEMM + fused CE; logits are never materialised for the whole batch."""
    M = hidden.shape[0]
    total = torch.zeros((), dtype=torch.float32, device=hidden.device)
    buf = torch.empty(min(M, CE_CHUNK_ROWS), arch.v_ld, dtype=torch.float32, device=hidden.device)
    for r0 in range(0, M, CE_CHUNK_ROWS):
        r1 = min(M, r0 + CE_CHUNK_ROWS)
        K.gemm_tn(hidden[r0:r1], pack.emb, arch.vocab, out=buf[: r1 - r0], alpha=1.0 / temperature)
        loss_row = K.ce_fwd_bwd(buf[: r1 - r0], shifted[r0:r1], arch.vocab, None, 1.0)
        total = total + loss_row.sum()
    return total * inv_norm.view(())


# ======================================================================================================================
# backward
# ======================================================================================================================
def _attention_bwd(arch, meta, s, dx_out, norm_w, w_qkv_t, qn, kn, w_o_t, kind, tabs, G, names, act_idx=None, embs=None,
                   gated=False, drop_p=None, drop_out=None):
    """Returns dx (grad wrt the sub-block input).  Parameter grads are accumulated into G[name] (fp32)."""
    dev = dx_out.device
    M = dx_out.shape[0]
    n_proj = w_qkv_t.shape[1]
    draw = torch.empty(M, n_proj, dtype=BF16, device=dev)
    if gated:
        dy = K.gate_residual_bwd(dx_out, s["y"], s["raw"][:, arch.qkv_w:], draw[:, arch.qkv_w:], drop=drop_out)
    else:
        dy = dx_out if drop_out is None else K.dropout_apply(dx_out, drop_out)
    d_o = K.gemm_tn(dy, w_o_t, arch.q_w)
    K.gemm_wgrad(dy, s["o"], arch.hidden, arch.q_w, G[names["o"]].view(1, arch.hidden, arch.q_w))
    drot = torch.empty(M, arch.qkv_w, dtype=BF16, device=dev)
    K.attn_bwd(s["rot"], s["o"], d_o, s["lse"], meta.B, meta.L, arch.n_q, arch.n_kv, arch.head_dim, kind, arch.P,
               meta.am, meta.act, meta.sess, arch.head_dim ** -0.5, drot, drop=drop_p, keep=s["keep"])
    qe, ke, ve = embs if embs is not None else (None, None, None)
    K.qk_norm_rope_bwd(s["raw"], drot, draw, meta.L, arch.n_q, arch.n_kv, arch.head_dim, tabs[0], tabs[1], qn, kn,
                       arch.eps, G[names["qn"]], G[names["kn"]], pos_ids=meta.rope_pos, q_emb=qe, k_emb=ke, v_emb=ve,
                       act_idx=act_idx, emb_rows=arch.n_beh + 1,
                       d_q_emb=G.get(names.get("qe")), d_k_emb=G.get(names.get("ke")), d_v_emb=G.get(names.get("ve")))
    dh = K.gemm_tn(draw, w_qkv_t, arch.hidden)
    K.gemm_wgrad(draw, s["h"], n_proj, arch.hidden, G[names["qkv"]].view(1, n_proj, arch.hidden))
    return K.rmsnorm_bwd(s["x"], norm_w, s["rstd"], arch.eps, dh, G[names["norm"]], dres=dx_out)


def backward_stack(arch: Arch, pack: Pack, meta: BatchMeta, ctx, d_hidden, G: dict, on_layer_done=None,
                   drop: DropCtx | None = None):
    """d_hidden: grad wrt the final normed hidden [M,H] bf16.  G: fused fp32 gradient buffers (see `grad_buffers`).
    Returns dx0 (grad wrt the embedding output)."""
    tabs = ctx["tabs"]
    k_self, k_cross = arch.mask_kinds()
    f = ctx["final"]
    dx = K.rmsnorm_bwd(f["x"], pack.norm, f["rstd"], arch.eps, d_hidden, G["model.norm.weight"])
    if on_layer_done is not None:
        on_layer_done(arch.n_layers)
    for l in reversed(range(arch.n_layers)):
        d = pack.layers[l]
        saved = ctx["layers"][l]
        p = f"L{l}."
        s = saved["ffn"]
        sparse = l in arch.sparse
        inject = l in arch.inject
        Kf = arch.hidden + (arch.beh_dim if inject else 0)
        E = arch.n_exp if sparse else 1
        if sparse:
            Mp = s["hp"].shape[0]
            dxp = K.gather_rows(dx, ctx["rows"], Mp, drop=_site(drop, l, SITE_FFN_OUT))
            seg = ctx["seg"]
        else:
            Mp = dx.shape[0]
            d_out = _site(drop, l, SITE_FFN_OUT)
            dxp, seg = (dx if d_out is None else K.dropout_apply(dx, d_out)), None
        da = K.gemm_tn(dxp, d["w_d_t"], arch.inter, rows=Mp, n_groups=E, seg_off=seg)
        K.gemm_wgrad(dxp, s["a"], arch.hidden, arch.inter, G[p + "w_d"], rows=Mp, n_groups=E, seg_off=seg)
        dgu = K.swiglu_bwd(s["gu"], da, arch.inter, row_ids=ctx["rows"] if sparse else None,
                           drop=_site(drop, l, SITE_FFN_INNER))
        dhp = K.gemm_tn(dgu, d["w_gu_t"], Kf, rows=Mp, n_groups=E, seg_off=seg)
        K.gemm_wgrad(dgu, s["hp"], 2 * arch.inter, Kf, G[p + "w_gu"], rows=Mp, n_groups=E, seg_off=seg)
        dx = K.rmsnorm_bwd(s["x"], d["post_norm"], s["rstd"], arch.eps, dhp, G[p + "post_norm"],
                           row_map=ctx["perm"] if sparse else None, dres=dx, cat_idx=ctx["beh_idx"] if inject else None,
                           cat_dim=arch.beh_dim if inject else 0, cat_rows=arch.n_beh + 1,
                           dcat=G.get(p + "beh_emb"))
        if l in arch.cross:
            names = dict(o=p + "c_w_o", qkv=p + "c_w_qkvg", qn=p + "c_qn", kn=p + "c_kn", norm=p + "ps_norm",
                         qe=p + "c_qe", ke=p + "c_ke", ve=p + "c_ve")
            dx = _attention_bwd(arch, meta, saved["cross"], dx, d["ps_norm"], d["c_w_qkvg_t"], d["c_qn"], d["c_kn"],
                                d["c_w_o_t"], k_cross, tabs, G, names, act_idx=ctx["act_idx"],
                                embs=(d["c_qe"], d["c_ke"], d["c_ve"]), gated=True,
                                drop_p=_site(drop, l, SITE_CROSS_P), drop_out=_site(drop, l, SITE_CROSS_OUT))
        names = dict(o=p + "w_o", qkv=p + "w_qkv", qn=p + "qn", kn=p + "kn", norm=p + "in_norm")
        dx = _attention_bwd(arch, meta, saved["self"], dx, d["in_norm"], d["w_qkv_t"], d["qn"], d["kn"], d["w_o_t"],
                            k_self, tabs, G, names, drop_p=_site(drop, l, SITE_SELF_P),
                            drop_out=_site(drop, l, SITE_SELF_OUT))
        saved.clear()
        if on_layer_done is not None:
            on_layer_done(l)
    return dx


def fused_blocks(arch: Arch):
    """Ordered (key, shape) list of the fused parameter blocks: the layout of the flat fp32 master-weight, gradient and
    optimizer-state buffers.  Forward order, so one layer is one contiguous range (= one all-reduce bucket)."""
    blocks = [("model.embed_tokens.weight", (arch.vocab, arch.hidden))]
    for l in range(arch.n_layers):
        p = f"L{l}."
        inject = l in arch.inject
        Kf = arch.hidden + (arch.beh_dim if inject else 0)
        E_ = arch.n_exp if l in arch.sparse else 1
        blocks += [(p + "in_norm", (arch.hidden,)), (p + "w_qkv", (arch.qkv_w, arch.hidden)),
                   (p + "w_o", (arch.hidden, arch.q_w)), (p + "qn", (arch.head_dim,)), (p + "kn", (arch.head_dim,))]
        if l in arch.cross:
            blocks += [(p + "ps_norm", (arch.hidden,)), (p + "c_w_qkvg", (arch.qkv_w + arch.hidden, arch.hidden)),
                       (p + "c_w_o", (arch.hidden, arch.q_w)), (p + "c_qn", (arch.head_dim,)),
                       (p + "c_kn", (arch.head_dim,)), (p + "c_qe", (arch.n_beh + 1, arch.q_w)),
                       (p + "c_ke", (arch.n_beh + 1, arch.kv_w)), (p + "c_ve", (arch.n_beh + 1, arch.kv_w))]
        blocks += [(p + "post_norm", (arch.hidden,)), (p + "w_gu", (E_, 2 * arch.inter, Kf)),
                   (p + "w_d", (E_, arch.hidden, arch.inter))]
        if inject:
            blocks += [(p + "beh_emb", (arch.n_beh + 1, arch.beh_dim))]
    blocks += [("model.norm.weight", (arch.hidden,))]
    return blocks


def flat_views(arch: Arch, flat: torch.Tensor) -> dict:
    """{fused key: view} over a flat buffer laid out by `fused_blocks` (every block start is 16-byte aligned)."""
    out, off = {}, 0
    for key, shape in fused_blocks(arch):
        n = 1
        for s in shape:
            n *= s
        out[key] = flat[off:off + n].view(*shape)
        off += (n + 3) // 4 * 4
    return out


def flat_size(arch: Arch) -> int:
    off = 0
    for _, shape in fused_blocks(arch):
        n = 1
        for s in shape:
            n *= s
        off += (n + 3) // 4 * 4
    return off


def layer_ranges(arch: Arch):
    """[(name, start, end)] element ranges of the flat layout: embedding, each layer, final norm — the gradient buckets."""
    ranges, off, cur, cur_start = [], 0, None, 0
    for key, shape in fused_blocks(arch):
        n = 1
        for s in shape:
            n *= s
        grp = key.split(".")[0] if key.startswith("L") else key
        if grp != cur:
            if cur is not None:
                ranges.append((cur, cur_start, off))
            cur, cur_start = grp, off
        off += (n + 3) // 4 * 4
    ranges.append((cur, cur_start, off))
    return ranges


def grad_buffers(arch: Arch, device, flat: torch.Tensor | None = None):
    """Fused fp32 gradient buffers (views of one flat allocation) in the layouts the wgrad kernels write."""
    if flat is None:
        flat = torch.zeros(flat_size(arch), dtype=torch.float32, device=device)
    G = flat_views(arch, flat)
    G["_flat"] = flat
    return G


def unfuse_grads(arch: Arch, G: dict) -> dict:
    """Fused gradient buffers -> {state-dict key: fp32 grad view}."""
    out = {"model.embed_tokens.weight": G["model.embed_tokens.weight"], "model.norm.weight": G["model.norm.weight"]}
    post = "post_attention_layernorm" if arch.variant in NO_CROSS_VARIANTS else "post_cross_attention_layernorm"
    q, kv, H, I = arch.q_w, arch.kv_w, arch.hidden, arch.inter
    for l in range(arch.n_layers):
        p, k = f"L{l}.", f"model.layers.{l}."
        out[k + "input_layernorm.weight"] = G[p + "in_norm"]
        g = G[p + "w_qkv"]
        out[k + "self_attn.q_proj.weight"] = g[:q]
        out[k + "self_attn.k_proj.weight"] = g[q:q + kv]
        out[k + "self_attn.v_proj.weight"] = g[q + kv:]
        out[k + "self_attn.o_proj.weight"] = G[p + "w_o"]
        out[k + "self_attn.q_norm.weight"] = G[p + "qn"]
        out[k + "self_attn.k_norm.weight"] = G[p + "kn"]
        if l in arch.cross:
            out[k + "post_self_attention_layernorm.weight"] = G[p + "ps_norm"]
            g = G[p + "c_w_qkvg"]
            out[k + "cross_attn.q_proj.weight"] = g[:q]
            out[k + "cross_attn.k_proj.weight"] = g[q:q + kv]
            out[k + "cross_attn.v_proj.weight"] = g[q + kv:q + 2 * kv]
            out[k + "cross_attn.gating.weight"] = g[q + 2 * kv:]
            out[k + "cross_attn.o_proj.weight"] = G[p + "c_w_o"]
            out[k + "cross_attn.q_norm.weight"] = G[p + "c_qn"]
            out[k + "cross_attn.k_norm.weight"] = G[p + "c_kn"]
            out[k + "cross_attn.q_behavior_embedding.weight"] = G[p + "c_qe"]
            out[k + "cross_attn.k_behavior_embedding.weight"] = G[p + "c_ke"]
            out[k + "cross_attn.v_behavior_embedding.weight"] = G[p + "c_ve"]
        out[k + post + ".weight"] = G[p + "post_norm"]
        gu, dn = G[p + "w_gu"], G[p + "w_d"]
        if l in arch.sparse:
            for e in range(arch.n_exp):
                q_ = k + f"mlp.experts.expert_{e}."
                out[q_ + "gate_proj.weight"] = gu[e, :I]
                out[q_ + "up_proj.weight"] = gu[e, I:]
                out[q_ + "down_proj.weight"] = dn[e]
        else:
            out[k + "mlp.mlp.gate_proj.weight"] = gu[0, :I]
            out[k + "mlp.mlp.up_proj.weight"] = gu[0, I:]
            out[k + "mlp.mlp.down_proj.weight"] = dn[0]
        if l in arch.inject:
            out[k + "mlp.behavior_embedding.weight"] = G[p + "beh_emb"]
    return out


def lm_head_backward(arch: Arch, pack: Pack, hidden, shifted, scale_dev, temperature, d_emb, loss_sum=None):
    """Recompute the logits chunk by chunk, emit dlogits (already scaled by grad_out * inv_norm / T), and run the
    lm_head dgrad + wgrad.  Returns d_hidden [M,H] bf16; accumulates the tied-weight gradient into d_emb.  The fused CE
    kernel produces the per-row losses in the same pass: when `loss_sum` (fp32 scalar tensor) is given they are added
    to it, so a caller that always runs forward and backward together (the native trainer) skips the forward head."""
    M = hidden.shape[0]
    dev = hidden.device
    d_hidden = torch.empty(M, arch.hidden, dtype=BF16, device=dev)
    rows_c = min(M, CE_CHUNK_ROWS)
    buf = torch.empty(rows_c, arch.v_ld, dtype=torch.float32, device=dev)
    dl = torch.empty(rows_c, arch.v_ld, dtype=BF16, device=dev)
    for r0 in range(0, M, CE_CHUNK_ROWS):
        r1 = min(M, r0 + CE_CHUNK_ROWS)
        n = r1 - r0
        K.gemm_tn(hidden[r0:r1], pack.emb, arch.vocab, out=buf[:n], alpha=1.0 / temperature)
        loss_row = K.ce_fwd_bwd(buf[:n], shifted[r0:r1], arch.vocab, scale_dev, 1.0 / temperature, dlogits=dl[:n])
        if loss_sum is not None:
            loss_sum += loss_row.sum()
        K.gemm_tn(dl[:n, :arch.vocab], pack.emb_t[:, :arch.vocab], arch.hidden, K=arch.vocab, out=d_hidden[r0:r1])
        K.gemm_wgrad(dl[:n, :arch.vocab], hidden[r0:r1], arch.vocab, arch.hidden,
                     d_emb.view(1, arch.vocab, arch.hidden))
    return d_hidden


def loss_forward(arch, pack, meta, lut, input_ids, shifted, inv_norm, temperature, drop: DropCtx | None = None,
                 defer_loss: bool = False):
    """Forward of the training step.  Returns (loss, state for `loss_backward`).  defer_loss=True skips the lm_head +
    CE forward: `loss_backward` then returns the loss from its own (fused forward + backward) pass over the logits."""
    hidden, saved = forward_stack(arch, pack, input_ids, meta, lut, save=True, drop=drop)
    loss = None if defer_loss else lm_head_loss(arch, pack, hidden, shifted, inv_norm, temperature)
    sort_buf = K.embed_sort(input_ids.view(-1), arch.vocab, arch.pad)
    return loss, dict(saved=saved, hidden=hidden, shifted=shifted, inv_norm=inv_norm, temperature=temperature,
                      sort_buf=sort_buf, meta=meta, drop=drop, defer_loss=defer_loss)


def loss_backward(arch, pack, st, grad_out, G, on_layer_done=None):
    """Backward of the training step into the fused gradient buffers G (accumulating).  grad_out: device scalar."""
    scale = (grad_out.float().reshape(()) * st["inv_norm"].view(())).reshape(1).contiguous()
    loss_sum = torch.zeros((), dtype=torch.float32, device=st["hidden"].device) if st.get("defer_loss") else None
    d_hidden = lm_head_backward(arch, pack, st["hidden"], st["shifted"], scale, st["temperature"],
                                G["model.embed_tokens.weight"], loss_sum=loss_sum)
    dx0 = backward_stack(arch, pack, st["meta"], st["saved"], d_hidden, G, on_layer_done=on_layer_done,
                         drop=st.get("drop"))
    K.embed_bwd(dx0, arch.vocab, st["sort_buf"], G["model.embed_tokens.weight"])
    if on_layer_done is not None:
        on_layer_done(-1)
    st["saved"] = None
    return None if loss_sum is None else loss_sum * st["inv_norm"].view(())


class DecoderLossFunction(torch.autograd.Function):
    """loss = CE(lm_head(decoder(input_ids)) / T).  Inputs after the fixed arguments are the fp32 master parameters in
    `param_names(arch)` order; backward returns their gradients (views of one flat fused buffer)."""

    @staticmethod
    def forward(ctx, arch, pack, meta, lut, input_ids, shifted, inv_norm, temperature, hooks, drop, *params):
        loss, st = loss_forward(arch, pack, meta, lut, input_ids, shifted, inv_norm, temperature, drop=drop)
        ctx.arch, ctx.pack, ctx.st, ctx.hooks = arch, pack, st, hooks
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        arch = ctx.arch
        G = grad_buffers(arch, ctx.st["hidden"].device)
        on_done = ctx.hooks.get("on_layer_done") if ctx.hooks else None
        loss_backward(arch, ctx.pack, ctx.st, grad_out, G, on_layer_done=(lambda l: on_done(l, G)) if on_done else None)
        named = unfuse_grads(arch, G)
        grads = tuple(named[n] for n in param_names(arch))
        ctx.st = None
        return (None,) * 10 + grads
