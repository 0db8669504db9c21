"""Trie-constrained beam search on the GPU: the evaluation decode of tasks/test_SMB_decoder.py:159-177.

What the reference does per batch (third-party HF `_beam_search`): expand every prompt x num_beams
(prefill is computed num_beams times), then per new token: forward one token with a DynamicCache, fp32 full-vocab
log-softmax, a Python double loop over batch x beams calling `Trie.get` on `.tolist()`-ed sequences, add the running
score, top-2K over K*V, gather, `index_select` the whole KV cache.

Here: the prompt is prefilled ONCE per user; its rotated K/V (the `rot` buffers of the forward pass) are the shared
prompt cache; each beam row only owns the K/V of its <= 4 generated tokens, reached through an ancestry table instead
of reordering caches; the trie is a CSR array in HBM and the per-step log-softmax + child mask + score add + per-user
top-K is one kernel.  Semantics kept: scores are log-probabilities over the FULL vocabulary (mask applied after
normalisation, Q6), running scores start at [0, -1e9, ...], the final score is sum/gen_len, sequences come back
best-first per user, prompt included.  Because the trie never allows EOS, no hypothesis can finish early and HF's
2K-candidate bookkeeping reduces to top-K.  Ties are broken towards the lower (beam, token) index.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import engine as E
from . import kernels as K
from .trie import FlatTrie, Trie, _PrefixFn


@dataclass
class BeamSearchOutput:
    """Mirrors the fields of HF's GenerateBeamDecoderOnlyOutput that test_SMB_decoder.py reads."""
    sequences: torch.Tensor
    sequences_scores: torch.Tensor
    beam_indices: torch.Tensor | None = None
    generated: torch.Tensor | None = None     # [B * num_return_sequences, max_new_tokens]: the new tokens alone (the decoded
                                              # item tuples) — what an evaluation loop needs on the host; `sequences` repeats
                                              # every prompt num_return_sequences times (20.7 MB at 256 users x 20 beams)

    def __getitem__(self, k):
        return getattr(self, k)


def _resolve_constraint(prefix_allowed_tokens_fn, candidate_trie, vocab, pad, device):
    trie, last = None, None
    if isinstance(prefix_allowed_tokens_fn, _PrefixFn):
        trie, last = prefix_allowed_tokens_fn.trie, prefix_allowed_tokens_fn.last_token_set
    elif candidate_trie is not None:
        trie = candidate_trie
    elif prefix_allowed_tokens_fn is not None:
        raise NotImplementedError(
            "generate() runs the prefix constraint on the GPU and needs the trie itself: build the callable with "
            "gamer_b200.trie.prefix_allowed_tokens_fn_by_last_token (same signature as SeqRec.generation.trie) or pass "
            "candidate_trie=")
    else:
        raise ValueError("constrained beam search needs prefix_allowed_tokens_fn or candidate_trie")
    flat = trie if isinstance(trie, FlatTrie) else trie.flat()
    # device copies are cached on the host objects: a captured decode graph holds their addresses
    cache = flat.__dict__.setdefault("_dev_cache", {})
    if device not in cache:
        # the beam step indexes the logits with the trie's tokens: validate them once, on the host copy
        if flat.child_tok.numel() and (int(flat.child_tok.max()) >= vocab or int(flat.child_tok.min()) < 0):
            raise ValueError(f"candidate trie holds token ids outside the vocabulary [0, {vocab})")
        cache[device] = flat.to(device)
    flat_dev = cache[device]
    owner = prefix_allowed_tokens_fn if isinstance(prefix_allowed_tokens_fn, _PrefixFn) else flat
    bcache = owner.__dict__.setdefault("_bitmap_cache", {})
    bkey = (device, vocab)
    if bkey not in bcache:
        bitmap = torch.zeros(vocab, dtype=torch.uint8)       # last is None: the whole sentence is the prefix
        if last is not None:
            idx = torch.tensor(sorted(t for t in last if 0 <= t < vocab), dtype=torch.long)
            bitmap[idx] = 1
        bcache[bkey] = bitmap.to(device)
    return flat_dev, bcache[bkey]


def _decode_layer_attention(arch, d, meta, x, s, R, B, beams, L0, S_max, tabs, gen, anc, prompt, kind, norm_w, w_qkv, qn,
                            kn, w_o, pos0, pos_ids, act_idx=None, embs=None, gated=False):
    h, _ = K.rmsnorm_fwd(x, norm_w, arch.eps)
    raw = K.gemm_tn(h, w_qkv, w_qkv.shape[0])
    qe, ke, ve = embs if embs is not None else (None, None, None)
    rot = gen[s]
    K.qk_norm_rope_fwd(raw, 1, arch.n_q, arch.n_kv, arch.head_dim, tabs[0], tabs[1], qn, kn, arch.eps, pos_ids=pos_ids,
                       pos0=pos0, q_emb=qe, k_emb=ke, v_emb=ve, act_idx=act_idx, out=rot)
    prompt_rot, vmean = prompt
    o = K.attn_decode(rot, prompt_rot, gen, gen.stride(0), anc, B, beams, L0, s + 1, arch.n_q, arch.n_kv, arch.head_dim,
                      S_max, meta.am, meta.act, meta.sess, kind, vmean, arch.head_dim ** -0.5)
    if gated:
        y = K.gemm_tn(o, w_o, arch.hidden)
        return K.gate_residual_fwd(x, y, raw[:, arch.qkv_w:])
    return K.gemm_tn(o, w_o, arch.hidden, resid=x)


def decode_step(arch, pack, lut, meta, state, tokens, s):
    """One cached step: `tokens` [R] int64 are appended at position L0+s.  Returns fp32 logits [R, V]."""
    B, beams, L0, S_max = state["B"], state["beams"], state["L0"], state["S_max"]
    R = B * beams
    pos0 = L0 + s
    x, pos_idx, beh_idx, act_idx = K.embed_route(tokens.view(R, 1), pack.emb, lut, arch.n_beh, arch.P, arch.pad, arch.eos,
                                                 ctx=state["ctx"], pos0=pos0)
    k_self, k_cross = arch.mask_kinds()
    pos_ids = None
    if state["rope_next"] is not None:
        pos_ids = (state["rope_next"] + s).to(torch.int32).contiguous()
    expert = (pos0 % arch.P) + 1          # position-routed expert of every row (trie tokens are never pad/eos)
    tabs = state["tabs"]
    for l in range(arch.n_layers):
        d = pack.layers[l]
        x = _decode_layer_attention(arch, d, meta, x, s, R, B, beams, L0, S_max, tabs, state["gen"][l]["self"], state["anc"],
                                    state["prompt"][l]["self"], k_self, d["in_norm"], d["w_qkv"], d["qn"], d["kn"],
                                    d["w_o"], pos0, pos_ids)
        if l in arch.cross:
            x = _decode_layer_attention(arch, d, meta, x, s, R, B, beams, L0, S_max, tabs, state["gen"][l]["cross"],
                                        state["anc"], state["prompt"][l]["cross"], k_cross, d["ps_norm"], d["c_w_qkvg"],
                                        d["c_qn"], d["c_kn"], d["c_w_o"], pos0, pos_ids, act_idx=act_idx,
                                        embs=(d["c_qe"], d["c_ke"], d["c_ve"]), gated=True)
        inject = l in arch.inject
        Kf = arch.hidden + (arch.beh_dim if inject else 0)
        e = expert if l in arch.sparse else 0
        hp, _ = K.rmsnorm_fwd(x, d["post_norm"], arch.eps, cat_table=d.get("beh_emb"), cat_idx=beh_idx if inject else None)
        gu = K.gemm_tn(hp, d["w_gu"][e * 2 * arch.inter:(e + 1) * 2 * arch.inter], 2 * arch.inter, K=Kf)
        a = K.swiglu_fwd(gu, arch.inter)
        x = K.gemm_tn(a, d["w_d"][e * arch.hidden:(e + 1) * arch.hidden], arch.hidden, resid=x)
    hidden, _ = K.rmsnorm_fwd(x, pack.norm, arch.eps)
    return E.lm_head_logits(arch, pack, hidden)


def _raise_on(code: int):
    if code == 1:
        raise ValueError("`prefix_allowed_tokens_fn` returned an empty list for a beam: the prompt suffix is not a "
                         "prefix of any candidate (cf. PrefixConstrainedLogitsProcessor)")
    if code == 2:
        raise RuntimeError("beam step candidate buffer overflow")


def beam_search_device(B, beams, S, vocab, flat, node0, logits0, advance, device):
    """The beam bookkeeping, independent of where the logits come from; no host synchronisation (graph-capturable).

    node0 [B] int32: trie node of every user's prompt suffix; logits0 [B*beams, >=vocab] fp32: next-token logits of the
    (identical) beams of each user; advance(s, parent_rows [R] int64, tokens [R] int64) -> logits [R, >=vocab] for the
    rows obtained by appending `tokens` to the hypotheses `parent_rows`.
    Returns (generated tokens [B, beams, S] int64, summed log-probabilities [B, beams] fp32, error word [1] int32),
    best-first per user.
    """
    node = node0.repeat_interleave(beams).contiguous()
    run = torch.zeros(B, beams, dtype=torch.float32, device=device)
    run[:, 1:] = -1e9                                                        # HF: beam_scores[:, 1:] = -1e9
    err = torch.zeros(1, dtype=torch.int32, device=device)
    toks, parents = [], []
    row_base = (torch.arange(B, device=device) * beams).view(B, 1)
    logits = logits0
    for s in range(S):
        run, parent, tok, node = K.beam_step(logits, vocab, B, beams, run, node, flat, err)
        toks.append(tok)
        parents.append(parent)
        if s + 1 == S:
            break
        prow = (parent.long() + row_base).view(-1)                           # parent row of every new beam row
        logits = advance(s, prow, tok.view(-1).long())
        node = node.view(-1).contiguous()
    # backtrack the token choices through the parent pointers
    gen = torch.empty(B, beams, S, dtype=torch.long, device=device)
    idx = torch.arange(beams, device=device).view(1, beams).expand(B, beams)
    for s in reversed(range(S)):
        gen[:, :, s] = torch.gather(toks[s].long(), 1, idx)
        idx = torch.gather(parents[s].long(), 1, idx)
    return gen, run, err


def beam_search_core(B, beams, S, vocab, flat, node0, logits0, advance, device):
    """beam_search_device + the error check (one host sync).  Returns (generated tokens, summed log-probabilities)."""
    gen, run, err = beam_search_device(B, beams, S, vocab, flat, node0, logits0, advance, device)
    _raise_on(int(err.item()))
    return gen, run


def _decode_on_device(arch, pack, lut, flat, last_bitmap, beams, S, input_ids, attention_mask, session_ids,
                      extended_session_ids, actions):
    """Prefill + constrained beam search, everything enqueued on the current stream without a host sync."""
    dev = input_ids.device
    B, L0 = input_ids.shape
    R = B * beams
    meta = E.make_meta(arch, input_ids, attention_mask, actions, session_ids, extended_session_ids)

    # ---- prefill once per user; the rotated q|k|v buffers are the prompt K/V cache ------------------------------
    sink = []
    hidden, ctx = E.forward_stack(arch, pack, input_ids, meta, lut, save=False, kv_sink=sink)
    last_hidden = hidden.view(B, L0, -1)[:, -1, :].contiguous()
    logits0 = E.lm_head_logits(arch, pack, last_hidden)                      # [B, V] fp32
    logits0 = logits0.repeat_interleave(beams, dim=0).contiguous()           # beams of a user start identical

    state = dict(B=B, beams=beams, L0=L0, S_max=S, tabs=ctx["tabs"], prompt=sink,
                 ctx=input_ids.repeat_interleave(beams, dim=0).contiguous(),
                 anc=torch.zeros(R, S, dtype=torch.int32, device=dev), rope_next=None, gen=[])
    for l in range(arch.n_layers):
        g = {"self": torch.empty(S, R, arch.qkv_w, dtype=torch.bfloat16, device=dev)}
        if l in arch.cross:
            g["cross"] = torch.empty(S, R, arch.qkv_w, dtype=torch.bfloat16, device=dev)
        state["gen"].append(g)
    if arch.session_rope() and extended_session_ids is not None:
        # Qwen3SessionMoe/model.py:688-701: the s-th new token is rotated at max(extended_session_ids) + 1 + s
        state["rope_next"] = (extended_session_ids.max(dim=-1)[0] + 1).repeat_interleave(beams)
    own_rows = torch.arange(R, dtype=torch.int32, device=dev)

    def advance(s, prow, tokens):
        anc = state["anc"].index_select(0, prow)
        anc[:, s] = own_rows                                                 # this step's K/V is stored at its own slot
        state["anc"] = anc.contiguous()
        return decode_step(arch, pack, lut, meta, state, tokens, s)

    node0 = K.trie_init(input_ids, arch.vocab, last_bitmap, flat)
    return beam_search_device(B, beams, S, arch.vocab, flat, node0, logits0, advance, dev)


class _DecodeGraph:
    """One captured evaluation call (prefill + 4 beam steps, ~330 kernel launches) for a fixed batch shape: replaying it
    costs one launch on the host, which is what keeps small per-GPU shards (32 users per GPU at 8 GPUs) on the device's
    clock instead of the host's."""

    def __init__(self, fn, inputs):
        self.static = {k: (None if v is None else v.clone()) for k, v in inputs.items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                                        # warm-up outside the capture
            fn(**self.static)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = fn(**self.static)

    def run(self, inputs):
        for k, v in inputs.items():
            if v is not None:
                self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.out


MAX_DECODE_GRAPHS = 3


@torch.no_grad()
def constrained_beam_search(model, input_ids, attention_mask, session_ids, extended_session_ids, actions,
                            max_new_tokens, prefix_allowed_tokens_fn, candidate_trie, num_beams, num_return_sequences,
                            return_dict_in_generate=True):
    if not input_ids.is_cuda:
        raise RuntimeError("gamer_b200 has no CPU path: move the model and inputs to a CUDA device")
    arch = model.arch
    pack, lut = model._get_pack(arch)
    dev = input_ids.device
    input_ids = input_ids.contiguous()
    B, L0 = input_ids.shape
    beams, S = int(num_beams), int(max_new_tokens)
    if num_return_sequences > beams:
        raise ValueError("`num_return_sequences` has to be smaller or equal to `num_beams`.")
    if S > arch.P - 1:
        raise NotImplementedError(f"max_new_tokens={S}: decode emits one item (<= {arch.P - 1} code tokens) per call")
    flat, last_bitmap = _resolve_constraint(prefix_allowed_tokens_fn, candidate_trie, arch.vocab, arch.pad, dev)
    inputs = dict(input_ids=input_ids, attention_mask=attention_mask, session_ids=session_ids,
                  extended_session_ids=extended_session_ids, actions=actions)
    fn = lambda **kw: _decode_on_device(arch, pack, lut, flat, last_bitmap, beams, S, **kw)

    # A batch shape seen for the second time is captured into a CUDA graph and replayed from then on (the pack, the
    # trie and the bitmap keep their addresses; a weight update builds a new pack, hence a new key).
    use_graphs = getattr(model.config, "gamer_decode_graphs", True) and not torch.cuda.is_current_stream_capturing()
    key = (id(pack), id(flat), id(last_bitmap), beams, S,
           tuple((k, None if v is None else (tuple(v.shape), v.dtype)) for k, v in inputs.items()))
    cache = model.__dict__.setdefault("_decode_graphs", {})
    entry = cache.get(key) if use_graphs else None
    if isinstance(entry, _DecodeGraph):
        gen, run, err = entry.run(inputs)
        gen = gen.clone()                                                    # the graph's outputs are reused by the next replay
    else:
        if use_graphs and entry == "seen":
            for k in [k for k in cache if k[0] != id(pack)]:                 # graphs of a retired pack
                del cache[k]
            while sum(isinstance(v, _DecodeGraph) for v in cache.values()) >= MAX_DECODE_GRAPHS:
                del cache[next(k for k, v in cache.items() if isinstance(v, _DecodeGraph))]
            cache[key] = _DecodeGraph(fn, inputs)
            gen, run, err = cache[key].run(inputs)
            gen = gen.clone()
        else:
            if use_graphs:
                cache[key] = "seen"
            gen, run, err = fn(**inputs)
    _raise_on(int(err.item()))                                               # single host sync of the whole decode
    K.check_token_ids(input_ids.device)                                      # (already synchronised: a second word read)
    seqs = torch.cat([input_ids.view(B, 1, L0).expand(B, beams, L0), gen], dim=2)
    scores = run / float(S)                                                  # length_penalty = 1: sum / generated length
    nret = num_return_sequences
    seqs = seqs[:, :nret].reshape(B * nret, L0 + S)
    scores = scores[:, :nret].reshape(B * nret)
    out = BeamSearchOutput(sequences=seqs, sequences_scores=scores, generated=gen[:, :nret].reshape(B * nret, S))
    return out if return_dict_in_generate else seqs
