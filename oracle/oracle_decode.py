"""TEST INFRASTRUCTURE ONLY — CPU restatement of the constrained beam-search decode and the ranking metrics.

Follows (paths relative to /root/reference/SeqRec/):
  * generation/trie.py:5-104 — prefix trie + `prefix_allowed_tokens_fn_by_last_token`;
  * third-party HF `GenerationMixin._beam_search` + `PrefixConstrainedLogitsProcessor` as driven by
    tasks/test_SMB_decoder.py:159-177 (transformers==4.51.0, un-vendored; algorithm read from the installed 5.5.0
    copy, generation/utils.py:2945-3370): full-vocab fp32 log-softmax, -inf on non-children *after* normalisation
    (Q6), + running beam score (init [0,-1e9,...]), top-2K over K*V, keep top-K, final score = sum/gen_len;
  * evaluation/ranking.py:5-90 — hit/recall/ndcg@K.
Pinned against the reference's own `generate()` output in tests/golden/ (made by oracle/make_golden.py).
"""
from __future__ import annotations

import math

import torch

from . import oracle_model as om


class PrefixTree:
    """Nested-dict trie over token lists (generation/trie.py:5-80)."""

    def __init__(self, sequences=()):
        self.root: dict = {}
        self.count = 0
        for s in sequences:
            self.add(s)

    def add(self, seq):
        node = self.root
        for t in seq:
            node = node.setdefault(int(t), {})
        self.count += 1

    def children(self, prefix) -> list[int]:
        node = self.root
        for t in prefix:
            node = node.get(int(t))
            if node is None:
                return []
        return list(node.keys())


def allowed_by_last_token(tree: PrefixTree, last_token_set: set[int], sentence: list[int]) -> list[int]:
    """generation/trie.py:92-104: suffix after the last token that ends an item (or pad), walked in the trie."""
    i = len(sentence) - 1
    while i >= 0 and sentence[i] not in last_token_set:
        i -= 1
    return tree.children(sentence[i + 1:])


def constrained_beam_search(spec: om.Spec, W: dict, tree: PrefixTree, last_token_set: set[int], input_ids,
                            attention_mask, session_ids=None, extended_session_ids=None, actions=None,
                            num_beams: int = 20, max_new_tokens: int = 4, trace: list | None = None):
    """Returns (sequences [B*K, L+new] int64, sequences_scores [B*K] fp32), best-first per user — the same
    contract as `generate(..., num_beams=K, num_return_sequences=K, output_scores=True)`.

    Every user's prompt is expanded to K identical rows (as HF does) so cache handling is a plain row gather.
    `trace` (test aid), when a list, receives per step (kept sequences [B,K,L+s+1], kept running scores [B,K],
    running score of the best pruned candidate [B]) so a test can tell how close a hypothesis came to the pruning cut.
    """
    B, L = input_ids.shape
    K, V = num_beams, spec.vocab_size
    rep = lambda t: None if t is None else t.repeat_interleave(K, dim=0)
    logits, st = om.prefill(spec, W, rep(input_ids), rep(attention_mask), rep(session_ids),
                            rep(extended_session_ids), rep(actions))
    running = torch.zeros(B, K)
    running[:, 1:] = -1e9
    seqs = rep(input_ids).view(B, K, L)
    for step in range(max_new_tokens):
        logp = torch.log_softmax(logits.float(), dim=-1)                       # [B*K, V]
        mask = torch.full_like(logp, -math.inf)
        flat = seqs.view(B * K, -1)
        for r in range(B * K):
            allowed = allowed_by_last_token(tree, last_token_set, flat[r].tolist())
            if not allowed:
                raise ValueError("prefix_allowed_tokens_fn returned an empty list")
            mask[r, allowed] = 0
        acc = (logp + mask).view(B, K, V) + running[:, :, None]
        top_val, top_idx = torch.topk(acc.view(B, K * V), k=2 * K)             # HF keeps max(2,1+n_eos)*K
        first_pruned = top_val[:, K].clone()
        top_val, top_idx = top_val[:, :K], top_idx[:, :K]                      # nothing can finish early: top-K
        beam = top_idx // V
        tok = top_idx % V
        seqs = torch.cat([torch.gather(seqs, 1, beam[:, :, None].expand(-1, -1, seqs.shape[2])),
                          tok[:, :, None]], dim=2)
        running = top_val
        if trace is not None:
            trace.append((seqs.clone(), running.clone(), first_pruned))
        if step + 1 < max_new_tokens:
            rows = (beam + torch.arange(B)[:, None] * K).view(-1)
            st.reorder(rows)
            logits = om.decode_step(spec, W, st, tok.reshape(-1))
    scores = running / float(max_new_tokens)                                   # length_penalty = 1
    return seqs.reshape(B * K, -1), scores.reshape(-1)


def teacher_forced_scores(spec: om.Spec, W: dict, input_ids, attention_mask, gen_tokens, session_ids=None,
                          extended_session_ids=None, actions=None):
    """Score given continuations with the cached path: mean full-vocabulary log-probability of `gen_tokens` [R, S]
    appended to `input_ids` [R, L] (= what beam search reports for a hypothesis that survives)."""
    logits, st = om.prefill(spec, W, input_ids, attention_mask, session_ids, extended_session_ids, actions)
    R, S = gen_tokens.shape
    total = torch.zeros(R)
    for s in range(S):
        logp = torch.log_softmax(logits.float(), dim=-1)
        total = total + logp.gather(1, gen_tokens[:, s:s + 1]).squeeze(1)
        if s + 1 < S:
            logits = om.decode_step(spec, W, st, gen_tokens[:, s])
    return total / float(S)


# ------------------------------------------------------------------------------------------------------------
# evaluation/ranking.py restated on id tuples (the reference compares decoded strings; ids are a bijection)
# ------------------------------------------------------------------------------------------------------------
def hit_lists(pred_items, scores, targets, k: int):
    """pred_items: list (B*k) of hashables; targets: list (B) of lists.  ranking.py:5-32 (stable sort by score)."""
    out = []
    for b in range(len(targets)):
        pairs = list(zip(pred_items[b * k:(b + 1) * k], [float(s) for s in scores[b * k:(b + 1) * k]]))
        pairs.sort(key=lambda x: x[1], reverse=True)
        out.append([1 if p in targets[b] else 0 for p, _ in pairs])
    return out


def metrics(hits, targets, names):
    """ranking.py:35-90: sums (not means) over the rows of `hits`."""
    res = {}
    for m in names:
        kind, k = m.split("@")
        k = int(k)
        tot = 0.0
        for row, tgt in zip(hits, targets):
            nT = len(set(tgt))
            r = row[:k]
            if kind.lower() == "hit":
                tot += 1.0 if sum(r) > 0 else 0.0
            elif kind.lower() == "recall":
                tot += min(sum(r), nT) / nT
            elif kind.lower() == "ndcg":
                dcg, cnt = 0.0, 0
                for j, h in enumerate(r):
                    cnt += h
                    dcg += h / math.log(j + 2, 2)
                    if cnt == nT:
                        break
                idcg = sum(1 / math.log(j + 2, 2) for j in range(min(k, nT)))
                tot += dcg / idcg
            else:
                raise NotImplementedError(m)
        res[m] = tot
    return res
