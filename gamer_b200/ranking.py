"""Ranking metrics of the evaluation task (reference: SeqRec/evaluation/ranking.py:5-90), same function names,
arguments and return values: hit lists per user from the beam output, then hit@K / recall@K / ndcg@K SUMS over users
(multi-target aware).  Host-side by design: it consumes K strings per user."""
from __future__ import annotations

import math


def get_topk_results(predictions, scores, targets, k):
    """predictions: B*k decoded strings; scores: B*k beam scores; targets: per user a string or a list of strings."""
    cleaned = [p.split("Response:")[-1].strip().replace(" ", "") for p in predictions]
    results = []
    for b in range(len(targets)):
        ranked = sorted(zip(cleaned[b * k:(b + 1) * k], scores[b * k:(b + 1) * k]), key=lambda pair: pair[1], reverse=True)
        tgt = targets[b]
        if isinstance(tgt, list):
            results.append([1 if pred in tgt else 0 for pred, _ in ranked])
        else:
            results.append([1 if pred == tgt else 0 for pred, _ in ranked])
    return results


def ndcg_k(topk_results, k, targets=None):
    total = 0.0
    for i, row in enumerate(topk_results):
        dcg, found = 0.0, 0
        for j, hit in enumerate(row[:k]):
            found += 1 if hit == 1 else 0
            dcg += hit / math.log(j + 2, 2)
            if (found == 1 and targets is None) or (targets is not None and found == len(targets[i])):
                break
        if targets is not None:
            ideal = sum(1 / math.log(j + 2, 2) for j in range(min(k, len(targets[i]))))
            assert ideal > 0, "Ideal DCG should be greater than 0"
            dcg /= ideal
        total += dcg
    return total


def recall_k(topk_results, k, targets=None):
    total = 0.0
    for i, row in enumerate(topk_results):
        hits = sum(row[:k])
        total += hits if targets is None else min(hits, len(targets[i])) / len(targets[i])
    return total


def hit_k(topk_results, k):
    return float(sum(1 for row in topk_results if sum(row[:k]) > 0))


def get_metrics_results(topk_results, metrics, targets=None):
    target_sets = [set(t) for t in targets] if targets is not None else None
    out = {}
    for m in metrics:
        kind, k = m.lower().split("@")[0], int(m.split("@")[1])
        if kind.startswith("hit"):
            out[m] = hit_k(topk_results, k)
        elif kind.startswith("ndcg"):
            out[m] = ndcg_k(topk_results, k, target_sets)
        elif kind.startswith("recall"):
            out[m] = recall_k(topk_results, k, target_sets)
        else:
            raise NotImplementedError
    return out


# ----------------------------------------------------------------------------------------------------------------------
# Tensor form of the same metrics: the evaluation loop of tasks/test_SMB_decoder.py:196-283
# detokenises K strings per user, compares strings on the host and pickles hit lists between ranks; here the decoded
# code tuples are compared as integers on the device they were decoded on, and only the metric sums leave it.
# ----------------------------------------------------------------------------------------------------------------------
def pack_targets(targets, width, device=None):
    """Per-user lists of target tuples -> (padded [B, T_max, width] int64 tensor filled with -1, counts [B])."""
    import torch
    B = len(targets)
    t_max = max(1, max(len(t) for t in targets))
    out = torch.full((B, t_max, width), -1, dtype=torch.int64)
    cnt = torch.zeros(B, dtype=torch.int64)
    for b, tl in enumerate(targets):
        uniq = sorted(set(tuple(int(x) for x in t) for t in tl))      # the reference compares against set(targets)
        cnt[b] = len(uniq)
        for j, t in enumerate(uniq):
            out[b, j] = torch.tensor(t, dtype=torch.int64)
    return out.to(device), cnt.to(device)


def topk_hits(generated, targets):
    """generated [B, K, S] int64 (best first), targets [B, T, S] (-1 padded) -> hits [B, K] (1 where the beam's tuple is
    one of the user's targets) — get_topk_results on id tuples."""
    eq = (generated[:, :, None, :] == targets[:, None, :, :]).all(-1)          # [B, K, T]
    return eq.any(-1).to(generated.dtype)


def metric_sums(hits, n_targets, metrics):
    """hits [B, K], n_targets [B] -> {metric: SUM over users} as device scalars; same definitions as hit_k / recall_k /
    ndcg_k above (multi-target: recall = min(hits, nT) / nT; ndcg stops counting once every target is found and is
    normalised by the ideal DCG of min(k, nT) hits)."""
    import torch
    out = {}
    B, K = hits.shape
    disc = 1.0 / torch.log2(torch.arange(K, device=hits.device, dtype=torch.float64) + 2.0)
    nt = n_targets.clamp(min=1)
    for m in metrics:
        kind, k = m.lower().split("@")[0], int(m.split("@")[1])
        h = hits[:, :k].to(torch.float64)
        if kind.startswith("hit"):
            out[m] = (h.sum(1) > 0).to(torch.float64).sum()
        elif kind.startswith("recall"):
            out[m] = (torch.minimum(h.sum(1), nt.to(torch.float64)) / nt).sum()
        elif kind.startswith("ndcg"):
            found_before = torch.cumsum(h, 1) - h                               # hits strictly before rank j
            live = found_before < nt.unsqueeze(1)                               # the reference breaks once all are found
            dcg = (h * live * disc[:k]).sum(1)
            ideal_len = torch.minimum(nt, torch.tensor(k, device=hits.device))
            ideal = torch.cumsum(disc[:k], 0)[(ideal_len - 1).clamp(min=0)]
            out[m] = (dcg / ideal).sum()
        else:
            raise NotImplementedError(m)
    return out
