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
