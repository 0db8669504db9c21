"""CPU: the tensor form of the ranking metrics equals the reference-shaped string functions on random beam outputs
(multi-target users, repeated hits, users without hits)."""
import random

import torch

from gamer_b200 import ranking as R

METRICS = ["hit@1", "hit@5", "hit@10", "recall@1", "recall@5", "recall@10", "ndcg@5", "ndcg@10"]


def test_tensor_metrics_equal_string_metrics():
    rng = random.Random(0)
    B, K, S = 64, 10, 4
    gen = torch.randint(0, 6, (B, K, S))
    targets = []
    for b in range(B):
        tl = [tuple(int(x) for x in torch.randint(0, 6, (S,))) for _ in range(rng.randint(1, 5))]
        for _ in range(rng.randint(0, 3)):                      # plant some of the user's own beams among the targets
            tl.append(tuple(int(x) for x in gen[b, rng.randrange(K)]))
        targets.append(tl)
    pred = ["".join(f"<{int(t)}>" for t in gen[b, k]) for b in range(B) for k in range(K)]
    scores = [float(K - k) for _ in range(B) for k in range(K)]                  # already best-first
    tgt = [["".join(f"<{t}>" for t in tup) for tup in tl] for tl in targets]
    hits_ref = R.get_topk_results(pred, scores, tgt, K)
    want = R.get_metrics_results(hits_ref, METRICS, tgt)
    tt, cnt = R.pack_targets(targets, S)
    hits = R.topk_hits(gen, tt)
    assert hits.tolist() == hits_ref
    got = R.metric_sums(hits, cnt, METRICS)
    for m in METRICS:
        assert abs(float(got[m]) - want[m]) < 1e-9, (m, float(got[m]), want[m])
