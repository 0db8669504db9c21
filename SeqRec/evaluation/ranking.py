"""Reference path SeqRec/evaluation/ranking.py -> gamer_b200.ranking."""
from gamer_b200.ranking import get_metrics_results, get_topk_results, hit_k, ndcg_k, recall_k  # noqa: F401
