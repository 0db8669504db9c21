"""One constrained-beam-search evaluation call at the bench's eval shape (the ncu target for the decode kernels).

    python tools/eval_one.py --users 256 --iters 2
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                   # noqa: E402
from gamer_b200 import modeling                # noqa: E402
from gamer_b200 import synthetic as syn        # noqa: E402
from gamer_b200.trie import flat_from_array, prefix_allowed_tokens_fn_by_last_token   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=256)
    ap.add_argument("--iters", type=int, default=2)
    a = ap.parse_args()
    cfg = bench.eval_config(100)
    cfg.gamer_decode_graphs = False            # eager launches: every kernel visible to ncu one by one
    torch.manual_seed(43)
    m = modeling.Qwen3SessionMoeWithTemperature(cfg).cuda().eval()
    cat = syn.make_catalogue(bench.EVAL_TRIE_ITEMS, 1234)
    items = cat.item_sequences(2)
    fn = prefix_allowed_tokens_fn_by_last_token(flat_from_array(items), set(int(t) for t in items[:, -1]) | {syn.PAD})
    batch, _ = syn.make_eval_batch(cat, a.users, max_his_len=100, target_behavior=2, seed=77, full_length=True)
    b = {k: v.cuda() for k, v in batch.items()}
    for _ in range(a.iters):
        m.generate(**b, max_new_tokens=4, prefix_allowed_tokens_fn=fn, num_beams=bench.EVAL_BEAMS,
                   num_return_sequences=bench.EVAL_BEAMS)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
