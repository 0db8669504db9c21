"""One optimizer-step's gradient at the headline shape, accumulated as 1 x 1024 rows and as 2 x 512 rows (dropout off):
the two flat gradient buffers must agree to accumulation-order noise.  Guards the index arithmetic of every kernel at the
larger launch (1024 x 505 = 517k token rows).   python tools/mb_check.py [--rows 1024]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from gamer_b200 import engine as E
from gamer_b200 import modeling
from gamer_b200 import synthetic as syn
from gamer_b200.trainer import NativeTrainer


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1024)
    ap.add_argument("--max-his-len", type=int, default=100)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(42)
    cfg = bench.model_config(a.max_his_len)
    cfg.dropout_rate = 0.0
    cfg.attention_dropout = 0.0
    model = modeling.Qwen3MultiWithTemperature(cfg)
    model.set_hyper(0.7)
    model = model.to(dev).train()
    tr = NativeTrainer(model, lr=0.0, use_cuda_graphs=False)
    cat = syn.make_catalogue(250_000, 1234)
    batch = {k: v.to(dev) for k, v in syn.make_train_batch(cat, a.rows, max_his_len=a.max_his_len, seed=7).items()}
    inv = (1.0 / (E.shift_labels(batch["labels"]) != -100).sum().clamp(min=1).float()).reshape(1)
    grads, losses = {}, {}
    for mb in (a.rows, a.rows // 2):
        tr.flat_g.zero_()
        tot = 0.0
        for b0 in range(0, a.rows, mb):
            sub = {k: v[b0:b0 + mb] for k, v in batch.items()}
            tot = tot + tr.forward_backward(sub, inv, last_micro=(b0 + mb >= a.rows))
        torch.cuda.synchronize()
        grads[mb] = tr.flat_g.clone()
        losses[mb] = float(tot)
    g1, g2 = grads[a.rows].double(), grads[a.rows // 2].double()
    rel = ((g1 - g2).norm() / g2.norm()).item()
    cos = (torch.dot(g1, g2) / (g1.norm() * g2.norm())).item()
    print({"rows": a.rows, "loss_1x": losses[a.rows], "loss_2x": losses[a.rows // 2], "rel_l2": rel, "cosine": cos,
           "finite": bool(torch.isfinite(g1).all())})
    assert abs(losses[a.rows] - losses[a.rows // 2]) <= 1e-3 * abs(losses[a.rows // 2])
    assert rel <= 2e-2 and cos >= 0.9995


if __name__ == "__main__":
    main()
