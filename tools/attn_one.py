"""One attention configuration, forward + backward a few times (the ncu target for the attention kernels).

    python tools/attn_one.py --kind 0 --p 0.2 --L 505 --batch 128 --iters 3
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gamer_b200 import kernels as K          # noqa: E402
from gamer_b200 import synthetic as syn      # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", type=int, default=0)
    ap.add_argument("--p", type=float, default=0.2)
    ap.add_argument("--L", type=int, default=505)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--bench-levels", action="store_true", help="behaviour levels drawn as the bench does (85/12/3 %)")
    a = ap.parse_args()
    dev = "cuda:0"
    nq, nkv, hd = 6, 3, 64
    B, L = a.batch, a.L
    g = torch.Generator().manual_seed(0)
    items = (L + 4) // 5
    am = torch.ones(B, L, dtype=torch.int32, device=dev)
    if a.bench_levels:
        lv = torch.multinomial(torch.tensor(syn.BEHAVIOR_PROB), B * items, replacement=True, generator=g).view(B, items)
    else:
        lv = torch.randint(0, 3, (B, items), generator=g)
    act = lv.repeat_interleave(5, dim=1)[:, :L].to(torch.int32).to(dev).contiguous()
    sess = torch.cumsum((torch.rand(B, items, generator=g) < 0.12).long(), 1).repeat_interleave(5, dim=1)[:, :L]
    sess = sess.to(torch.int32).to(dev).contiguous()
    qkv = torch.randn(B * L, 768, device=dev).to(torch.bfloat16)
    d_o = torch.randn(B * L, nq * hd, device=dev).to(torch.bfloat16)
    dqkv = torch.empty_like(qkv)
    drop = K.Dropout(1234, 0, 8 + a.kind, a.p) if a.p > 0 else None
    for _ in range(a.iters):
        o, lse, _, keep = K.attn_fwd(qkv, B, L, nq, nkv, hd, a.kind, 5, am, act, sess, hd ** -0.5, drop=drop)
        K.attn_bwd(qkv, o, d_o, lse, B, L, nq, nkv, hd, a.kind, 5, am, act, sess, hd ** -0.5, dqkv, drop=drop, keep=keep)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
