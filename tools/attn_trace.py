"""In-kernel timeline of CTA 0 of the attention kernels (debug hook gamer_attn_set_trace): per role, the clock at each
hand-off point.  Prints per-step durations between consecutive tags.

    python tools/attn_trace.py [--kind 0] [--p 0.2] [--which bwd|fwd]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gamer_b200 import kernels as K          # noqa: E402
from gamer_b200._cabi import lib             # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", type=int, default=0)
    ap.add_argument("--p", type=float, default=0.2)
    ap.add_argument("--L", type=int, default=505)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--which", default="bwd")
    ap.add_argument("--cap", type=int, default=400)
    a = ap.parse_args()
    dev = "cuda:0"
    nq, nkv, hd = 6, 3, 64
    B, L = a.batch, a.L
    g = torch.Generator().manual_seed(0)
    items = (L + 4) // 5
    am = torch.ones(B, L, dtype=torch.int32, device=dev)
    act = torch.randint(0, 3, (B, items), generator=g).repeat_interleave(5, dim=1)[:, :L].to(torch.int32).to(dev).contiguous()
    sess = torch.cumsum((torch.rand(B, items, generator=g) < 0.12).long(), 1).repeat_interleave(5, dim=1)[:, :L]
    sess = sess.to(torch.int32).to(dev).contiguous()
    qkv = torch.randn(B * L, 768, device=dev).to(torch.bfloat16)
    d_o = torch.randn(B * L, nq * hd, device=dev).to(torch.bfloat16)
    dqkv = torch.empty_like(qkv)
    drop = K.Dropout(1234, 0, 8 + a.kind, a.p) if a.p > 0 else None
    run_f = lambda: K.attn_fwd(qkv, B, L, nq, nkv, hd, a.kind, 5, am, act, sess, hd ** -0.5, drop=drop)
    o, lse, _, keep = run_f()
    run_b = lambda: K.attn_bwd(qkv, o, d_o, lse, B, L, nq, nkv, hd, a.kind, 5, am, act, sess, hd ** -0.5, dqkv, drop=drop,
                               keep=keep)
    run_b()
    torch.cuda.synchronize()
    buf = torch.zeros(4, a.cap, 2, dtype=torch.int64, device=dev)
    lib().gamer_attn_set_trace(buf.data_ptr(), a.cap)
    (run_b if a.which == "bwd" else run_f)()
    torch.cuda.synchronize()
    lib().gamer_attn_set_trace(None, 0)
    t = buf.cpu()
    t0 = min(int(t[r, 0, 1]) for r in range(4) if int(t[r, 0, 0]) != 0)
    for r in range(4):
        ev = [(int(t[r, i, 0]), int(t[r, i, 1]) - t0) for i in range(a.cap) if int(t[r, i, 0]) != 0]
        if not ev:
            continue
        print(f"role {r}: {len(ev)} events")
        line = []
        for i, (tag, clk) in enumerate(ev[:160]):
            d = clk - ev[i - 1][1] if i else 0
            line.append(f"{tag}@{clk}(+{d})")
        print("  " + " ".join(line))


if __name__ == "__main__":
    main()
