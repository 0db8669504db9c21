"""Masked-attention roofline sweep (BASELINE.json configs[4]): forward and backward TFLOP/s of the tcgen05 kernels on the
causal pair count (4*64*n_q*L(L+1)/2 FLOP per sequence forward, 2.5x backward) for the four mask
kinds at L = 505 / 1005 / 2505, full-length rows, with and without dropout.  One JSON line per point.

    python tools/attn_bench.py [--tokens 65536] [--iters 10]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gamer_b200 import kernels as K          # noqa: E402


def timed(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=65536, help="tokens per call (batch = tokens // L)")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--his", type=int, nargs="*", default=[100, 200, 500], help="history lengths (L = 5 (his + 1))")
    ap.add_argument("--kinds", type=int, nargs="*", default=[0, 1, 2, 3])
    a = ap.parse_args()
    dev = "cuda:0"
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    except Exception:
        peak = 1400.0
    nq, nkv, hd = 6, 3, 64
    g = torch.Generator().manual_seed(0)
    for his in a.his:
        L = 5 * (his + 1)
        B = max(1, a.tokens // L)
        am = torch.ones(B, L, dtype=torch.int32, device=dev)
        act = torch.randint(0, 3, (B, his + 1), generator=g).repeat_interleave(5, dim=1).to(torch.int32).to(dev).contiguous()
        sess = torch.cumsum((torch.rand(B, his + 1, generator=g) < 0.12).long(), 1).repeat_interleave(5, dim=1).to(torch.int32).to(dev).contiguous()
        qkv = torch.randn(B * L, 768, device=dev).to(torch.bfloat16)
        d_o = torch.randn(B * L, nq * hd, device=dev).to(torch.bfloat16)
        dqkv = torch.empty_like(qkv)
        for kind in a.kinds:
            for p in (0.0, 0.2):
                drop = K.Dropout(1234, 0, 8 + kind, p) if p > 0 else None
                o, lse, _, keep = K.attn_fwd(qkv, B, L, nq, nkv, hd, kind, 5, am, act, sess, hd ** -0.5, drop=drop)
                ms_f = timed(lambda: K.attn_fwd(qkv, B, L, nq, nkv, hd, kind, 5, am, act, sess, hd ** -0.5, drop=drop), a.iters)
                ms_b = timed(lambda: K.attn_bwd(qkv, o, d_o, lse, B, L, nq, nkv, hd, kind, 5, am, act, sess, hd ** -0.5, dqkv,
                                                drop=drop, keep=keep), a.iters)
                flops = 4 * hd * nq * B * L * (L + 1) // 2
                print(json.dumps({"kind": kind, "L": L, "batch": B, "dropout": p, "fwd_ms": ms_f, "bwd_ms": ms_b,
                                  "fwd_tflops": flops / ms_f / 1e9, "bwd_tflops": 2.5 * flops / ms_b / 1e9,
                                  "fwd_frac": flops / ms_f / 1e9 / peak, "bwd_frac": 2.5 * flops / ms_b / 1e9 / peak,
                                  "peak_tflops": peak}))


if __name__ == "__main__":
    main()
