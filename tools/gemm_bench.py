"""Per-shape throughput of the tcgen05 GEMM entry points at the training shapes (M = 64 640 token rows): TFLOP/s and
the HBM-side GB/s of every projection of one layer, forward and backward.  One JSON line per shape.

    python tools/gemm_bench.py [--rows 64640] [--iters 20]
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
    for _ in range(3):
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
    ap.add_argument("--rows", type=int, default=64640)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    dev = "cuda:0"
    M = a.rows
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    except Exception:
        peak = 1400.0
    shapes = [("qkv fwd", 768, 256, False), ("cross qkv+gate fwd", 1024, 256, False), ("o_proj fwd (+resid)", 256, 384, True),
              ("gate|up fwd", 1024, 256, False), ("gate|up fwd (inject)", 1024, 320, False), ("down dgrad", 512, 256, False),
              ("o_proj dgrad", 384, 256, False), ("qkv dgrad", 256, 768, False), ("gate|up dgrad", 256, 1024, False),
              ("gate|up dgrad (inject)", 320, 1024, False), ("down fwd (plain)", 256, 512, True)]
    for name, N, Kd, resid in shapes:
        A = torch.randn(M, Kd, device=dev).to(torch.bfloat16)
        B = (0.05 * torch.randn(N, Kd, device=dev)).to(torch.bfloat16)
        R = torch.randn(M, N, device=dev).to(torch.bfloat16) if resid else None
        out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        ms = timed(lambda: K.gemm_tn(A, B, N, out=out, resid=R), a.iters)
        flops = 2.0 * M * N * Kd
        nbytes = M * Kd * 2 + N * Kd * 2 + M * N * 2 * (2 if resid else 1)
        print(json.dumps({"gemm": name, "rows": M, "N": N, "K": Kd, "ms": ms, "tflops": flops / ms / 1e9,
                          "frac_of_peak": flops / ms / 1e9 / peak, "hbm_gbs": nbytes / ms / 1e6}))
    # weight gradients
    for name, N, Kd in [("qkv wgrad", 768, 256), ("gate|up wgrad", 1024, 256), ("o_proj wgrad", 256, 384), ("down wgrad", 256, 512)]:
        dY = torch.randn(M, N, device=dev).to(torch.bfloat16)
        X = torch.randn(M, Kd, device=dev).to(torch.bfloat16)
        dW = torch.zeros(1, N, Kd, dtype=torch.float32, device=dev)
        ms = timed(lambda: K.gemm_wgrad(dY, X, N, Kd, dW), a.iters)
        flops = 2.0 * M * N * Kd
        print(json.dumps({"gemm": name, "rows": M, "N": N, "K": Kd, "ms": ms, "tflops": flops / ms / 1e9,
                          "frac_of_peak": flops / ms / 1e9 / peak, "hbm_gbs": (M * (N + Kd) * 2) / ms / 1e6}))


if __name__ == "__main__":
    main()
