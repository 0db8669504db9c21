"""HBM roofline of the K1 kernels (fused embedding gather + router indices; sparse embedding gradient) at the
long-history sweep size of BASELINE.json configs[4] (max_his_len=500: L = 2505 tokens), where launch latency is
amortised.  Prints one JSON line per kernel: achieved GB/s on the algorithmic bytes (8 B id + 512 B
bf16 row + 12 B indices per token forward; 4 B sorted index + 512 B row per token + the 1.07 MB fp32 table backward)
against MEASURED_PEAKS.json.

    python tools/embed_bench.py [--batch 512] [--his 500] [--iters 20]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gamer_b200 import kernels as K          # noqa: E402
from gamer_b200 import synthetic as syn      # noqa: E402


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


def run(batch=512, his=500, iters=20, dev="cuda:0"):
    """The K1 kernels alone at one launch of batch x 5 (his + 1) tokens; returns one dict per kernel."""
    a = argparse.Namespace(batch=batch, his=his, iters=iters)
    L = 5 * (a.his + 1)
    V, H = syn.VOCAB, 256
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        src = "measured (MEASURED_PEAKS.json)"
    except Exception:
        peak, src = 6650.0, "fallback (B200_PROFILING.md)"
    g = torch.Generator().manual_seed(0)
    # ShortVideoAD-shaped ids: items of 5 tokens (behaviour token + 4 codes), a few right-padded rows
    codes = torch.stack([torch.randint(lo, lo + 256, (a.batch, a.his + 1), generator=g) for lo in (syn.A0, syn.B0, syn.C0, syn.D0)], -1)
    beh = torch.randint(526, 529, (a.batch, a.his + 1, 1), generator=g)
    ids = torch.cat([beh, codes], -1).view(a.batch, L)
    ids[: a.batch // 8, -50:] = syn.PAD
    ids = ids.to(dev)
    table = (0.02 * torch.randn(V, H, generator=g)).to(torch.bfloat16).to(dev)
    lut = torch.arange(V, dtype=torch.int32)
    for i, t in enumerate((526, 527, 528)):
        lut[t] = i + 1
    lut = lut.to(dev)
    M = a.batch * L
    ms_f = timed(lambda: K.embed_route(ids, table, lut, 3, 5, syn.PAD, syn.EOS), a.iters)
    bytes_f = M * (8 + 2 * H + 12)
    dx = torch.randn(M, H, device=dev).to(torch.bfloat16)
    sort_buf = K.embed_sort(ids.view(-1), V, syn.PAD)
    dtab = torch.zeros(V, H, dtype=torch.float32, device=dev)
    ms_b = timed(lambda: K.embed_bwd(dx, V, sort_buf, dtab), a.iters)
    bytes_b = M * (4 + 2 * H) + V * H * 4
    ms_s = timed(lambda: K.embed_sort(ids.view(-1), V, syn.PAD), a.iters)
    out = []
    for name, ms, nbytes in (("gamer_embed_route_fwd", ms_f, bytes_f), ("gamer_embed_bwd", ms_b, bytes_b),
                             ("gamer_embed_sort_build (ids only: 8 B read + 4 B written per token)", ms_s, M * 12)):
        gbs = nbytes / (ms / 1e3) / 1e9
        out.append({"kernel": name, "tokens": M, "batch": a.batch, "seq_len": L, "ms": ms, "algorithmic_bytes": nbytes,
                    "achieved_gbs": gbs, "peak_gbs": peak, "frac": gbs / peak, "peak_source": src})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--his", type=int, default=500)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    for row in run(a.batch, a.his, a.iters):
        print(json.dumps(row))


if __name__ == "__main__":
    main()
