"""GPU diagnostic for the attention kernels: per-head / per-tile errors against the fp32 oracle, plus timings.
    python tools/attn_debug.py [--time]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gamer_b200 import kernels as k          # noqa: E402
from oracle import oracle_model as om        # noqa: E402
from tests.test_kernels_gpu import _attn_inputs, bf, rel_err   # noqa: E402

DEV = "cuda:0"


def check(kind, B, L, left_pad, seed=0, verbose=True):
    torch.manual_seed(seed)
    nq, nkv, hd = 6, 3, 64
    M = B * L
    am, act, sess = _attn_inputs(B, L, 11 + kind, left_pad)
    qkv = bf(torch.randn(M, 768, device=DEV))
    scale = hd ** -0.5
    i32 = lambda t: t.to(torch.int32).to(DEV).contiguous()
    o, lse, vmean = k.attn_fwd(qkv, B, L, nq, nkv, hd, kind, 5, i32(am), i32(act), i32(sess), scale)
    torch.cuda.synchronize()
    allow = om.allow_matrix(kind, am, act, sess, 5).to(DEV)
    qf = qkv.float().requires_grad_(True)
    q = qf[:, :384].view(B, L, nq, hd).transpose(1, 2)
    kk = qf[:, 384:576].view(B, L, nkv, hd).transpose(1, 2)
    v = qf[:, 576:].view(B, L, nkv, hd).transpose(1, 2)
    ref = om.masked_attention(q, kk, v, allow, scale).transpose(1, 2).reshape(M, nq * hd)
    uni = ~allow.any(-1)
    msg = [f"kind={kind} B={B} L={L} left_pad={left_pad}: fwd {rel_err(o, ref):.3e} uni_ok={torch.equal(torch.isinf(lse[:, 0, :]), uni)}"]
    if verbose and rel_err(o, ref) > 8e-3:
        ov, rv = o.float().view(B, L, nq, hd), ref.view(B, L, nq, hd)
        for h in range(nq):
            for t in range(0, L, 128):
                e = rel_err(ov[:, t:t + 128, h], rv[:, t:t + 128, h])
                msg.append(f"   fwd head {h} rows {t}: {e:.3e}")
    d_o = bf(torch.randn(M, nq * hd, device=DEV))
    ref.backward(d_o.float())
    dqkv = torch.zeros(M, 768, dtype=torch.bfloat16, device=DEV)
    k.attn_bwd(qkv, o, d_o, lse, B, L, nq, nkv, hd, kind, 5, i32(am), i32(act), i32(sess), scale, dqkv)
    torch.cuda.synchronize()
    g = qf.grad
    for name, sl in (("dq", slice(0, 384)), ("dk", slice(384, 576)), ("dv", slice(576, 768))):
        e = rel_err(dqkv[:, sl], g[:, sl])
        msg.append(f"   {name} {e:.3e}")
        if verbose and e > 2e-2:
            mv, gv = dqkv[:, sl].float().view(B, L, -1, hd), g[:, sl].view(B, L, -1, hd)
            for h in range(mv.shape[2]):
                for t in range(0, L, 128):
                    msg.append(f"      {name} head {h} rows {t}: {rel_err(mv[:, t:t + 128, h], gv[:, t:t + 128, h]):.3e}")
    print("\n".join(msg), flush=True)


def timeit(kind, B=128, L=505, iters=10):
    nq, nkv, hd = 6, 3, 64
    M = B * L
    am, act, sess = _attn_inputs(B, L, 3, False)
    am[:] = 1
    act = torch.where(act == 100, torch.zeros_like(act), act)
    qkv = bf(torch.randn(M, 768, device=DEV))
    i32 = lambda t: t.to(torch.int32).to(DEV).contiguous()
    a, c, s = i32(am), i32(act), i32(sess)
    scale = hd ** -0.5
    d_o = bf(torch.randn(M, nq * hd, device=DEV))
    dqkv = torch.zeros(M, 768, dtype=torch.bfloat16, device=DEV)
    for _ in range(3):
        o, lse, _ = k.attn_fwd(qkv, B, L, nq, nkv, hd, kind, 5, a, c, s, scale)
        k.attn_bwd(qkv, o, d_o, lse, B, L, nq, nkv, hd, kind, 5, a, c, s, scale, dqkv)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize()
    e[0].record()
    for _ in range(iters):
        o, lse, _ = k.attn_fwd(qkv, B, L, nq, nkv, hd, kind, 5, a, c, s, scale)
    e[1].record()
    for _ in range(iters):
        k.attn_bwd(qkv, o, d_o, lse, B, L, nq, nkv, hd, kind, 5, a, c, s, scale, dqkv)
    e[2].record()
    torch.cuda.synchronize()
    fl = 4 * hd * nq * B * L * (L + 1) // 2
    tf, tb = e[0].elapsed_time(e[1]) / iters, e[1].elapsed_time(e[2]) / iters
    print(f"time kind={kind} B={B} L={L}: fwd {tf * 1e3:.1f} us ({fl / tf / 1e9:.1f} TFLOP/s)  bwd {tb * 1e3:.1f} us "
          f"({2.5 * fl / tb / 1e9:.1f} TFLOP/s)", flush=True)


if __name__ == "__main__":
    cases = [(0, 2, 65, False), (0, 2, 505, False), (1, 2, 505, False), (2, 2, 300, True), (3, 2, 200, True),
             (0, 3, 129, True), (1, 3, 37, False)]
    for c in ([] if "--time-only" in sys.argv else cases):
        try:
            check(*c)
        except Exception as ex:  # keep going: one failing case should not hide the others
            print(f"case {c} raised {type(ex).__name__}: {ex}", flush=True)
            if "CUDA" in str(ex) or "cuda" in str(ex):
                break
    if "--time" in sys.argv or "--time-only" in sys.argv:
        for kind in (0, 1):
            try:
                timeit(kind)
            except Exception as ex:
                print(f"timing kind {kind} raised {ex}", flush=True)
                break
