"""In-kernel timeline of the tcgen05 attention forward (CTA 0): where do the MMA thread and the softmax warps wait?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gamer_b200 import kernels as k          # noqa: E402
from gamer_b200._cabi import call            # noqa: E402
from tests.test_kernels_gpu import _attn_inputs, bf   # noqa: E402

DEV = "cuda:0"
kind = int(sys.argv[1]) if len(sys.argv) > 1 else 0
B, L, nq, nkv, hd = 128, 505, 6, 3, 64
am, act, sess = _attn_inputs(B, L, 3, False)
am[:] = 1
act = torch.where(act == 100, torch.zeros_like(act), act)
qkv = bf(torch.randn(B * L, 768, device=DEV))
i32 = lambda t: t.to(torch.int32).to(DEV).contiguous()
a, c, s = i32(am), i32(act), i32(sess)
for _ in range(2):
    k.attn_fwd(qkv, B, L, nq, nkv, hd, kind, 5, a, c, s, hd ** -0.5)
CAP = 400
buf = torch.zeros(3, CAP, 2, dtype=torch.int64, device=DEV)
call("gamer_attn_set_trace", buf.data_ptr(), CAP)
BWD = "--bwd" in sys.argv
if BWD:
    call("gamer_attn_set_trace", 0, 0)
    o, lse, _ = k.attn_fwd(qkv, B, L, nq, nkv, hd, kind, 5, a, c, s, hd ** -0.5)
    d_o = bf(torch.randn(B * L, nq * hd, device=DEV))
    dqkv = torch.zeros(B * L, 768, dtype=torch.bfloat16, device=DEV)
    k.attn_bwd(qkv, o, d_o, lse, B, L, nq, nkv, hd, kind, 5, a, c, s, hd ** -0.5, dqkv)
    call("gamer_attn_set_trace", buf.data_ptr(), CAP)
    k.attn_bwd(qkv, o, d_o, lse, B, L, nq, nkv, hd, kind, 5, a, c, s, hd ** -0.5, dqkv)
else:
    k.attn_fwd(qkv, B, L, nq, nkv, hd, kind, 5, a, c, s, hd ** -0.5)
torch.cuda.synchronize()
call("gamer_attn_set_trace", 0, 0)
t = buf.cpu()
t0 = int(t[:, 0, 1][t[:, 0, 1] > 0].min())
if BWD:
    names = {1: "kv_full", 2: "qdo_full", 3: "S(n+1) issued", 4: "pds_full", 5: "dq_free", 20: "step start", 21: "s_full",
             22: "S loaded", 23: "P computed", 24: "p_free", 25: "P stored", 26: "dp_full", 27: "ds_free", 28: "dS stored",
             29: "arrived", 40: "drain start", 41: "dq_full", 42: "stg free h0", 43: "stg free h1"}
    role_names = ("MMA", "softmax w4", "drain w12")
else:
    role_names = ("MMA", "softmax A", "softmax B")
names_f = {1: "q_full", 2: "kv_full", 10: "p_full[A]", 11: "p_full[B]", 20: "tile start", 21: "s_full", 22: "tmem ld done",
         23: "max done", 24: "o_full/rescale done", 25: "exp done", 26: "P stored+arrive", 30: "epi o_full", 31: "epi done"}
if not BWD:
    names = names_f
for role, nm in enumerate(role_names):
    print(f"--- {nm}")
    prev = None
    for n in range(CAP):
        tag, clk = int(t[role, n, 0]), int(t[role, n, 1])
        if clk == 0:
            break
        print(f"  {clk - t0:8d} (+{0 if prev is None else clk - prev:6d}) {names.get(tag, tag)}")
        prev = clk
        if n > 70:
            break
