"""GPU: training-mode dropout of the CUDA path.  The masks are Philox bits regenerated in the backward; the numpy
restatement (oracle/dropout_masks.py) reproduces them bit-exactly, so every kernel — and the whole model, loss and
gradients — is compared with the oracle fed the SAME masks (tolerances as in the dropout-free tests)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"
SEED, OFFSET = 0x1234_5678_9ABC_DEF0, 5


def _k():
    from gamer_b200 import kernels
    return kernels


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def bf(x):
    return x.to(torch.bfloat16)


def test_hidden_masks_bit_exact():
    """dropout_apply / gather_rows / GEMM epilogues / SwiGLU / gate-residual all realise the oracle's mask exactly."""
    from oracle import dropout_masks as dm
    k = _k()
    torch.manual_seed(0)
    R, W = 777, 256
    site = 3 * 8 + dm.SITE_SELF_OUT
    d = k.Dropout(SEED, OFFSET, site, 0.2)
    keep, scale = dm.hidden_keep(SEED, OFFSET, site, R, W, 0.2)
    keep = keep.to(DEV)
    ones = torch.ones(R, W, dtype=torch.bfloat16, device=DEV)
    out = k.dropout_apply(ones, d)
    assert torch.equal(out != 0, keep), "mask bits differ from the numpy restatement"
    assert torch.allclose(out.float()[keep], torch.tensor(scale, device=DEV).to(torch.bfloat16).float().expand(int(keep.sum())))
    assert abs(keep.float().mean().item() - 0.8) < 5e-3
    # staged GEMM epilogue: out = x + dropout(A B^T)
    A, B_ = bf(torch.randn(R, 384, device=DEV)), bf(torch.randn(W, 384, device=DEV) * 0.1)
    x = bf(torch.randn(R, W, device=DEV))
    got = k.gemm_tn(A, B_, W, resid=x, drop=d)
    ref = x.float() + (A.float() @ B_.float().t()) * keep * scale
    assert rel_err(got, ref) < 6e-3
    # scattered (row_map) epilogue: the mask is indexed by the OUTPUT row
    perm = torch.randperm(R, device=DEV).to(torch.int32)
    got2 = torch.zeros(R, W, dtype=torch.bfloat16, device=DEV)
    k.gemm_tn(A, B_, W, resid=x, row_map=perm, out=got2, drop=d)
    ref2 = torch.zeros(R, W, device=DEV)
    pl = perm.long()
    ref2[pl] = x.float()[pl] + (A.float() @ B_.float().t()) * keep[pl] * scale
    assert rel_err(got2, ref2) < 6e-3
    # gather_rows: backward of the scattered epilogue
    rows = torch.randint(-1, R, (1000,), dtype=torch.int32, device=DEV)
    g = k.gather_rows(x, rows, 1000, drop=d)
    rl = rows.clamp(min=0).long()
    refg = torch.where((rows >= 0)[:, None], x.float()[rl] * keep[rl] * scale, torch.zeros(1, device=DEV))
    assert rel_err(g, refg) < 4e-3 and torch.equal(g != 0, (refg != 0) & (x[rl] != 0))
    # SwiGLU with row ids (mask indexed by token row) fwd + bwd
    I = 512
    site_i = 2 * 8 + dm.SITE_FFN_INNER
    di = k.Dropout(SEED, OFFSET, site_i, 0.2)
    keep_i, scale_i = dm.hidden_keep(SEED, OFFSET, site_i, R, I, 0.2)
    keep_i = keep_i.to(DEV)
    gu = bf(torch.randn(1000, 2 * I, device=DEV))
    act = k.swiglu_fwd(gu, I, row_ids=rows, drop=di)
    guf = gu.float().requires_grad_(True)
    z = torch.where((rows >= 0)[:, None], keep_i[rl].float() * scale_i, torch.ones(1, device=DEV))
    ref = torch.nn.functional.silu(guf[:, :I]) * guf[:, I:] * z
    assert rel_err(act, ref) < 4e-3
    dact = bf(torch.randn(1000, I, device=DEV))
    ref.backward(dact.float())
    assert rel_err(k.swiglu_bwd(gu, dact, I, row_ids=rows, drop=di), guf.grad) < 5e-3
    # gate residual fwd + bwd
    buf = bf(torch.randn(R, 1024, device=DEV))
    y = bf(torch.randn(R, W, device=DEV))
    out = k.gate_residual_fwd(x, y, buf[:, 768:], drop=d)
    yf, gf = y.float().requires_grad_(True), buf[:, 768:].float().requires_grad_(True)
    ref = x.float() + yf * torch.nn.functional.silu(gf) * keep * scale
    assert rel_err(out, ref) < 4e-3
    dout = bf(torch.randn(R, W, device=DEV))
    ref.backward(dout.float())
    dbuf = torch.zeros(R, 1024, dtype=torch.bfloat16, device=DEV)
    dy = k.gate_residual_bwd(dout, y, buf[:, 768:], dbuf[:, 768:], drop=d)
    assert rel_err(dy, yf.grad) < 5e-3 and rel_err(dbuf[:, 768:], gf.grad) < 5e-3


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
@pytest.mark.parametrize("L,left_pad", [(505, False), (200, True), (129, False)])
def test_attention_dropout_fwd_bwd(kind, L, left_pad):
    """attention with dropout on the probabilities vs the oracle given the same Philox keep mask."""
    from oracle import dropout_masks as dm
    from oracle import oracle_model as om
    from tests.test_kernels_gpu import _attn_inputs
    k = _k()
    torch.manual_seed(18 + kind)
    B, nq, nkv, hd = 3, 6, 3, 64
    M = B * L
    am, act, sess = _attn_inputs(B, L, 21 + kind, left_pad)
    qkv = bf(torch.randn(M, 768, device=DEV))
    scale = hd ** -0.5
    i32 = lambda t: t.to(torch.int32).to(DEV).contiguous()
    site = 5 * 8 + (dm.SITE_CROSS_P if kind in (1, 3) else dm.SITE_SELF_P)
    d = k.Dropout(SEED, OFFSET, site, 0.2)
    keep, zs = dm.attn_keep(SEED, OFFSET, site, B, nq, L, 0.2)
    zp = keep.float().to(DEV) * zs
    o, lse, _, kw = k.attn_fwd(qkv, B, L, nq, nkv, hd, kind, 5, i32(am), i32(act), i32(sess), scale, drop=d)
    o0, lse0, _, _ = k.attn_fwd(qkv, B, L, nq, nkv, hd, kind, 5, i32(am), i32(act), i32(sess), scale)
    assert torch.equal(lse, lse0), "lse must not depend on dropout (softmax is normalised before dropping)"
    allow = om.allow_matrix(kind, am, act, sess, 5).to(DEV)
    # the keep words the forward stored (layout [b*n_q + h][query tile][32-key block][128 rows]) equal the oracle's wherever
    # a pair is allowed (blocks without an allowed pair are never generated)
    words, _ = dm.attn_keep_words(SEED, OFFSET, site, B, nq, L, 0.2)
    qt_n, nw_p = (L + 127) // 128, 4 * ((L + 127) // 128)
    got = kw.view(B * nq, qt_n, nw_p, 128).permute(0, 1, 3, 2).reshape(B * nq, qt_n * 128, nw_p)[:, :L, :words.shape[-1]]
    want = torch.from_numpy(words.astype("int64")).to(DEV)
    blk_any = torch.nn.functional.pad(allow, (0, words.shape[-1] * 32 - L)).view(B, L, -1, 32).any(-1)   # [B, L, nw]
    blk_any = blk_any[:, None].expand(B, nq, L, -1).reshape(B * nq, L, -1)
    assert torch.equal((got.long() & 0xffffffff)[blk_any], want[blk_any]), "stored keep words differ from the oracle generator"
    qf = qkv.float().requires_grad_(True)
    q = qf[:, :384].view(B, L, nq, hd).transpose(1, 2)
    kk = qf[:, 384:576].view(B, L, nkv, hd).transpose(1, 2)
    v = qf[:, 576:].view(B, L, nkv, hd).transpose(1, 2)
    ref = om.masked_attention(q, kk, v, allow, scale, zp).transpose(1, 2).reshape(M, nq * hd)
    err = rel_err(o, ref)
    assert err < 1e-2, err
    assert rel_err(o0, ref) > 0.1, "dropout had no effect"
    d_o = bf(torch.randn(M, nq * hd, device=DEV))
    ref.backward(d_o.float())
    dqkv = torch.zeros(M, 768, dtype=torch.bfloat16, device=DEV)
    k.attn_bwd(qkv, o, d_o, lse, B, L, nq, nkv, hd, kind, 5, i32(am), i32(act), i32(sess), scale, dqkv, drop=d, keep=kw)
    g = qf.grad
    for name, sl in (("dq", slice(0, 384)), ("dk", slice(384, 576)), ("dv", slice(576, 768))):
        e = rel_err(dqkv[:, sl], g[:, sl])
        assert e < 2e-2, (name, e)


@pytest.mark.parametrize("name", ["train_qwen3multi.pt", "train_qwen3sessionmoe.pt", "train_qwen3sessionmulti.pt"])
def test_model_train_mode_dropout_vs_oracle(name):
    """Whole model in train mode with dropout_rate = attention_dropout = 0.2: loss and gradients vs the oracle run
    with the same masks; a second forward draws fresh masks (the offset advances); eval mode is dropout-free."""
    from oracle import dropout_masks as dm
    from oracle import oracle_model as om
    from tests.helpers import load_golden, spec_from_golden, weights_from_golden
    from tests.test_model_gpu import build_model
    g = load_golden(name)
    m = build_model(g).train()
    m.config.dropout_rate = 0.2
    m.config.attention_dropout = 0.2
    m.set_dropout_seed(SEED)
    batch = {k_: v.to(DEV) for k_, v in g["batch"].items()}
    out = m(**batch)
    out.loss.backward()
    spec = spec_from_golden(g, g["temperature"])
    W = weights_from_golden(g, requires_grad=True)
    ref = om.forward(spec, W, **g["batch"], drop=dm.OracleDropout(SEED, 0, 0.2, 0.2))
    ref["loss"].backward()
    lo, lr = out.loss.item(), ref["loss"].item()
    assert abs(lo - lr) <= 5e-3 * max(1.0, abs(lr)), (lo, lr)
    assert abs(lo - g["loss"].item()) > 1e-3, "train-mode loss equals the dropout-free golden: dropout not applied"
    params = dict(m.named_parameters())
    worst = (0.0, None)
    for key, wref in W.items():
        if key == "lm_head.weight" or wref.grad is None:
            continue
        gr, rr = params[key].grad.float().cpu(), wref.grad
        if rr.norm() < 1e-6:
            continue
        rn = abs(gr.norm().item() - rr.norm().item()) / rr.norm().item()
        cos = torch.nn.functional.cosine_similarity(gr.reshape(-1), rr.reshape(-1), dim=0).item()
        worst = max(worst, (rn, key))
        assert rn <= 3e-2 and cos >= 0.998, (key, rn, cos)
    print(f"{name}: dropout loss {lo:.5f} vs oracle {lr:.5f}; worst grad-norm rel diff {worst[0]:.3e} ({worst[1]})")
    out2 = m(**batch)                                              # offset 1: different masks
    ref2 = om.forward(spec, W, **g["batch"], drop=dm.OracleDropout(SEED, 1, 0.2, 0.2))
    assert abs(out2.loss.item() - ref2["loss"].item()) <= 5e-3 * max(1.0, abs(ref2["loss"].item()))
    assert out2.loss.item() != lo
    m.eval()
    with torch.no_grad():
        oute = m(**batch)
    assert abs(oute.loss.item() - g["loss"].item()) <= 5e-3 * max(1.0, abs(g["loss"].item()))


def test_trainer_steps_with_dropout():
    """NativeTrainer with the packaged config's dropout (0.2 / 0.2): loss finite and decreasing over a few steps."""
    import os
    from transformers.models.qwen3_moe import Qwen3MoeConfig
    from gamer_b200 import modeling
    from gamer_b200 import synthetic as syn
    from gamer_b200.trainer import NativeTrainer
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = Qwen3MoeConfig.from_pretrained(os.path.join(root, "config", "s2s-models", "Qwen3Multi"))
    cfg.vocab_size = 1041; cfg.num_behavior = 3; cfg.behavior_maps = {"526": 0, "527": 1, "528": 2}
    cfg.use_behavior_token = True; cfg.num_positions = 5; cfg.num_experts = 6; cfg.n_positions = 21
    cfg.use_user_token = False; cfg.model_max_length = 1024; cfg.num_hidden_layers = 2
    cfg.sparse_layers_decoder = [0, 1]; cfg.behavior_injection_decoder = [0]; cfg.cross_attention_decoder = [1]
    assert cfg.dropout_rate == 0.2 and cfg.attention_dropout == 0.2
    torch.manual_seed(0)
    m = modeling.Qwen3MultiWithTemperature(cfg).to(DEV).train()
    m.set_hyper(0.7)
    tr = NativeTrainer(m, lr=2e-3, warmup_steps=0)
    cat = syn.make_catalogue(2000, 1)
    batch = {k_: v.to(DEV) for k_, v in syn.make_train_batch(cat, 16, max_his_len=20, seed=3).items()}
    losses = [tr.step(batch, micro_batch=8).item() for _ in range(8)]
    assert all(l == l and abs(l) < 1e4 for l in losses), losses
    assert losses[-1] < losses[0], losses
    assert tr.micro_batches == 16
