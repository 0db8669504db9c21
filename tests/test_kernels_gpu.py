"""GPU: each CUDA kernel through the C-ABI against a plain fp32 torch computation of the same op (and against the
oracle's helpers for router / mask semantics).  Integer outputs are compared bit-exactly; bf16 paths with the stated
tolerances."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _k():
    from gamer_b200 import kernels
    return kernels


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def bf(x):
    return x.to(torch.bfloat16)


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("rows,N,K", [(1000, 768, 256), (300, 256, 384), (129, 1024, 320), (5000, 256, 512),
                                      (64, 1041, 256), (20000, 1024, 256)])
def test_gemm_tn_plain(rows, N, K):
    k = _k()
    torch.manual_seed(0)
    a = bf(torch.randn(rows, K, device=DEV))
    b = bf(torch.randn(N, K, device=DEV) * 0.1)
    ldc = (N + 63) // 64 * 64          # output rows must stay 16-byte aligned (N = 1041 -> 1088-wide buffer)
    out = torch.empty(rows, ldc, dtype=torch.bfloat16, device=DEV)
    k.gemm_tn(a, b, N, out=out)
    ref = a.float() @ b.float().t()
    assert rel_err(out[:, :N], ref) < 6e-3, rel_err(out[:, :N], ref)
    chk = k.ref_gemm_tn(a, b, N, K)
    assert rel_err(chk, ref) < 1e-5
    out32 = torch.empty(rows, ldc, dtype=torch.float32, device=DEV)
    k.gemm_tn(a, b, N, out=out32, alpha=0.5)
    assert rel_err(out32[:, :N], 0.5 * ref) < 1e-5, rel_err(out32[:, :N], 0.5 * ref)
    if ldc != N:
        with pytest.raises(Exception, match="16-byte aligned"):
            k.gemm_tn(a, b, N)


def test_gemm_tn_k_tail_and_stride():
    """dgrad of the lm_head: reduction over V=1041 (not a multiple of 64) read from a 1088-wide padded buffer."""
    k = _k()
    torch.manual_seed(1)
    rows, V, H, ld = 777, 1041, 256, 1088
    a_full = torch.zeros(rows, ld, device=DEV)
    a_full[:, :V] = torch.randn(rows, V, device=DEV)
    a_full[:, V:] = 7.0  # garbage in the pad columns must not leak: K extent is V
    a = bf(a_full)
    wt_full = torch.zeros(H, ld, device=DEV)
    wt_full[:, :V] = torch.randn(H, V, device=DEV) * 0.1
    wt_full[:, V:] = 3.0
    wt = bf(wt_full)
    out = k.gemm_tn(a[:, :V], wt[:, :V], H, K=V)
    ref = a[:, :V].float() @ wt[:, :V].float().t()
    assert rel_err(out, ref) < 6e-3, rel_err(out, ref)


def test_gemm_tn_grouped_scatter_residual():
    k = _k()
    torch.manual_seed(2)
    E, N, K = 6, 256, 512
    counts = [0, 300, 128, 77, 1000, 1]
    seg = [0]
    for c in counts:
        seg.append(seg[-1] + (c + 127) // 128 * 128)
    Mp = seg[-1]
    M = sum(counts)
    a = torch.zeros(Mp + 256, K, device=DEV)
    rows = torch.full((Mp + 256,), -1, dtype=torch.int32, device=DEV)
    perm_src = torch.randperm(M, device=DEV).to(torch.int32)
    off = 0
    for e, c in enumerate(counts):
        a[seg[e]:seg[e] + c] = torch.randn(c, K, device=DEV)
        rows[seg[e]:seg[e] + c] = perm_src[off:off + c]
        off += c
    a = bf(a)
    w = bf(torch.randn(E * N, K, device=DEV) * 0.05)
    resid = bf(torch.randn(M, N, device=DEV))
    seg_t = torch.tensor(seg, dtype=torch.int32, device=DEV)
    out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=DEV)
    k.gemm_tn(a, w, N, rows=a.shape[0], n_groups=E, seg_off=seg_t, out=out, resid=resid, row_map=rows)
    ref = torch.empty(M, N, device=DEV)
    for e, c in enumerate(counts):
        if c:
            r = rows[seg[e]:seg[e] + c].long()
            ref[r] = resid[r].float() + a[seg[e]:seg[e] + c].float() @ w[e * N:(e + 1) * N].float().t()
    assert not torch.isnan(out.float()).any()
    assert rel_err(out, ref) < 6e-3, rel_err(out, ref)
    # permuted-space output (no row map): padding rows are zero because the A rows are zero
    out2 = k.gemm_tn(a, w, N, rows=a.shape[0], n_groups=E, seg_off=seg_t)
    for e, c in enumerate(counts):
        r2 = a[seg[e]:seg[e + 1]].float() @ w[e * N:(e + 1) * N].float().t()
        assert rel_err(out2[seg[e]:seg[e + 1]], r2) < 6e-3 or c == 0


@pytest.mark.parametrize("rows,N_out,K_in", [(1000, 768, 256), (4100, 256, 384), (2048, 1024, 320), (333, 1041, 256),
                                             (70000, 256, 512)])
def test_gemm_wgrad_plain(rows, N_out, K_in):
    k = _k()
    torch.manual_seed(3)
    ld = (N_out + 63) // 64 * 64
    dy = bf(torch.randn(rows, ld, device=DEV) * 0.1)[:, :N_out]
    x = bf(torch.randn(rows, K_in, device=DEV))
    dw = torch.zeros(1, N_out, K_in, device=DEV)
    k.gemm_wgrad(dy, x, N_out, K_in, dw)
    ref = dy.float().t() @ x.float()
    assert rel_err(dw[0], ref) < 1e-4, rel_err(dw[0], ref)
    k.gemm_wgrad(dy, x, N_out, K_in, dw)   # accumulates
    assert rel_err(dw[0], 2 * ref) < 1e-4


def test_gemm_wgrad_grouped_and_strided():
    k = _k()
    torch.manual_seed(4)
    E, N_out, K_in = 6, 1024, 320
    counts = [5, 300, 128, 77, 3000, 0]
    seg = [0]
    for c in counts:
        seg.append(seg[-1] + (c + 127) // 128 * 128)
    Mp = seg[-1]
    dy = torch.zeros(Mp + 128, N_out, device=DEV)
    xbuf = torch.zeros(Mp + 128, K_in + 64, device=DEV)   # wider buffer: row stride != K_in
    for e, c in enumerate(counts):
        dy[seg[e]:seg[e] + c] = torch.randn(c, N_out, device=DEV) * 0.1
        xbuf[seg[e]:seg[e] + c, :K_in] = torch.randn(c, K_in, device=DEV)
    dy, xbuf = bf(dy), bf(xbuf)
    x = xbuf[:, :K_in]
    dw = torch.zeros(E, N_out, K_in, device=DEV)
    k.gemm_wgrad(dy, x, N_out, K_in, dw, rows=dy.shape[0], n_groups=E,
                 seg_off=torch.tensor(seg, dtype=torch.int32, device=DEV))
    for e, c in enumerate(counts):
        ref = dy[seg[e]:seg[e + 1]].float().t() @ x[seg[e]:seg[e + 1]].float()
        if c == 0:
            assert dw[e].abs().max().item() == 0.0
        else:
            assert rel_err(dw[e], ref) < 1e-4, (e, rel_err(dw[e], ref))


# ------------------------------------------------------------------------------------------------ K1
def _spec():
    from oracle import oracle_model as om
    return om.Spec()


def _beh_lut(spec, V):
    lut = torch.arange(V, dtype=torch.int32)
    for t, i in spec.behavior_maps.items():
        lut[t] = i + 1
    return lut.to(DEV)


@pytest.mark.parametrize("left_pad", [False, True])
def test_embed_route_and_perm(left_pad):
    from gamer_b200 import synthetic as syn
    from oracle import oracle_model as om
    k = _k()
    spec = _spec()
    cat = syn.make_catalogue(2000, 1)
    if left_pad:
        batch, _ = syn.make_eval_batch(cat, 9, max_his_len=20, seed=7, median_len=8)
    else:
        batch = syn.make_train_batch(cat, 9, max_his_len=20, seed=7, median_len=8)
    ids = batch["input_ids"]
    B, L = ids.shape
    torch.manual_seed(0)
    table = bf(torch.randn(spec.vocab_size, 256))
    x, pos, beh, act = k.embed_route(ids.to(DEV), table.to(DEV), _beh_lut(spec, spec.vocab_size), spec.n_behavior, 5,
                                     spec.pad, spec.eos)
    rp, rb, ra = om.route(spec, ids, torch.arange(L))
    assert torch.equal(pos.cpu().view(B, L).long(), rp)
    assert torch.equal(beh.cpu().view(B, L).long(), rb)
    assert torch.equal(act.cpu().view(B, L).long(), ra)
    assert torch.equal(x.cpu(), table[ids.view(-1)])
    # routing permutation: a bijection onto 128-aligned expert segments, in token order inside each segment
    perm, rows, seg = k.route_perm(pos, B, L, 6)
    perm, rows, seg = perm.cpu().long(), rows.cpu().long(), seg.cpu().long()
    counts = torch.bincount(rp.view(-1), minlength=6)
    assert seg[0] == 0 and all(int(seg[e + 1] - seg[e]) == (int(counts[e]) + 127) // 128 * 128 for e in range(6))
    for e in range(6):
        toks = torch.nonzero(rp.view(-1) == e).view(-1)
        assert torch.equal(rows[seg[e]:seg[e] + len(toks)], toks)
        assert (rows[seg[e] + len(toks):seg[e + 1]] == -1).all()
    assert torch.equal(rows[perm], torch.arange(B * L))


def test_embed_route_flags_out_of_vocabulary_ids():
    """nn.Embedding raises IndexError on an id outside the table; the kernel gathers the pad row and flags it."""
    k = _k()
    spec = _spec()
    V = spec.vocab_size
    table = bf(torch.randn(V, 256)).to(DEV)
    lut = _beh_lut(spec, V)
    good = torch.randint(14, V, (2, 10), device=DEV)
    k.check_token_ids(DEV)                                   # clears anything an earlier test left behind
    k.embed_route(good, table, lut, spec.n_behavior, 5, spec.pad, spec.eos)
    k.check_token_ids(DEV)                                   # nothing flagged
    for bad_id in (V, V + 77, -1):
        bad = good.clone()
        bad[1, 3] = bad_id
        x, *_ = k.embed_route(bad, table, lut, spec.n_behavior, 5, spec.pad, spec.eos)
        assert torch.equal(x[13], table[spec.pad])
        with pytest.raises(IndexError):
            k.check_token_ids(DEV)
        k.check_token_ids(DEV)                               # the word was reset by the failed check


def test_embed_decode_step_route():
    from oracle import oracle_model as om
    k = _k()
    spec = _spec()
    torch.manual_seed(3)
    B, T = 6, 51   # 10 items + the target behaviour token, then generated tokens at positions 51..53
    ctx = torch.randint(14, 526, (B, T + 3))
    ctx[:, 0:T:5] = torch.randint(526, 529, (B, 11))
    table = bf(torch.randn(spec.vocab_size, 256)).to(DEV)
    lut = _beh_lut(spec, spec.vocab_size)
    for step in range(3):
        pos0 = T + step
        new = ctx[:, pos0:pos0 + 1].contiguous()
        x, pos, beh, act = k.embed_route(new.to(DEV), table, lut, spec.n_behavior, 5, spec.pad, spec.eos,
                                         ctx=ctx[:, :pos0 + 1].contiguous().to(DEV), pos0=pos0)
        rp, rb, ra = om.route(spec, new, torch.tensor([pos0]), ctx[:, :pos0 + 1])
        assert torch.equal(pos.cpu().view(B, 1).long(), rp)
        assert torch.equal(beh.cpu().view(B, 1).long(), rb)
        assert torch.equal(act.cpu().view(B, 1).long(), ra)


def test_embed_bwd_scatter_add():
    k = _k()
    torch.manual_seed(5)
    V, H, M, pad = 1041, 256, 40000, 4
    ids = torch.randint(0, V, (M,), device=DEV)
    ids[::5] = torch.randint(526, 529, (len(ids[::5]),), device=DEV)   # heavy hitters
    ids[-3000:] = pad
    dx = bf(torch.randn(M, H, device=DEV))
    sort_buf = k.embed_sort(ids, V, pad)
    dt = k.embed_bwd(dx, V, sort_buf)
    ref = torch.zeros(V, H, device=DEV)
    keep = ids != pad
    ref.index_add_(0, ids[keep], dx[keep].float())
    assert dt[pad].abs().max().item() == 0.0
    assert torch.allclose(dt, ref, rtol=1e-4, atol=1e-3), (dt - ref).abs().max().item()


# ------------------------------------------------------------------------------------------------ norms
def test_rmsnorm_fwd_bwd():
    k = _k()
    torch.manual_seed(6)
    M, H, eps = 3001, 256, 1e-6
    x = bf(torch.randn(M, H, device=DEV))
    w = (1 + 0.1 * torch.randn(H, device=DEV))
    tab = bf(torch.randn(4, 64, device=DEV))
    idx = torch.randint(0, 4, (M,), dtype=torch.int32, device=DEV)
    perm = torch.randperm(M + 100, device=DEV)[:M].to(torch.int32)
    out, rstd = k.rmsnorm_fwd(x, w, eps, row_map=perm, cat_table=tab, cat_idx=idx, out_rows=M + 100)
    xf = x.float().requires_grad_(True)
    wf = w.clone().requires_grad_(True)
    tf = tab.float().requires_grad_(True)
    ref = torch.cat([wf * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)), tf[idx.long()]], dim=-1)
    assert rel_err(out[perm.long()], ref) < 4e-3
    dh_p = torch.zeros(M + 100, H + 64, device=DEV)
    dh_p[perm.long()] = torch.randn(M, H + 64, device=DEV)
    dh_p = bf(dh_p)
    dres = bf(torch.randn(M, H, device=DEV))
    ref.backward(dh_p[perm.long()].float())
    dw = torch.zeros(H, device=DEV)
    dcat = torch.zeros(4, 64, device=DEV)
    dx = k.rmsnorm_bwd(x, w, rstd, eps, dh_p, dw, row_map=perm, dres=dres, cat_idx=idx, cat_dim=64, cat_rows=4, dcat=dcat)
    assert rel_err(dx, xf.grad + dres.float()) < 5e-3
    assert rel_err(dw, wf.grad) < 1e-3
    assert rel_err(dcat, tf.grad) < 1e-3


@pytest.mark.parametrize("cross", [False, True])
def test_qk_norm_rope_fwd_bwd(cross):
    from oracle import oracle_model as om
    k = _k()
    torch.manual_seed(7)
    spec = _spec()
    B, L, nq, nkv, hd, eps = 3, 37, 6, 3, 64, 1e-6
    M = B * L
    ld = 1024 if cross else 768
    raw = bf(torch.randn(M, ld, device=DEV))
    qn = 1 + 0.1 * torch.randn(hd, device=DEV)
    kn = 1 + 0.1 * torch.randn(hd, device=DEV)
    pos_ids = torch.randint(0, 200, (M,), dtype=torch.int32, device=DEV)
    cos, sin = om.rope_cos_sin(spec, torch.arange(200).unsqueeze(0))
    cos_t, sin_t = cos[0, :, :32].contiguous().to(DEV), sin[0, :, :32].contiguous().to(DEV)
    act = torch.randint(0, 4, (M,), dtype=torch.int32, device=DEV)
    qe = bf(torch.randn(4, nq * hd, device=DEV)) if cross else None
    ke = bf(torch.randn(4, nkv * hd, device=DEV)) if cross else None
    ve = bf(torch.randn(4, nkv * hd, device=DEV)) if cross else None
    out = k.qk_norm_rope_fwd(raw, L, nq, nkv, hd, cos_t, sin_t, qn, kn, eps, pos_ids=pos_ids, q_emb=qe, k_emb=ke,
                             v_emb=ve, act_idx=act if cross else None)
    # fp32 torch reference
    rawf = raw.float().requires_grad_(True)
    qnf, knf = qn.clone().requires_grad_(True), kn.clone().requires_grad_(True)
    embs = [t.float().requires_grad_(True) if t is not None else None for t in (qe, ke, ve)]
    q = rawf[:, :384].view(M, nq, hd)
    kk = rawf[:, 384:576].view(M, nkv, hd)
    v = rawf[:, 576:768].view(M, nkv, hd)
    if cross:
        q = q + embs[0][act.long()].view(M, nq, hd)
        kk = kk + embs[1][act.long()].view(M, nkv, hd)
        v = v + embs[2][act.long()].view(M, nkv, hd)
    c = cos[0].to(DEV)[pos_ids.long()].unsqueeze(1)
    s = sin[0].to(DEV)[pos_ids.long()].unsqueeze(1)

    def rope(t):
        rot = torch.cat([-t[..., 32:], t[..., :32]], dim=-1)
        return t * c + rot * s

    qo = rope(om.rmsnorm(q, qnf, eps))
    ko = rope(om.rmsnorm(kk, knf, eps))
    ref = torch.cat([qo.reshape(M, -1), ko.reshape(M, -1), v.reshape(M, -1)], dim=-1)
    assert rel_err(out[:, :768], ref) < 4e-3, rel_err(out[:, :768], ref)
    dout = bf(torch.randn(M, 768, device=DEV))
    ref.backward(dout.float())
    draw = torch.zeros(M, ld, dtype=torch.bfloat16, device=DEV)
    dqn, dkn = torch.zeros(hd, device=DEV), torch.zeros(hd, device=DEV)
    dqe = torch.zeros(4, nq * hd, device=DEV) if cross else None
    dke = torch.zeros(4, nkv * hd, device=DEV) if cross else None
    dve = torch.zeros(4, nkv * hd, device=DEV) if cross else None
    k.qk_norm_rope_bwd(raw, dout, draw, L, nq, nkv, hd, cos_t, sin_t, qn, kn, eps, dqn, dkn, pos_ids=pos_ids, q_emb=qe,
                       k_emb=ke, v_emb=ve, act_idx=act if cross else None, emb_rows=4, d_q_emb=dqe, d_k_emb=dke,
                       d_v_emb=dve)
    assert rel_err(draw[:, :768], rawf.grad[:, :768]) < 5e-3
    assert rel_err(dqn, qnf.grad) < 2e-3 and rel_err(dkn, knf.grad) < 2e-3
    if cross:
        for mine, r in zip((dqe, dke, dve), embs):
            assert rel_err(mine, r.grad) < 2e-3


def test_qk_norm_rope_positions_are_clamped_to_the_table():
    """A position outside [0, n_pos) reads the nearest table row (never out of bounds): the result equals the one for the
    clamped positions."""
    from oracle import oracle_model as om
    k = _k()
    torch.manual_seed(3)
    spec = _spec()
    M, nq, nkv, hd, eps, n_pos = 64, 6, 3, 64, 1e-6, 50
    raw = bf(torch.randn(M, 768, device=DEV))
    qn = 1 + 0.1 * torch.randn(hd, device=DEV)
    kn = 1 + 0.1 * torch.randn(hd, device=DEV)
    cos, sin = om.rope_cos_sin(spec, torch.arange(n_pos).unsqueeze(0))
    cos_t, sin_t = cos[0, :, :32].contiguous().to(DEV), sin[0, :, :32].contiguous().to(DEV)
    wild = torch.randint(-40, 4000, (M,), dtype=torch.int32, device=DEV)
    a = k.qk_norm_rope_fwd(raw, 1, nq, nkv, hd, cos_t, sin_t, qn, kn, eps, pos_ids=wild)
    b = k.qk_norm_rope_fwd(raw, 1, nq, nkv, hd, cos_t, sin_t, qn, kn, eps, pos_ids=wild.clamp(0, n_pos - 1))
    assert torch.equal(a, b)


# ------------------------------------------------------------------------------------------------ attention
def _attn_inputs(B, L, seed, left_pad):
    g = torch.Generator().manual_seed(seed)
    am = torch.ones(B, L, dtype=torch.int64)
    for b in range(B):
        n = int(torch.randint(L // 3, L + 1, (1,), generator=g))
        n = max(5, n - n % 5)
        if left_pad:
            am[b, :L - n] = 0
        else:
            am[b, n:] = 0
    act = torch.randint(0, 3, (B, (L + 4) // 5), generator=g).repeat_interleave(5, dim=1)[:, :L]
    act = torch.where(am.bool(), act, torch.full_like(act, 100))
    sess = torch.cumsum((torch.rand(B, (L + 4) // 5, generator=g) < 0.3).long(), dim=1).repeat_interleave(5, dim=1)[:, :L]
    sess = torch.where(am.bool(), sess, torch.zeros_like(sess))
    return am, act, sess


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
@pytest.mark.parametrize("L,left_pad", [(65, False), (200, True), (37, False), (505, False), (256, True), (129, False)])
def test_attention_fwd_bwd(kind, L, left_pad):
    from oracle import oracle_model as om
    k = _k()
    torch.manual_seed(8 + kind)
    B, nq, nkv, hd = 3, 6, 3, 64
    M = B * L
    am, act, sess = _attn_inputs(B, L, 11 + kind, left_pad)
    qkv = bf(torch.randn(M, 768, device=DEV))
    scale = hd ** -0.5
    i32 = lambda t: t.to(torch.int32).to(DEV).contiguous()
    o, lse, vmean, _ = k.attn_fwd(qkv, B, L, nq, nkv, hd, kind, 5, i32(am), i32(act), i32(sess), scale)
    allow = om.allow_matrix(kind, am, act, sess, 5).to(DEV)
    qf = qkv.float().requires_grad_(True)
    q = qf[:, :384].view(B, L, nq, hd).transpose(1, 2)
    kk = qf[:, 384:576].view(B, L, nkv, hd).transpose(1, 2)
    v = qf[:, 576:].view(B, L, nkv, hd).transpose(1, 2)
    ref = om.masked_attention(q, kk, v, allow, scale).transpose(1, 2).reshape(M, nq * hd)
    n_uniform = int((~allow.any(-1)).sum())
    err = rel_err(o, ref)
    assert err < 8e-3, (err, n_uniform)
    uni = ~allow.any(-1)                                        # [B, L]
    assert torch.equal(torch.isinf(lse[:, 0, :]), uni), "uniform-row flags differ"
    d_o = bf(torch.randn(M, nq * hd, device=DEV))
    ref.backward(d_o.float())
    dqkv = torch.zeros(M, 768, dtype=torch.bfloat16, device=DEV)
    k.attn_bwd(qkv, o, d_o, lse, B, L, nq, nkv, hd, kind, 5, i32(am), i32(act), i32(sess), scale, dqkv)
    g = qf.grad
    assert rel_err(dqkv[:, :384], g[:, :384]) < 2e-2, ("dq", rel_err(dqkv[:, :384], g[:, :384]), n_uniform)
    assert rel_err(dqkv[:, 384:576], g[:, 384:576]) < 2e-2, ("dk", rel_err(dqkv[:, 384:576], g[:, 384:576]))
    assert rel_err(dqkv[:, 576:], g[:, 576:]) < 2e-2, ("dv", rel_err(dqkv[:, 576:], g[:, 576:]))


# ------------------------------------------------------------------------------------------------ elementwise + CE
def test_swiglu_gate_gather():
    k = _k()
    torch.manual_seed(9)
    R, I = 1500, 512
    gu = bf(torch.randn(R, 2 * I, device=DEV))
    act = k.swiglu_fwd(gu, I)
    guf = gu.float().requires_grad_(True)
    ref = torch.nn.functional.silu(guf[:, :I]) * guf[:, I:]
    assert rel_err(act, ref) < 4e-3
    dact = bf(torch.randn(R, I, device=DEV))
    ref.backward(dact.float())
    assert rel_err(k.swiglu_bwd(gu, dact, I), guf.grad) < 5e-3
    # gate residual: g lives at columns 768.. of a 1024-wide buffer
    W = 256
    buf = bf(torch.randn(R, 1024, device=DEV))
    x, y = bf(torch.randn(R, W, device=DEV)), bf(torch.randn(R, W, device=DEV))
    out = k.gate_residual_fwd(x, y, buf[:, 768:])
    yf, gf = y.float().requires_grad_(True), buf[:, 768:].float().requires_grad_(True)
    ref = x.float() + yf * torch.nn.functional.silu(gf)
    assert rel_err(out, ref) < 4e-3
    dout = bf(torch.randn(R, W, device=DEV))
    ref.backward(dout.float())
    dbuf = torch.zeros(R, 1024, dtype=torch.bfloat16, device=DEV)
    dy = k.gate_residual_bwd(dout, y, buf[:, 768:], dbuf[:, 768:])
    assert rel_err(dy, yf.grad) < 5e-3 and rel_err(dbuf[:, 768:], gf.grad) < 5e-3
    rows = torch.randint(-1, R, (2000,), dtype=torch.int32, device=DEV)
    gth = k.gather_rows(x, rows, 2000)
    ref = torch.where((rows >= 0)[:, None], x[rows.clamp(min=0).long()].float(), torch.zeros(1, device=DEV))
    assert torch.equal(gth.float(), ref)


def test_cross_entropy_fused():
    k = _k()
    torch.manual_seed(10)
    R, V, ld = 4000, 1041, 1088
    logits = torch.randn(R, V, device=DEV) * 3
    labels = torch.randint(0, V, (R,), device=DEV)
    labels[::7] = -100
    n_valid = (labels != -100).sum()
    inv = (1.0 / n_valid.float()).reshape(1)
    dl = torch.full((R, ld), 5.0, dtype=torch.bfloat16, device=DEV)
    loss_row = k.ce_fwd_bwd(logits, labels, V, inv, 1.0 / 0.7, dlogits=dl)
    lf = logits.clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lf, labels, ignore_index=-100, reduction="mean")
    assert abs((loss_row.sum() * inv).item() - ref.item()) < 1e-4
    ref.backward()
    assert rel_err(dl[:, :V], lf.grad / 0.7) < 5e-3
    assert dl[:, V:].abs().max().item() == 0.0


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_attention_long_history(kind):
    """BASELINE.json configs[4]: max_his_len = 500 (L = 2505, 20 key tiles) through the tcgen05 kernels, ragged rows."""
    from oracle import oracle_model as om
    k = _k()
    torch.manual_seed(40 + kind)
    B, L, nq, nkv, hd = 2, 2505, 6, 3, 64
    M = B * L
    am, act, sess = _attn_inputs(B, L, 31 + kind, kind == 2)
    qkv = bf(torch.randn(M, 768, device=DEV))
    scale = hd ** -0.5
    i32 = lambda t: t.to(torch.int32).to(DEV).contiguous()
    o, lse, _, _ = k.attn_fwd(qkv, B, L, nq, nkv, hd, kind, 5, i32(am), i32(act), i32(sess), scale)
    allow = om.allow_matrix(kind, am, act, sess, 5).to(DEV)
    qf = qkv.float().requires_grad_(True)
    q = qf[:, :384].view(B, L, nq, hd).transpose(1, 2)
    kk = qf[:, 384:576].view(B, L, nkv, hd).transpose(1, 2)
    v = qf[:, 576:].view(B, L, nkv, hd).transpose(1, 2)
    ref = om.masked_attention(q, kk, v, allow, scale).transpose(1, 2).reshape(M, nq * hd)
    assert rel_err(o, ref) < 8e-3
    d_o = bf(torch.randn(M, nq * hd, device=DEV))
    ref.backward(d_o.float())
    dqkv = torch.zeros(M, 768, dtype=torch.bfloat16, device=DEV)
    k.attn_bwd(qkv, o, d_o, lse, B, L, nq, nkv, hd, kind, 5, i32(am), i32(act), i32(sess), scale, dqkv)
    g = qf.grad
    for name, sl in (("dq", slice(0, 384)), ("dk", slice(384, 576)), ("dv", slice(576, 768))):
        e = rel_err(dqkv[:, sl], g[:, sl])
        assert e < 2e-2, (name, e)


def test_zero_unmapped_rows():
    """Rows whose map entry is negative are cleared, the others are left alone (ragged row count, strided rows)."""
    k = _k()
    torch.manual_seed(5)
    n, W = 1000 + 37, 320
    buf = bf(torch.randn(n, W + 64, device=DEV))[:, :W]
    before = buf.clone()
    rows = torch.randint(0, 50, (n,), dtype=torch.int32, device=DEV)
    rows[torch.rand(n, device=DEV) < 0.2] = -1
    rows[-1] = -1
    k.zero_unmapped_rows(buf, rows)
    neg = rows < 0
    assert (buf[neg] == 0).all() and torch.equal(buf[~neg], before[~neg])
