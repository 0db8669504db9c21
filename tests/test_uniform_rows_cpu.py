"""CPU: the closed form DESIGN.md §8.2 gives for the gradient of "uniform rows" (quirk Q1: a query with no allowed key
attends uniformly to all L keys) against autograd of the reference's eager attention arithmetic
(SeqRec/models/generative/Qwen3Multi/model.py:123-143 with the additive finfo.min masks of :573-741).
With U the uniform rows, Vm the column mean of V, s = sum_{i in U} dO_i, C = sum_j (V_j - Vm)^T K_j and
G = sum_{i in U} dO_i^T Q_i:   dV_j = s / L,   dQ_i = (scale / L) dO_i C  (i in U),   dK_j = (scale / L) (V_j - Vm) G."""
import torch


def _eager(q, k, v, mask_add, scale):
    s = (q @ k.t()) * scale + mask_add
    p = torch.softmax(s, dim=-1)          # (the reference computes it in fp32: the masked scores absorb s either way)
    return p @ v


def test_uniform_row_gradient_closed_form():
    torch.manual_seed(0)
    L, D, scale = 37, 16, 16 ** -0.5
    q = torch.randn(L, D, dtype=torch.float64, requires_grad=True)
    k = torch.randn(L, D, dtype=torch.float64, requires_grad=True)
    v = torch.randn(L, D, dtype=torch.float64, requires_grad=True)
    d_o = torch.randn(L, D, dtype=torch.float64)
    uni = torch.rand(L) < 0.6                                  # rows with no allowed key
    uni[0] = True
    causal = torch.tril(torch.ones(L, L, dtype=torch.bool))
    allowed = causal & ~uni.unsqueeze(1)
    neg = torch.finfo(torch.float32).min
    mask_add = torch.where(allowed, 0.0, neg).to(torch.float64)
    out = _eager(q, k, v, mask_add, scale)
    # the forward of a uniform row is the column mean of V over ALL keys
    assert torch.allclose(out[uni], v.mean(0, keepdim=True).expand(int(uni.sum()), D).double(), atol=1e-10)
    # gradient of the uniform rows alone
    (out * d_o * uni.unsqueeze(1)).sum().backward()
    vm = v.detach().mean(0, keepdim=True)
    do_u = d_o * uni.unsqueeze(1)
    s = do_u.sum(0, keepdim=True)
    C = (v.detach() - vm).t() @ k.detach()
    G = do_u.t() @ q.detach()
    dv = s.expand(L, D) / L
    dq = (scale / L) * do_u @ C
    dk = (scale / L) * (v.detach() - vm) @ G
    # float64 throughout: agreement to rounding
    assert torch.allclose(v.grad, dv, atol=1e-10)
    assert torch.allclose(q.grad, dq, atol=1e-10)
    assert torch.allclose(k.grad, dk, atol=1e-10)
