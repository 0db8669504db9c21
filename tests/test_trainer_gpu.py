"""GPU: the native training step (flat buffers, gradient accumulation, clip, fused AdamW) against the oracle's gradients
and torch.optim.AdamW on the CPU."""
import pytest
import torch

from tests.helpers import load_golden, spec_from_golden, weights_from_golden
from tests.test_model_gpu import build_model

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_native_step_matches_oracle_adamw():
    from gamer_b200 import engine as E
    from gamer_b200.trainer import NativeTrainer
    from oracle import oracle_model as om
    g = load_golden("train_qwen3multi.pt")
    m = build_model(g).train()
    lr, wd = 1e-3, 0.01
    tr = NativeTrainer(m, lr=lr, weight_decay=wd, max_grad_norm=1.0)
    # parameters became views of the flat buffer without changing value or key
    sd0 = weights_from_golden(g)
    for k, v in m.state_dict().items():
        assert torch.equal(v.cpu(), sd0[k]), k
    batch = {k: v.to(DEV) for k, v in g["batch"].items()}
    loss = tr.step(batch, micro_batch=4)            # 6 rows -> micro-batches of 4 + 2, num_items normalisation
    assert abs(loss.item() - g["loss"].item()) <= 5e-3 * max(1.0, abs(g["loss"].item()))
    # oracle: same batch, same normalisation, clip 1.0, torch AdamW; decay groups by the transformers 4.51 name rule
    # (bias / layernorm / rmsnorm exempt; q_norm, k_norm and model.norm are decayed) the reference pins
    spec = spec_from_golden(g, g["temperature"])
    W = weights_from_golden(g, requires_grad=True)
    om.forward(spec, W, **g["batch"])["loss"].backward()
    named = {k: v for k, v in W.items() if k != "lm_head.weight"}
    exempt = lambda k: any(t in k.lower() for t in ("bias", "layernorm", "rmsnorm"))
    decay = [v for k, v in named.items() if not exempt(k)]
    no_decay = [v for k, v in named.items() if exempt(k)]
    for v in named.values():
        if v.grad is None:
            v.grad = torch.zeros_like(v)
    # gradient parity of the accumulated flat buffer
    mine = E.unfuse_grads(tr.arch, tr.G)
    num = sum(((mine[k].cpu() - v.grad) ** 2).sum() for k, v in named.items()) ** 0.5
    den = sum((v.grad ** 2).sum() for v in named.values()) ** 0.5
    assert (num / den).item() <= 3e-2, (num / den).item()
    torch.nn.utils.clip_grad_norm_(list(named.values()), 1.0)
    opt = torch.optim.AdamW([{"params": decay, "weight_decay": wd}, {"params": no_decay, "weight_decay": 0.0}], lr=lr)
    opt.step()
    # first AdamW step moves every weight by ~lr*sign(g): compare where the oracle gradient is not noise-level
    agree, total = 0, 0
    for k, p in m.named_parameters():
        ref, gr = named[k].detach(), named[k].grad
        sel = gr.abs() > 1e-3 * gr.abs().max().clamp(min=1e-12)
        d = (p.detach().cpu() - ref)[sel].abs()
        agree += int((d <= 0.25 * lr).sum())
        total += int(sel.sum())
    assert agree / total >= 0.98, agree / total
    # the bf16 operand copy written by the AdamW kernel is the rounded master copy
    assert torch.equal(tr.flat_bf16.float(), tr.flat_p.to(torch.bfloat16).float())
    # the batched transpose left every dgrad copy equal to the transposed operand the AdamW kernel just wrote
    pairs = tr.pack._transposes()
    assert len(pairs) > 10
    for src, dst in pairs:
        assert torch.equal(dst, src.t())
    # a second step runs on the refreshed pack and lowers the loss on the same batch
    loss2 = tr.step(batch, micro_batch=6)
    assert loss2.item() < loss.item()


def test_cuda_graph_replay_matches_eager():
    """The captured micro-batch graph (trainer default) reproduces the eager launch sequence: same losses over several
    optimizer steps with dropout ON (the device-side offset word gives every replay the masks the eager path draws),
    and the graph really is replayed."""
    from gamer_b200.trainer import NativeTrainer
    g = load_golden("train_qwen3multi.pt")
    batch = {k: v.to(DEV) for k, v in g["batch"].items()}
    losses = {}
    for mode in (False, True):
        m = build_model(g).train()
        m.config.dropout_rate = 0.2
        m.config.attention_dropout = 0.2
        m.set_dropout_seed(1234)
        tr = NativeTrainer(m, lr=1e-3, max_grad_norm=1.0, use_cuda_graphs=mode)
        losses[mode] = [tr.step(batch, micro_batch=3).item() for _ in range(5)]   # 6 rows -> 2 micro-batches of 3
        if mode:
            assert len(tr.graphs) == 1 and tr.micro_batches == 10
    for a, b in zip(losses[False], losses[True]):
        assert abs(a - b) <= 2e-3 * max(1.0, abs(a)), (losses[False], losses[True])
    assert losses[True][-1] < losses[True][0]
    # without dropout the replayed loss of the first step equals the golden loss
    m = build_model(g).train()
    tr = NativeTrainer(m, lr=1e-3, use_cuda_graphs=True)
    l0 = tr.step(batch, micro_batch=2).item()                                      # 3 micro-batches: eager, capture, replay
    assert abs(l0 - g["loss"].item()) <= 5e-3 * max(1.0, abs(g["loss"].item())), (l0, g["loss"].item())
