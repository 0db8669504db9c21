"""CPU: the oracle restatement against golden vectors frozen from the reference (oracle/make_golden.py)."""
import pytest
import torch

from gamer_b200 import synthetic as syn
from oracle import oracle_decode as od
from oracle import oracle_model as om
from tests.helpers import load_golden, spec_from_golden, weights_from_golden

TRAIN = ["train_qwen3multi.pt", "train_qwen3multi_numitems.pt", "train_qwen3sessionmoe.pt",
         "train_qwen3sessionmulti.pt", "train_qwen3moe.pt"]


@pytest.mark.parametrize("name", TRAIN)
def test_forward_loss_grads(name):
    g = load_golden(name)
    spec = spec_from_golden(g, g["temperature"])
    W = weights_from_golden(g, requires_grad=True)
    out = om.forward(spec, W, **g["batch"], num_items_in_batch=g["num_items_in_batch"])
    # router indices: bit-exact
    for a, b in zip(g["route"], out["route"]):
        assert torch.equal(a, b)
    # fp32 vs fp32, different summation order only
    assert torch.allclose(out["logits"], g["logits"], rtol=1e-4, atol=2e-5)
    assert abs(out["loss"].item() - g["loss"].item()) < 1e-5 * max(1.0, abs(g["loss"].item()))
    assert abs(g["loss"].item() - g["loss_eager"].item()) < 1e-5 * max(1.0, abs(g["loss"].item()))
    out["loss"].backward()
    for k, d in g["grads"].items():
        gr = W[k].grad if W[k].grad is not None else torch.zeros_like(W[k])
        ref_norm = d["norm"].item()
        assert abs(gr.norm().item() - ref_norm) <= 1e-4 * ref_norm + 1e-7, k
        mine = gr.reshape(-1)[::d["stride"]][: d["samples"].numel()]
        assert torch.allclose(mine, d["samples"], rtol=1e-3, atol=1e-5 * max(ref_norm, 1e-3)), k
    assert torch.allclose(W["model.embed_tokens.weight"].grad, g["embed_grad"], rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("name,tol", [("decode_qwen3multi_lvl2.pt", 1e-5), ("decode_qwen3multi_lvl1.pt", 1e-5)])
def test_constrained_beam_search(name, tol):
    g = load_golden(name)
    spec = spec_from_golden(g)
    W = weights_from_golden(g)
    cat = syn.make_catalogue(g["catalogue_size"], g["catalogue_seed"])
    items = cat.item_sequences(g["target_behavior"]).tolist()
    tree = od.PrefixTree(items)
    last = set(t[-1] for t in items) | {spec.pad}
    b = g["batch"]
    with torch.no_grad():
        seqs, scores = od.constrained_beam_search(spec, W, tree, last, b["input_ids"], b["attention_mask"],
                                                  b["session_ids"], b["extended_session_ids"], b["actions"],
                                                  num_beams=g["num_beams"])
    assert torch.equal(seqs, g["sequences"])
    assert torch.allclose(scores, g["sequences_scores"], atol=tol, rtol=0)
    # trie children: bit-exact against the reference Trie
    for prefix, allowed in g["trie_probes"]:
        assert sorted(tree.children(prefix)) == allowed


def test_ranking_metrics():
    g = load_golden("decode_qwen3multi_lvl2.pt")
    K = g["num_beams"]
    hits = od.hit_lists(g["rank_pred"], g["sequences_scores"], g["rank_targets"], K)
    assert hits == g["rank_hits"]
    met = od.metrics(hits, g["rank_targets"], g["rank_names"])
    for k, v in g["rank_metrics"].items():
        assert abs(met[k] - v) < 1e-12, k


FULL_SIZE = ["train_qwen3multi_headline.pt", "train_qwen3moe_mb4.pt"]


@pytest.mark.parametrize("name", FULL_SIZE)
def test_full_size_forward_loss_grads(name):
    """BASELINE.json's shapes (all 8 layers; L = 505 Qwen3Multi / L = 1005 four-behaviour Qwen3Moe): the oracle against
    the unmodified reference's loss, strided logits, gradient digests and full embedding gradient."""
    g = load_golden(name)
    spec = spec_from_golden(g, g["temperature"])
    W = weights_from_golden(g, requires_grad=True)
    out = om.forward(spec, W, **g["batch"])
    mine = out["logits"].reshape(-1)[::g["logits_stride"]]
    assert torch.allclose(mine, g["logits_samples"], rtol=1e-4, atol=5e-5)
    assert abs(out["loss"].item() - g["loss"].item()) < 1e-5 * max(1.0, abs(g["loss"].item()))
    out["loss"].backward()
    for k, d in g["grads"].items():
        gr = W[k].grad if W[k].grad is not None else torch.zeros_like(W[k])
        ref_norm = d["norm"].item()
        assert abs(gr.norm().item() - ref_norm) <= 1e-4 * ref_norm + 1e-7, k
        mine = gr.reshape(-1)[::d["stride"]][: d["samples"].numel()]
        assert torch.allclose(mine, d["samples"], rtol=1e-3, atol=1e-5 * max(ref_norm, 1e-3)), k
    assert torch.allclose(W["model.embed_tokens.weight"].grad, g["embed_grad"], rtol=1e-3, atol=1e-6)
