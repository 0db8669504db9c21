"""Freeze golden vectors from the UNMODIFIED reference (run in the build container only; /root/reference required).

    python -m oracle.make_golden          # writes tests/golden/*.pt

TEST INFRASTRUCTURE ONLY.  The reference has no tests of its own, so these fixtures — outputs of the
reference's classes executed here through oracle/ref_shim.py — are the pins the oracle (and through it the CUDA
path) is checked against.  Weights are not stored: they are regenerated from a seed with
`gamer_b200.synthetic.seeded_state_dict` (a checksum is stored to detect RNG drift).

Gradient goldens use `_attn_implementation="eager"` (the reference's own `eager_attention_forward` branch,
Qwen3Multi/model.py:123): torch's fused SDPA *backward* recomputes P = exp(s - lse) with lse saturated at finfo.min on
fully-masked rows (quirk Q1), which inflates those rows' gradients by the key count — an artefact of the fused kernel,
not the derivative of the forward (finite differences agree with eager).  Forward goldens are taken with the default
"sdpa"; the two agree to 1e-6 in the forward.  See DESIGN.md "Q15".
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from gamer_b200 import synthetic as syn  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
GRAD_SAMPLES = 64


def tiny_config(variant: str, n_layers: int = 3):
    cfg = ref_shim.reference_config(variant, max_his_len=20, num_layers=n_layers)
    cfg.behavior_injection_decoder = [0]
    if variant not in ("Qwen3SessionMoe", "Qwen3Moe"):
        cfg.cross_attention_decoder = [1, 2]
    return cfg


def build(R, variant, cfg, seed, attn_impl="sdpa", temperature=0.7):
    cfg._attn_implementation = attn_impl
    m = R[variant](cfg)
    m.set_hyper(temperature)
    m.eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = syn.seeded_state_dict(shapes, seed=seed)
    sd["lm_head.weight"] = sd["model.embed_tokens.weight"]
    m.load_state_dict(sd)
    return m, sd, shapes


def checksum(sd):
    return float(sum(v.double().abs().sum() for k, v in sorted(sd.items()) if k != "lm_head.weight"))


def grad_digest(named_grads):
    """Per-parameter: L2 norm + GRAD_SAMPLES strided samples (full grads would be tens of MB)."""
    out = {}
    for k, g in named_grads:
        flat = g.reshape(-1)
        stride = max(1, flat.numel() // GRAD_SAMPLES)
        out[k] = {"norm": g.norm().double(), "samples": flat[::stride][:GRAD_SAMPLES].clone(), "stride": stride}
    return out


def cfg_public(cfg):
    keys = ["vocab_size", "hidden_size", "num_attention_heads", "num_key_value_heads", "head_dim", "intermediate_size",
            "num_hidden_layers", "behavior_embedding_dim", "num_behavior", "num_positions", "num_experts",
            "sparse_layers_decoder", "behavior_injection_decoder", "behavior_maps", "pad_token_id", "eos_token_id",
            "rms_norm_eps", "n_positions", "model_max_length"]
    d = {k: getattr(cfg, k) for k in keys}
    d["cross_attention_decoder"] = list(getattr(cfg, "cross_attention_decoder", []) or [])
    d["rope_theta"] = 1e6
    return d


def train_case(R, variant, seed, batch_seed, num_items=None):
    cfg = tiny_config(variant)
    cat = syn.make_catalogue(2000, 1)
    batch = syn.make_train_batch(cat, 6, max_his_len=12, seed=batch_seed, median_len=6)
    extra = {} if num_items is None else {"num_items_in_batch": num_items}
    m, sd, shapes = build(R, variant, cfg, seed, "sdpa")
    with torch.no_grad():
        out = m(**batch, **extra)
        L = batch["input_ids"].shape[1]
        r = m.model.router(batch["input_ids"], cache_position=torch.arange(L))
    m2, _, _ = build(R, variant, cfg, seed, "eager")
    out2 = m2(**batch, **extra)
    out2.loss.backward()
    grads = [(k, (p.grad if p.grad is not None else torch.zeros_like(p))) for k, p in m2.named_parameters()
             if k != "lm_head.weight"]
    return {
        "variant": variant, "config": cfg_public(cfg), "weight_seed": seed, "weight_checksum": checksum(sd),
        "shapes": shapes, "temperature": 0.7, "batch": batch, "num_items_in_batch": num_items,
        "logits": out.logits.clone(), "loss": out.loss.double(), "loss_eager": out2.loss.detach().double(),
        "route": [t.clone() for t in r], "grads": grad_digest(grads),
        "embed_grad": dict(grads)["model.embed_tokens.weight"].clone(),
    }


LOGIT_STRIDE = 97      # full-size cases keep every 97th logit (the full tensors are 8 MB)


def mb_batch(vocab, behavior_tokens, n_items, rows, seed):
    """train_MB_decoder-shaped batch (multi-behaviour sequences without sessions, tasks/train_MB_decoder.py:317-365):
    rows of `n_items` items, each item = behaviour token + 4 code tokens, all rows full length."""
    rng = np.random.default_rng(seed)
    cat = syn.make_catalogue(20_000, 7)
    ct = cat.tokens()
    ids = np.empty((rows, n_items, 5), dtype=np.int64)
    for r in range(rows):
        it = rng.integers(0, cat.n_items, size=n_items)
        ids[r, :, 1:] = ct[it]
        ids[r, :, 0] = np.asarray(behavior_tokens)[rng.integers(0, len(behavior_tokens), size=n_items)]
    ids = torch.from_numpy(ids.reshape(rows, n_items * 5))
    labels = ids.clone()
    for t in behavior_tokens:
        labels[labels == t] = -100
    return {"input_ids": ids, "attention_mask": torch.ones_like(ids), "labels": labels}


def full_size_case(R, variant, seed, batch, max_his_len, vocab=1041, behavior_tokens=(526, 527, 528)):
    """Whole-model parity at BASELINE.json's shapes (all layers of the packaged config, L = 5 (max_his_len + 1)): loss,
    strided logits, gradient digests and the full embedding gradient of the unmodified reference."""
    cfg = ref_shim.reference_config(variant, vocab_size=vocab, behavior_tokens=behavior_tokens, max_his_len=max_his_len,
                                    model_max_length=max(1024, 5 * (max_his_len + 1)))
    m, sd, shapes = build(R, variant, cfg, seed, "sdpa")
    with torch.no_grad():
        out = m(**batch)
    m2, _, _ = build(R, variant, cfg, seed, "eager")
    out2 = m2(**batch)
    out2.loss.backward()
    grads = [(k, (p.grad if p.grad is not None else torch.zeros_like(p))) for k, p in m2.named_parameters()
             if k != "lm_head.weight"]
    pub = cfg_public(cfg)
    return {
        "variant": variant, "config": pub, "weight_seed": seed, "weight_checksum": checksum(sd), "shapes": shapes,
        "temperature": 0.7, "batch": batch, "num_items_in_batch": None,
        "logits_stride": LOGIT_STRIDE, "logits_samples": out.logits.reshape(-1)[::LOGIT_STRIDE].clone(),
        "logits_absmax": out.logits.abs().max().double(),
        "loss": out.loss.double(), "loss_eager": out2.loss.detach().double(), "grads": grad_digest(grads),
        "embed_grad": dict(grads)["model.embed_tokens.weight"].clone(),
    }


def decode_case(R, variant, seed, batch_seed, target_behavior, K):
    cfg = tiny_config(variant)
    cat = syn.make_catalogue(3000, 1)
    batch, targets = syn.make_eval_batch(cat, 4, max_his_len=12, target_behavior=target_behavior, seed=batch_seed,
                                         median_len=6)
    items = cat.item_sequences(target_behavior).tolist()
    trie = R["Trie"](items)
    last = set(t[-1] for t in items) | {cfg.pad_token_id}
    fn = R["prefix_allowed_tokens_fn_by_last_token"](trie, last)
    m, sd, shapes = build(R, variant, cfg, seed, "sdpa", temperature=1.0)
    with torch.no_grad():
        out = m.generate(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"],
                         session_ids=batch["session_ids"], extended_session_ids=batch["extended_session_ids"],
                         actions=batch["actions"], max_new_tokens=4, prefix_allowed_tokens_fn=fn, num_beams=K,
                         num_return_sequences=K, output_scores=True, return_dict_in_generate=True,
                         early_stopping=True, do_sample=False)
    # allowed-token goldens straight from the reference trie, for a few prefixes
    probes = []
    for it in items[:20]:
        for n in range(0, 5):
            probes.append((it[:n], sorted(trie.get(it[:n]))))
    # ranking goldens from the reference's ranking.py on the decoded ids (strings = joined token ids)
    pred = ["".join(f"<{t}>" for t in row[-4:].tolist()) for row in out.sequences]
    # make some targets hit: put the user's 2nd and 5th beams among the targets
    tg = []
    for b in range(4):
        tl = ["".join(f"<{t}>" for t in tup) for tup in targets[b]]
        tl += [pred[b * K + 1], pred[b * K + 4]] if b % 2 == 0 else []
        tg.append(tl)
    hits = R["ranking"].get_topk_results(pred, out.sequences_scores, tg, K)
    names = ["hit@1", "hit@5", "recall@5", "ndcg@5", "recall@10", "ndcg@10"]
    names = [n for n in names if int(n.split("@")[1]) <= K]
    met = R["ranking"].get_metrics_results(hits, names, tg)
    return {
        "variant": variant, "config": cfg_public(cfg), "weight_seed": seed, "weight_checksum": checksum(sd),
        "shapes": shapes, "batch": batch, "catalogue_seed": 1, "catalogue_size": 3000,
        "target_behavior": target_behavior, "num_beams": K, "sequences": out.sequences.clone(),
        "sequences_scores": out.sequences_scores.clone(), "trie_probes": probes,
        "rank_pred": pred, "rank_targets": tg, "rank_hits": hits, "rank_metrics": met, "rank_names": names,
    }


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    R = ref_shim.load_reference()
    torch.manual_seed(0)
    jobs = {
        "train_qwen3multi.pt": lambda: train_case(R, "Qwen3Multi", 42, 3),
        "train_qwen3multi_numitems.pt": lambda: train_case(R, "Qwen3Multi", 43, 4, num_items=777),
        "train_qwen3sessionmoe.pt": lambda: train_case(R, "Qwen3SessionMoe", 44, 5),
        "train_qwen3sessionmulti.pt": lambda: train_case(R, "Qwen3SessionMulti", 45, 6),
        "train_qwen3moe.pt": lambda: train_case(R, "Qwen3Moe", 47, 8),
        # BASELINE.json configs[1] shape (Qwen3Multi, 8 layers, max_his_len 100: L = 505, full-length rows)
        "train_qwen3multi_headline.pt": lambda: full_size_case(
            R, "Qwen3Multi", 48, syn.make_train_batch(syn.make_catalogue(50_000, 1234), 4, max_his_len=100, seed=11,
                                                      full_length=True), 100),
        # BASELINE.json configs[3] shape (train_MB_decoder: Qwen3Moe, 4 behaviour types, max_his_len 200: L = 1005)
        "train_qwen3moe_mb4.pt": lambda: full_size_case(
            R, "Qwen3Moe", 49, mb_batch(1042, (526, 527, 528, 1041), 201, 2, 12), 200, vocab=1042,
            behavior_tokens=(526, 527, 528, 1041)),
        "decode_qwen3multi_lvl2.pt": lambda: decode_case(R, "Qwen3Multi", 42, 5, 2, 8),
        "decode_qwen3multi_lvl1.pt": lambda: decode_case(R, "Qwen3Multi", 46, 7, 1, 6),
    }
    only = set(sys.argv[1:])                   # python -m oracle.make_golden [file.pt ...]: regenerate a subset
    for name, fn in jobs.items():
        if only and name not in only:
            continue
        obj = fn()
        path = os.path.join(GOLDEN, name)
        torch.save(obj, path)
        print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
