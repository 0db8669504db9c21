"""Shared test helpers: golden loading, oracle Spec/weights from a golden's public config."""
import os

import torch

from gamer_b200 import synthetic as syn
from oracle import oracle_model as om

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def spec_from_golden(g, temperature=1.0):
    c = g["config"]
    variant = g["variant"]
    cross = tuple(c["cross_attention_decoder"]) if variant not in ("Qwen3SessionMoe", "Qwen3Moe") else ()
    return om.Spec(variant=variant, vocab_size=c["vocab_size"], hidden=c["hidden_size"], n_q=c["num_attention_heads"],
                   n_kv=c["num_key_value_heads"], head_dim=c["head_dim"], inter=c["intermediate_size"],
                   n_layers=c["num_hidden_layers"], beh_dim=c["behavior_embedding_dim"], n_behavior=c["num_behavior"],
                   n_positions=c["num_positions"], n_experts=c["num_experts"],
                   sparse_layers=tuple(c["sparse_layers_decoder"]), inject_layers=tuple(c["behavior_injection_decoder"]),
                   cross_layers=cross, behavior_maps={int(k): int(v) for k, v in c["behavior_maps"].items()},
                   pad=c["pad_token_id"], eos=c["eos_token_id"], eps=c["rms_norm_eps"], rope_theta=c["rope_theta"],
                   temperature=temperature)


def weights_from_golden(g, requires_grad=False):
    sd = syn.seeded_state_dict(g["shapes"], seed=g["weight_seed"])
    sd["lm_head.weight"] = sd["model.embed_tokens.weight"]
    chk = float(sum(v.double().abs().sum() for k, v in sorted(sd.items()) if k != "lm_head.weight"))
    assert abs(chk - g["weight_checksum"]) <= 1e-6 * abs(chk), "seeded weights drifted from the golden's checksum"
    if requires_grad:
        W = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k != "lm_head.weight"}
        W["lm_head.weight"] = W["model.embed_tokens.weight"]
        return W
    return sd
