"""TEST INFRASTRUCTURE ONLY — never imported by the product (gamer_b200/).

Makes the *unmodified* reference classes under /root/reference importable with the installed
transformers 5.5.0 (the reference pins 4.51.0, requirements.txt:9).  It adds six names that carry no
arithmetic and maps one renamed kwarg.  /root/reference only exists in the build
container, so this module is used by oracle/make_golden.py and by the `needs_reference` tests, never on
the GPU box.
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("GAMER_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "SeqRec", "models", "generative"))


def install() -> None:
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import transformers
    import transformers.utils as tu
    import transformers.cache_utils as cu
    import transformers.models.qwen3.modeling_qwen3 as mq
    import transformers.models.qwen3_moe.modeling_qwen3_moe as mm
    from typing import TypedDict

    class _Kw(TypedDict, total=False):
        pass

    for mod, name, val in [
        (mq, "KwargsForCausalLM", _Kw), (mq, "QWEN3_INPUTS_DOCSTRING", ""),
        (mm, "KwargsForCausalLM", _Kw), (mm, "QWEN3_MOE_INPUTS_DOCSTRING", ""),
        (mm, "logger", transformers.utils.logging.get_logger("ref_shim")),
        (cu, "SlidingWindowCache", type("SlidingWindowCache", (), {})),
    ]:
        if not hasattr(mod, name):
            setattr(mod, name, val)
    if not hasattr(tu, "replace_return_docstrings"):
        tu.replace_return_docstrings = lambda **kw: (lambda f: f)
    if not getattr(mm.Qwen3MoeAttention.forward, "_gamer_shim", False):
        _orig = mm.Qwen3MoeAttention.forward

        def _fwd(self, hidden_states, position_embeddings=None, attention_mask=None,
                 past_key_values=None, past_key_value=None, **kw):
            return _orig(self, hidden_states, position_embeddings, attention_mask,
                         past_key_values if past_key_values is not None else past_key_value, **kw)

        _fwd._gamer_shim = True
        mm.Qwen3MoeAttention.forward = _fwd


def load_reference():
    """Import the reference's SeqRec package (from REFERENCE_ROOT) and return the model classes.

    The repo root also holds a `SeqRec/` facade (the drop-in module paths).  The reference must own the
    name `SeqRec` in this process (HF looks classes up through sys.modules), so this refuses to run if the
    facade was imported first; facade tests run in a subprocess.
    """
    install()
    mod = sys.modules.get("SeqRec")
    if mod is not None and not str(getattr(mod, "__file__", "") or "").startswith(REFERENCE_ROOT):
        paths = [str(p) for p in getattr(mod, "__path__", [])]
        if not any(p.startswith(REFERENCE_ROOT) for p in paths):
            raise RuntimeError("the repo's SeqRec facade is already imported; load the reference in a fresh process")
    if sys.path[0] != REFERENCE_ROOT:
        sys.path.insert(0, REFERENCE_ROOT)
    from SeqRec.models.generative.Qwen3Multi.model import Qwen3MultiWithTemperature
    from SeqRec.models.generative.Qwen3SessionMoe.model import Qwen3SessionMoeWithTemperature
    from SeqRec.models.generative.Qwen3SessionMulti.model import Qwen3SessionMultiWithTemperature
    from SeqRec.models.generative.Qwen3Moe.model import MyQwen3MoeForCausalLM, Qwen3MoeWithTemperature
    # 4.x-style list at Qwen3Moe/model.py:464; transformers 5.x expects {tied key: source key}
    MyQwen3MoeForCausalLM._tied_weights_keys = {"lm_head.weight": "model.embed_tokens.weight"}
    from SeqRec.generation.trie import Trie, prefix_allowed_tokens_fn_by_last_token
    from SeqRec.evaluation import ranking
    return dict(Qwen3Multi=Qwen3MultiWithTemperature, Qwen3SessionMoe=Qwen3SessionMoeWithTemperature,
                Qwen3SessionMulti=Qwen3SessionMultiWithTemperature, Qwen3Moe=Qwen3MoeWithTemperature, Trie=Trie,
                prefix_allowed_tokens_fn_by_last_token=prefix_allowed_tokens_fn_by_last_token,
                ranking=ranking)


def reference_config(name="Qwen3Multi", vocab_size=1041, behavior_tokens=(526, 527, 528), max_his_len=100,
                     num_layers=None, model_max_length=1024):
    """Config the way train_SMB_decoder.py:321-360 mutates it (no tokenizer / dataset needed)."""
    from transformers.models.qwen3_moe import Qwen3MoeConfig
    cfg = Qwen3MoeConfig.from_pretrained(os.path.join(REFERENCE_ROOT, "config", "s2s-models", name))
    cfg.vocab_size = vocab_size
    cfg.num_behavior = len(behavior_tokens)
    cfg.behavior_maps = {str(t): i for i, t in enumerate(behavior_tokens)}
    cfg.use_behavior_token = True
    cfg.num_positions = 5
    cfg.num_experts = 6
    cfg.n_positions = max_his_len + 1
    cfg.use_user_token = False
    cfg.model_max_length = model_max_length
    if num_layers is not None:
        cfg.num_hidden_layers = num_layers
        for key in ("sparse_layers_decoder", "behavior_injection_decoder", "cross_attention_decoder"):
            if hasattr(cfg, key):
                setattr(cfg, key, [i for i in getattr(cfg, key) if i < num_layers])
    return cfg
