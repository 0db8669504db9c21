"""Drop-in model classes: same names, constructor, forward/generate signatures and state-dict keys as the reference's
SeqRec/models/generative/{Qwen3Multi,Qwen3SessionMoe,Qwen3SessionMulti}/model.py, with the arithmetic done by the
sm_100a kernels behind include/gamer_b200.h (via gamer_b200.engine).  There is no PyTorch/CPU fallback: calling
forward on CPU tensors raises.

The nn.Module tree below exists only to hold parameters under the reference's checkpoint key names;
the modules' own forward methods are never called.
"""
from __future__ import annotations

import torch
from torch import nn
from transformers import PreTrainedModel
from transformers.modeling_outputs import CausalLMOutputWithPast
from transformers.models.qwen3_moe import Qwen3MoeConfig

from . import engine as E

try:  # transformers >= 5: flag-aware initialisers (skip parameters already loaded from a checkpoint)
    from transformers import initialization as _init
except ImportError:  # transformers 4.x (the reference's pin)
    from torch.nn import init as _init


class _Weight(nn.Module):
    """Parameter holder named like Qwen3RMSNorm (`.weight`, ones-initialised)."""

    def __init__(self, n):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(n))


class _Attention(nn.Module):
    # mirrors Qwen3MultiAttention.__init__ (Qwen3Multi/model.py:28-67)
    def __init__(self, cfg, is_cross: bool):
        super().__init__()
        hd = cfg.head_dim
        H, nq, nkv = cfg.hidden_size, cfg.num_attention_heads, cfg.num_key_value_heads
        self.q_proj = nn.Linear(H, nq * hd, bias=False)
        self.k_proj = nn.Linear(H, nkv * hd, bias=False)
        self.v_proj = nn.Linear(H, nkv * hd, bias=False)
        self.o_proj = nn.Linear(nq * hd, H, bias=False)
        self.q_norm = _Weight(hd)
        self.k_norm = _Weight(hd)
        if is_cross:
            bd = cfg.behavior_embedding_dim
            self.q_behavior_embedding = nn.Embedding(cfg.num_behavior + 1, nq * bd)
            self.k_behavior_embedding = nn.Embedding(cfg.num_behavior + 1, nkv * bd)
            self.v_behavior_embedding = nn.Embedding(cfg.num_behavior + 1, nkv * bd)
            self.gating = nn.Linear(H, H, bias=False)


class _Expert(nn.Module):
    # MyQwen3MoeMLP (Qwen3Moe/FFN.py:8-23)
    def __init__(self, cfg, inject: bool):
        super().__init__()
        k = cfg.moe_intermediate_size + (cfg.behavior_embedding_dim if inject else 0)
        self.gate_proj = nn.Linear(k, cfg.intermediate_size, bias=False)
        self.up_proj = nn.Linear(k, cfg.intermediate_size, bias=False)
        self.down_proj = nn.Linear(cfg.intermediate_size, cfg.moe_intermediate_size, bias=False)


class _SparseMLP(nn.Module):
    # MyQwen3SparseMLP (Qwen3Moe/FFN.py:30-51)
    def __init__(self, cfg, sparse: bool, inject: bool):
        super().__init__()
        if sparse:
            self.experts = nn.ModuleDict({f"expert_{i}": _Expert(cfg, inject) for i in range(cfg.num_experts)})
        else:
            self.mlp = _Expert(cfg, inject)
        if inject:
            self.behavior_embedding = nn.Embedding(cfg.num_behavior + 1, cfg.behavior_embedding_dim)


class _Layer(nn.Module):
    def __init__(self, cfg, idx: int, variant: str):
        super().__init__()
        sparse = idx in cfg.sparse_layers_decoder
        inject = idx in cfg.behavior_injection_decoder
        cross = variant not in E.NO_CROSS_VARIANTS and idx in (getattr(cfg, "cross_attention_decoder", None) or [])
        self.self_attn = _Attention(cfg, is_cross=False)
        if cross:
            self.cross_attn = _Attention(cfg, is_cross=True)
            self.post_self_attention_layernorm = _Weight(cfg.hidden_size)
        self.mlp = _SparseMLP(cfg, sparse, inject)
        self.input_layernorm = _Weight(cfg.hidden_size)
        if variant in E.NO_CROSS_VARIANTS:
            self.post_attention_layernorm = _Weight(cfg.hidden_size)
        else:
            self.post_cross_attention_layernorm = _Weight(cfg.hidden_size)


class _Backbone(nn.Module):
    def __init__(self, cfg, variant: str):
        super().__init__()
        self.embed_tokens = nn.Embedding(cfg.vocab_size, cfg.hidden_size, cfg.pad_token_id)
        self.layers = nn.ModuleList([_Layer(cfg, i, variant) for i in range(cfg.num_hidden_layers)])
        self.norm = _Weight(cfg.hidden_size)


class _GamerCausalLM(PreTrainedModel):
    config_class = Qwen3MoeConfig
    base_model_prefix = "model"
    _tied_weights_keys = {"lm_head.weight": "model.embed_tokens.weight"}
    _no_split_modules = ["_Layer"]
    supports_gradient_checkpointing = False
    VARIANT = "Qwen3Multi"

    def __init__(self, config: Qwen3MoeConfig):
        assert "num_positions" in config and isinstance(config.num_positions, int), \
            "Config must have 'num_positions' attribute."
        assert "model_max_length" in config and isinstance(config.model_max_length, int), \
            "Config must have 'model_max_length' attribute."
        if getattr(config, "mlp_type", None) != "Qwen3":
            raise NotImplementedError("gamer_b200 implements mlp_type='Qwen3' (MyQwen3SparseMLP); the PBATransformer "
                                      "FFN variant is outside the hot path (SURVEY.md §8)")
        if getattr(config, "Moe_behavior_only", False):
            # Qwen3Multi/router.py:31-48: two experts only (behaviour token vs code tokens); the packaged configs and both
            # SMB tasks run with False.  Silently using the per-position routing instead would train a different model.
            raise NotImplementedError("gamer_b200 implements the per-position expert routing (Moe_behavior_only=False)")
        super().__init__(config)
        self.model = _Backbone(config, self.VARIANT)
        self.vocab_size = config.vocab_size
        self.lm_head = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        self.post_init()
        self.temperature = 1.0
        self._pack = None
        self._pack_key = None
        self._lut = None
        self.grad_hooks = {}
        self._drop_seed = None          # None: derived from torch.initial_seed() and the rank at first use
        self._drop_calls = 0

    # ---- HF plumbing ---------------------------------------------------------------------------------------------
    def _init_weights(self, module):
        # N(0, initializer_range) for Linear/Embedding, zeroed pad row, ones for the norms — as Qwen3PreTrainedModel
        std = self.config.initializer_range
        if isinstance(module, nn.Linear):
            _init.normal_(module.weight, mean=0.0, std=std)
        elif isinstance(module, nn.Embedding):
            loaded = getattr(module.weight, "_is_hf_initialized", False)
            _init.normal_(module.weight, mean=0.0, std=std)
            if module.padding_idx is not None and not loaded:
                with torch.no_grad():
                    module.weight[module.padding_idx].zero_()
        elif isinstance(module, _Weight):
            _init.ones_(module.weight)

    def get_input_embeddings(self):
        return self.model.embed_tokens

    def set_input_embeddings(self, value):
        self.model.embed_tokens = value

    def get_output_embeddings(self):
        return self.lm_head

    def set_output_embeddings(self, new):
        self.lm_head = new

    def set_hyper(self, temperature: float):
        self.temperature = temperature

    # ---- engine glue ---------------------------------------------------------------------------------------------
    @property
    def arch(self) -> E.Arch:
        a = E.Arch.from_config(self.config, self.VARIANT)
        a.vocab = self.model.embed_tokens.weight.shape[0]
        return a

    def _named_weights(self):
        sd = dict(self.named_parameters())
        sd.setdefault("lm_head.weight", self.model.embed_tokens.weight)
        return sd

    def check_token_ids(self):
        """nn.Embedding raises IndexError on a token id outside the vocabulary; the gather kernel cannot raise, it flags
        the id in a device word.  This reads the word (one host sync) and raises: call it wherever the host synchronises
        anyway (logging steps, end of an epoch); `generate` does so itself."""
        from . import kernels as K
        K.check_token_ids(self.device)

    def set_dropout_seed(self, seed: int):
        """Fix the Philox seed of the training-mode dropout masks (default: torch.initial_seed() mixed with the rank)."""
        self._drop_seed = int(seed)
        self._drop_calls = 0

    def _drop_config(self):
        """(seed, p_hidden, p_attn) of the training-mode dropout, or None when the model is in eval mode / p = 0."""
        d = self._next_drop(advance=False)
        return None if d is None else (d.seed, d.p_hidden, d.p_attn)

    def _next_drop(self, advance=True):
        """Dropout context of the next training-mode forward (nn.Dropout(config.dropout_rate) / SDPA
        dropout_p=config.attention_dropout in the reference, Qwen3Multi/model.py:139,177); None in eval mode."""
        p_h = float(getattr(self.config, "dropout_rate", 0.0) or 0.0)
        p_a = float(getattr(self.config, "attention_dropout", 0.0) or 0.0)
        if not self.training or (p_h <= 0.0 and p_a <= 0.0):
            return None
        if self._drop_seed is None:
            rank = torch.distributed.get_rank() if torch.distributed.is_available() and torch.distributed.is_initialized() else 0
            self._drop_seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + 0xD1B54A32D192ED03 * (rank + 1)) & (2 ** 64 - 1)
        d = E.DropCtx(self._drop_seed, self._drop_calls, p_h, p_a)
        if advance:
            self._drop_calls += 1
        return d

    def _get_pack(self, arch):
        """bf16 operand pack, rebuilt whenever any master weight changed (optimizer step, load_state_dict)."""
        params = list(self.parameters())
        key = (tuple(p._version for p in params), params[0].device, arch.vocab)
        if self._pack is None or self._pack_key != key:
            with torch.no_grad():
                self._pack = E.Pack(arch, self._named_weights())
            self._pack_key = key
        if self._lut is None or self._lut.device != params[0].device or self._lut.numel() != arch.vocab:
            self._lut = E.behaviour_lut(arch, params[0].device)
        return self._pack, self._lut

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None,
                labels=None, use_cache=None, output_attentions=None, output_hidden_states=None, cache_position=None,
                logits_to_keep=0, session_ids=None, extended_session_ids=None, actions=None, **kwargs):
        if (input_ids is None) ^ (inputs_embeds is not None):
            raise ValueError("You must specify exactly one of input_ids or inputs_embeds")
        if inputs_embeds is not None:
            raise NotImplementedError("gamer_b200 fuses the embedding gather with the router: pass input_ids")
        if past_key_values is not None or cache_position is not None:
            raise NotImplementedError("incremental decoding is served by .generate() (constrained beam search); "
                                      "forward() is the full-sequence path")
        if output_attentions:
            raise NotImplementedError("attention probabilities are never materialised by the fused kernel")
        if not input_ids.is_cuda:
            raise RuntimeError("gamer_b200 has no CPU path: move the model and inputs to a CUDA device")
        if not self.lm_head.weight.data_ptr() == self.model.embed_tokens.weight.data_ptr():
            raise RuntimeError("lm_head and embed_tokens must stay tied (config.tie_word_embeddings)")
        arch = self.arch
        pack, lut = self._get_pack(arch)
        input_ids = input_ids.contiguous()
        B, L = input_ids.shape
        meta = E.make_meta(arch, input_ids, attention_mask, actions, session_ids, extended_session_ids)
        num_items = kwargs.get("num_items_in_batch", None)
        want_grad = labels is not None and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        loss, logits = None, None
        drop = self._next_drop()
        if want_grad:
            shifted = E.shift_labels(labels)
            inv_norm = self._inv_norm(shifted, num_items)
            names = E.param_names(arch)
            W = self._named_weights()
            loss = E.DecoderLossFunction.apply(arch, pack, meta, lut, input_ids, shifted, inv_norm, float(self.temperature),
                                               self.grad_hooks, drop, *[W[n] for n in names])
            if getattr(self.config, "gamer_return_train_logits", False):
                with torch.no_grad():
                    hidden, _ = E.forward_stack(arch, pack, input_ids, meta, lut, save=False, drop=drop)
                    logits = E.lm_head_logits(arch, pack, hidden, 1.0 / self.temperature).view(B, L, -1)
        else:
            with torch.no_grad():
                hidden, _ = E.forward_stack(arch, pack, input_ids, meta, lut, save=False, drop=drop)
                hidden = hidden.view(B, L, -1)
                if isinstance(logits_to_keep, int) and logits_to_keep > 0:
                    hidden_k = hidden[:, -logits_to_keep:, :]
                elif isinstance(logits_to_keep, int):
                    hidden_k = hidden
                else:
                    hidden_k = hidden[:, logits_to_keep, :]
                Lk = hidden_k.shape[1]
                alpha = 1.0 / self.temperature if labels is not None else 1.0      # Q7: logits /= T only with labels
                logits = E.lm_head_logits(arch, pack, hidden_k.reshape(B * Lk, -1).contiguous(), alpha).view(B, Lk, -1)
                if labels is not None:
                    shifted = E.shift_labels(labels)
                    inv_norm = self._inv_norm(shifted, num_items)
                    loss = E.lm_head_loss(arch, pack, hidden.reshape(B * L, -1), shifted, inv_norm, float(self.temperature))
        return CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=None, hidden_states=None, attentions=None)

    @staticmethod
    def _inv_norm(shifted, num_items):
        """1/N for the mean over non-ignored labels, or 1/num_items_in_batch (HF Trainer's accumulation-aware sum)."""
        if num_items is None:
            n = (shifted != -100).sum().clamp(min=1).float()
        elif torch.is_tensor(num_items):
            n = num_items.to(shifted.device).float()
        else:
            n = torch.tensor(float(num_items), device=shifted.device)
        return (1.0 / n).reshape(1)

    @torch.no_grad()
    def generate(self, input_ids=None, attention_mask=None, session_ids=None, extended_session_ids=None, actions=None,
                 max_new_tokens=4, prefix_allowed_tokens_fn=None, num_beams=1, num_return_sequences=None,
                 output_scores=True, return_dict_in_generate=True, early_stopping=True, do_sample=False,
                 candidate_trie=None, **kwargs):
        from .generation import constrained_beam_search
        if do_sample:
            raise NotImplementedError("the evaluation path is deterministic constrained beam search (do_sample=False)")
        nret = num_beams if num_return_sequences is None else num_return_sequences
        return constrained_beam_search(self, input_ids, attention_mask, session_ids, extended_session_ids, actions,
                                       max_new_tokens, prefix_allowed_tokens_fn, candidate_trie, num_beams, nret,
                                       return_dict_in_generate)


class Qwen3MultiWithTemperature(_GamerCausalLM):
    """Drop-in for SeqRec.models.generative.Qwen3Multi.Qwen3MultiWithTemperature (Qwen3Multi/model.py:883-1013)."""
    VARIANT = "Qwen3Multi"


class Qwen3SessionMoeWithTemperature(_GamerCausalLM):
    """Drop-in for SeqRec.models.generative.Qwen3SessionMoe.Qwen3SessionMoeWithTemperature
    (Qwen3SessionMoe/model.py:590-735)."""
    VARIANT = "Qwen3SessionMoe"


class Qwen3MoeWithTemperature(_GamerCausalLM):
    """Drop-in for SeqRec.models.generative.Qwen3Moe.Qwen3MoeWithTemperature (Qwen3Moe/model.py:604-640), the backbone
    of `train_MB_decoder` (multi-behaviour sequences without sessions): HF causal + padding mask, token-position RoPE,
    position-routed expert FFN.  `session_ids` / `actions` are accepted and ignored, as the reference's **kwargs do."""
    VARIANT = "Qwen3Moe"


class Qwen3SessionMultiWithTemperature(_GamerCausalLM):
    """Drop-in for SeqRec.models.generative.Qwen3SessionMulti.Qwen3SessionMultiWithTemperature."""
    VARIANT = "Qwen3SessionMulti"
