"""Signature-compatible stand-ins for the reference's BERT-style building blocks
(SeqRec/modules/layers/transformer.py:12-183): `MultiHeadAttention(input_tensor, attention_mask)`, `FeedForward`,
`TransformerEncoderLayer`, `TransformerEncoder(hidden_states, attention_mask)`, with the reference's parameter names so
its state dicts load unchanged.

These blocks belong to the discriminative recommenders (SASRec, BERT4Rec, ...: SURVEY.md §2.1 rows 16-17), which are
OUTSIDE the accelerated hot path; `north_star` only asks that the Transformer/Attention interfaces stay importable
with unchanged signatures.  They run on stock PyTorch ops (fused SDPA with the additive
[B,1,L,L] / [B,1,1,L] mask of `SeqModel.get_attention_mask`), not on the sm_100a kernels of this package.
"""
from __future__ import annotations

import copy
from typing import Callable

import torch
import torch.nn.functional as F
from torch import nn

_ACT = {"gelu": F.gelu, "relu": F.relu, "swish": F.silu, "tanh": torch.tanh, "sigmoid": torch.sigmoid, "elu": F.elu}


class MultiHeadAttention(nn.Module):
    """Post-LN self attention: LayerNorm(x + dropout(dense(softmax(QK^T / sqrt(d) + mask) V)))."""
    _PROJ = ("query", "key", "value")

    def __init__(self, embed_dim: int, num_heads: int, dropout: float, layer_norm_eps: float):
        super().__init__()
        head, rest = divmod(embed_dim, num_heads)
        if rest:
            raise ValueError("The hidden size (%d) is not a multiple of the number of attention heads (%d)"
                             % (embed_dim, num_heads))
        self.num_attention_heads, self.attention_head_size, self.all_head_size = num_heads, head, num_heads * head
        for name in self._PROJ:                                   # query / key / value: the reference's parameter names
            setattr(self, name, nn.Linear(embed_dim, self.all_head_size))
        self.attn_dropout = nn.Dropout(dropout)
        self.dense = nn.Linear(embed_dim, embed_dim)
        self.LayerNorm = nn.LayerNorm(embed_dim, eps=layer_norm_eps)
        self.out_dropout = nn.Dropout(dropout)

    def transpose_for_scores(self, x: torch.Tensor) -> torch.Tensor:
        """[B, L, heads * d] -> [B, heads, L, d]"""
        B, L, _ = x.shape
        return x.view(B, L, self.num_attention_heads, self.attention_head_size).transpose(1, 2)

    def forward(self, input_tensor: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
        q, k, v = (self.transpose_for_scores(getattr(self, name)(input_tensor)) for name in self._PROJ)
        bias = attention_mask if attention_mask is None else attention_mask.to(q.dtype)     # additive, broadcastable
        p_drop = self.attn_dropout.p if self.training else 0.0
        mixed = F.scaled_dot_product_attention(q, k, v, attn_mask=bias, dropout_p=p_drop)   # scale = d ** -0.5
        mixed = mixed.transpose(1, 2).flatten(2)
        return self.LayerNorm(input_tensor + self.out_dropout(self.dense(mixed)))


class FeedForward(nn.Module):
    """dense_2(act(dense_1(x))).  As in the reference (transformer.py:115-124) the dropout + residual LayerNorm branch
    is guarded by `if not self.residual`, while those sub-modules only exist when `residual` is True: the default
    residual=True returns the bare projection, residual=False raises AttributeError.  Reproduced, not repaired."""

    def __init__(self, d_model: int, dim_feedforward: int, dropout: float,
                 activation: str | Callable[[torch.Tensor], torch.Tensor], layer_norm_eps: float, residual: bool = True):
        super().__init__()
        self.dense_1, self.dense_2 = nn.Linear(d_model, dim_feedforward), nn.Linear(dim_feedforward, d_model)
        self.intermediate_act_fn = self.get_hidden_act(activation) if isinstance(activation, str) else activation
        self.residual = residual
        if residual:
            self.LayerNorm, self.dropout = nn.LayerNorm(d_model, eps=layer_norm_eps), nn.Dropout(dropout)

    @staticmethod
    def get_hidden_act(act: str):
        return _ACT[act]

    def forward(self, input_tensor: torch.Tensor) -> torch.Tensor:
        y = self.dense_2(self.intermediate_act_fn(self.dense_1(input_tensor)))
        return y if self.residual else self.LayerNorm(input_tensor + self.dropout(y))


class TransformerEncoderLayer(nn.Module):
    def __init__(self, d_model: int, nhead: int, dim_feedforward: int = 2048, dropout: float = 0.1,
                 activation: str | Callable[[torch.Tensor], torch.Tensor] = F.relu, layer_norm_eps: float = 1e-5) -> None:
        super().__init__()
        self.multi_head_attention = MultiHeadAttention(d_model, nhead, dropout, layer_norm_eps)
        self.feed_forward = FeedForward(d_model, dim_feedforward, dropout, activation, layer_norm_eps)

    def forward(self, hidden_states: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
        return self.feed_forward(self.multi_head_attention(hidden_states, attention_mask))


class TransformerEncoder(nn.Module):
    """`num_layers` independent copies of `encoder_layer`, applied in order with the same mask."""

    def __init__(self, encoder_layer: nn.Module, num_layers: int):
        super().__init__()
        self.layer = nn.ModuleList(copy.deepcopy(encoder_layer) for _ in range(num_layers))

    def forward(self, hidden_states, attention_mask: torch.Tensor, **kwargs):
        for block in self.layer:
            hidden_states = block(hidden_states, attention_mask, **kwargs)
        return hidden_states
