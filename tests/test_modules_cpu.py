"""CPU: the `SeqRec/modules` Transformer/Attention interfaces (north_star) — same constructor and
forward signatures, same parameter names, and (in the build container, where /root/reference exists) the same outputs
as the reference's modules for the same state dict and the additive mask of SeqModel.get_attention_mask."""
import importlib.util
import inspect
import os

import pytest
import torch

from gamer_b200 import modules as M

REF = "/root/reference/SeqRec/modules/layers/transformer.py"


def _mask(B, L, causal):
    keep = torch.ones(B, L, dtype=torch.bool)
    keep[0, L - 3:] = False
    ext = keep[:, None, None, :]
    if causal:
        ext = ext & torch.tril(torch.ones(L, L, dtype=torch.bool))[None, None]
    return torch.where(ext, 0.0, -10000.0)


def test_signatures_and_parameter_names():
    layer = M.TransformerEncoderLayer(64, 4, 128, 0.1, "gelu", 1e-12)
    enc = M.TransformerEncoder(layer, 2)
    assert list(inspect.signature(M.MultiHeadAttention.forward).parameters) == ["self", "input_tensor", "attention_mask"]
    assert list(inspect.signature(M.TransformerEncoder.forward).parameters)[:3] == ["self", "hidden_states", "attention_mask"]
    names = set(enc.state_dict())
    for k in ("layer.0.multi_head_attention.query.weight", "layer.1.multi_head_attention.LayerNorm.bias",
              "layer.0.multi_head_attention.dense.weight", "layer.1.feed_forward.dense_1.weight",
              "layer.0.feed_forward.dense_2.bias", "layer.0.feed_forward.LayerNorm.weight"):
        assert k in names, k
    x = torch.randn(2, 9, 64)
    assert enc.eval()(x, _mask(2, 9, True)).shape == x.shape
    with pytest.raises(ValueError):
        M.MultiHeadAttention(10, 3, 0.1, 1e-5)
    with pytest.raises(AttributeError):                      # reference quirk: residual=False has no LayerNorm
        M.FeedForward(8, 16, 0.1, "relu", 1e-5, residual=False)(torch.randn(1, 2, 8))


@pytest.mark.needs_reference
@pytest.mark.parametrize("causal", [False, True])
def test_outputs_match_reference_modules(causal):
    spec = importlib.util.spec_from_file_location("ref_transformer", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.manual_seed(0)
    r_enc = ref.TransformerEncoder(ref.TransformerEncoderLayer(64, 4, 128, 0.2, "gelu", 1e-12), 3).eval()
    m_enc = M.TransformerEncoder(M.TransformerEncoderLayer(64, 4, 128, 0.2, "gelu", 1e-12), 3).eval()
    assert set(r_enc.state_dict()) == set(m_enc.state_dict())
    m_enc.load_state_dict(r_enc.state_dict())
    x = torch.randn(3, 11, 64)
    mask = _mask(3, 11, causal)
    with torch.no_grad():
        a, b = r_enc(x, mask), m_enc(x, mask)
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-5), (a - b).abs().max()
    r_att = ref.MultiHeadAttention(64, 8, 0.0, 1e-5).eval()
    m_att = M.MultiHeadAttention(64, 8, 0.0, 1e-5).eval()
    m_att.load_state_dict(r_att.state_dict())
    with torch.no_grad():
        assert torch.allclose(r_att(x, mask), m_att(x, mask), rtol=1e-5, atol=1e-5)
