"""CPU: the numpy restatement of the dropout-mask generator (oracle/dropout_masks.py) — Philox4x32 known-answer vectors
(Random123's kat_vectors: 10 rounds, and the 7-round zero vector), mask statistics, and the oracle's mask injection."""
import numpy as np
import torch

from oracle import dropout_masks as dm
from oracle import oracle_model as om


def test_philox_known_answers():
    kat10 = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
             ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
             ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
              (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat10:
        got = tuple(int(x) for x in dm.philox4x32(*ctr, *key, rounds=10))
        assert got == want, (ctr, [hex(g) for g in got])
    got7 = tuple(int(x) for x in dm.philox4x32(0, 0, 0, 0, 0, 0, rounds=7))
    assert got7 == (0x5f6fb709, 0x0d893f64, 0x4f121f81, 0x4f730a48)


def test_mask_rates_and_determinism():
    keep, scale = dm.hidden_keep(1234, 7, 13, 512, 256, 0.2)
    assert keep.shape == (512, 256) and abs(scale - 65536 / (65536 - 13107)) < 1e-6
    rate = keep.float().mean().item()
    assert abs(rate - 0.8) < 5e-3, rate
    keep2, _ = dm.hidden_keep(1234, 7, 13, 512, 256, 0.2)
    assert torch.equal(keep, keep2)
    other, _ = dm.hidden_keep(1234, 8, 13, 512, 256, 0.2)            # another step offset: independent mask
    assert abs((keep == other).float().mean().item() - (0.64 + 0.04)) < 1e-2
    ka, sa = dm.attn_keep(99, 0, 2, 2, 6, 70, 0.2)
    assert ka.shape == (2, 6, 70, 70) and abs(sa - 256 / 205) < 1e-6
    assert abs(ka.float().mean().item() - 205 / 256) < 5e-3
    # rows / heads / sites decorrelated
    assert abs((ka[0, 0] == ka[0, 1]).float().mean().item() - (0.8008 ** 2 + 0.1992 ** 2)) < 3e-2


def test_oracle_dropout_injection():
    """oracle forward with the Philox masks injected: finite, differs from the dropout-free loss, deterministic in
    (seed, offset), and every mask has unit mean (E[dropout(x)] = x)."""
    from tests.helpers import load_golden, spec_from_golden, weights_from_golden
    g = load_golden("train_qwen3multi.pt")
    spec = spec_from_golden(g, g["temperature"])
    W = weights_from_golden(g)
    with torch.no_grad():
        base = om.forward(spec, W, **g["batch"])["loss"].item()
        d1 = om.forward(spec, W, **g["batch"], drop=dm.OracleDropout(42, 0, 0.2, 0.2))["loss"].item()
        d1b = om.forward(spec, W, **g["batch"], drop=dm.OracleDropout(42, 0, 0.2, 0.2))["loss"].item()
        d2 = om.forward(spec, W, **g["batch"], drop=dm.OracleDropout(42, 1, 0.2, 0.2))["loss"].item()
    assert abs(base - g["loss"].item()) < 1e-5
    assert np.isfinite(d1) and d1 == d1b and d1 != base and d1 != d2
    z = dm.OracleDropout(42, 0, 0.2, 0.2).hidden(3, dm.SITE_FFN_OUT, 8, 100, 256)
    assert abs(z.mean().item() - 1.0) < 5e-3
    za = dm.OracleDropout(42, 0, 0.2, 0.2).attn(3, dm.SITE_SELF_P, 2, 6, 120)
    assert abs(za.mean().item() - 1.0) < 5e-3
