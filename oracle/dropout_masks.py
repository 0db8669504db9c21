"""TEST INFRASTRUCTURE ONLY — numpy restatement of the dropout-mask generator of the CUDA path
(gamer_b200/csrc/common.cuh: philox4x32_7, drop_keep8, make_drop; gamer_b200/csrc/attention_tc.cu: keep_word).

The reference draws its dropout masks from torch's generator (nn.Dropout, SDPA dropout_p: Qwen3Multi/model.py:139,177,
217,235,241; Qwen3Moe/FFN.py:23-26), so masks can never be bit-identical to it; parity with dropout ON is therefore
checked by handing the *same* Philox masks to the oracle (oracle_model.forward(..., drop=OracleDropout(...))) and
comparing losses and gradients under the usual tolerances.  Philox4x32 is the published Random123 generator
(Salmon et al., SC'11); the 10-round variant is pinned to the Random123 known-answer vectors in
tests/test_dropout_cpu.py, the kernels use 7 rounds (the fewest that pass BigCrush).
"""
from __future__ import annotations

import numpy as np
import torch

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
HIDDEN_TAG = 0x64726F70
MASK32 = np.uint64(0xFFFFFFFF)
# site kinds within a layer (gamer_b200.engine.SITE_*): site = layer * 8 + kind
SITE_SELF_P, SITE_SELF_OUT, SITE_CROSS_P, SITE_CROSS_OUT, SITE_FFN_INNER, SITE_FFN_OUT = range(6)


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=7):
    """Counter words c0..c3 (broadcastable uint arrays), key (k0, k1) python ints -> 4 uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*(np.asarray(c, dtype=np.uint64) & MASK32 for c in (c0, c1, c2, c3)))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(rounds):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def _key(seed: int, offset: int):
    return seed & 0xFFFFFFFF, ((seed >> 32) & 0xFFFFFFFF) ^ (offset & 0xFFFFFFFF)


def _thresh(p: float, bits: int):
    full = 1 << bits
    t = min(int(float(np.float32(p)) * full), full - 1)      # the C struct carries p as a float
    return t, float(np.float32(full / (full - t)))


def hidden_keep(seed, offset, site, rows, cols, p):
    """-> (keep bool [rows, cols], scale).  16 random bits per element; one Philox call per 8 columns."""
    assert cols % 8 == 0
    t, scale = _thresh(p, 16)
    k0, k1 = _key(seed, offset)
    r = np.arange(rows, dtype=np.uint64).reshape(rows, 1)
    c8 = np.arange(cols // 8, dtype=np.uint64).reshape(1, cols // 8)
    w = philox4x32(c8, r, site, HIDDEN_TAG, k0, k1)
    vals = np.empty((rows, cols // 8, 8), dtype=np.uint32)
    for q in range(4):
        vals[:, :, 2 * q] = w[q] & np.uint32(0xFFFF)
        vals[:, :, 2 * q + 1] = w[q] >> np.uint32(16)
    return torch.from_numpy((vals >= t).reshape(rows, cols)), scale


def attn_keep_words(seed, offset, site, B, n_q, L, p):
    """-> (uint32 [B*n_q, L, ceil(L/32)] keep words, scale): the words the forward kernel stores (attention_tc.cu:
    keep_word).  Word (bh, i, jw) covers keys [32 jw, 32 jw + 32): one Philox call gives the four high bit planes r4..r7,
    the low planes r0..r3 are those words rotated left by 5, 13, 21, 29; lane bit c of r_k is bit k of an 8-bit number R_c,
    and the lane is dropped iff R_c < thresh."""
    t, scale = _thresh(p, 8)
    k0, k1 = _key(seed, offset)
    nw = (L + 31) // 32
    bh = np.arange(B * n_q, dtype=np.uint64).reshape(B * n_q, 1, 1)
    i = np.arange(L, dtype=np.uint64).reshape(1, L, 1)
    jw = np.arange(nw, dtype=np.uint64).reshape(1, 1, nw)
    hi = philox4x32(jw, i, bh, site, k0, k1)                        # the four high bit planes
    rotl = lambda x, n: ((x << np.uint32(n)) | (x >> np.uint32(32 - n))).astype(np.uint32)
    r = tuple(rotl(hi[k].astype(np.uint32), 5 + 8 * k) for k in range(4)) + tuple(h.astype(np.uint32) for h in hi)
    lanes = np.zeros((B * n_q, L, nw, 32), dtype=np.uint32)       # R_c per lane
    bit = np.arange(32, dtype=np.uint32)
    for k in range(8):
        lanes |= ((r[k][..., None] >> bit) & np.uint32(1)) << np.uint32(k)
    keep_lane = lanes >= t                                           # lane = bit position of the keep word
    words = (keep_lane.astype(np.uint32) << bit).sum(axis=-1, dtype=np.uint64).astype(np.uint32)
    return words, scale


def attn_keep(seed, offset, site, B, n_q, L, p):
    """-> (keep bool [B, n_q, L, L(keys)], scale).  Key 32 jw + 4 g + e of a word sits at bit g + 8 e."""
    words, scale = attn_keep_words(seed, offset, site, B, n_q, L, p)
    nw = words.shape[-1]
    c = np.arange(32)
    bitpos = ((c >> 2) + 8 * (c & 3)).astype(np.uint32)              # key column within the block -> bit
    keep = ((words[..., None] >> bitpos) & np.uint32(1)).astype(bool).reshape(B, n_q, L, nw * 32)[..., :L]
    return torch.from_numpy(np.ascontiguousarray(keep)), scale


class OracleDropout:
    """Mask provider for oracle_model.forward(..., drop=...): the masks of one forward pass of the CUDA path."""

    def __init__(self, seed: int, offset: int, p_hidden: float, p_attn: float):
        self.seed, self.offset, self.p_hidden, self.p_attn = seed, offset, p_hidden, p_attn

    def hidden(self, layer, kind, B, S, width):
        """multiplicative mask Z [B, S, width] (0 or 1/keep_prob), rows indexed by the flat token row b * S + s"""
        if self.p_hidden <= 0:
            return None
        keep, scale = hidden_keep(self.seed, self.offset, layer * 8 + kind, B * S, width, self.p_hidden)
        return keep.view(B, S, width).float() * scale

    def attn(self, layer, kind, B, n_q, L):
        if self.p_attn <= 0:
            return None
        keep, scale = attn_keep(self.seed, self.offset, layer * 8 + kind, B, n_q, L, self.p_attn)
        return keep.float() * scale
