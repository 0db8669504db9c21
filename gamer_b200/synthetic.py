"""ShortVideoAD-shaped synthetic sessions.

The reference's datasets are Git-LFS pointers, so inputs are generated as *token-id tensors* with the shapes and
semantics its dataset + collators produce (datasets/SMB_dataset.py:194-234,526-610; datasets/collator.py:47-107,
149-207):

  * item = 5 tokens [<behavior_x>, <a_i>, <b_j>, <c_k>, <d_l>]; ids: 0-13 stub specials (pad=bos=4, eos=8),
    <a_*> 14-269, <b_*> 270-525, <behavior_*> 526-528, <c_*> 529-784, <d_*> 785-1040  => V = 1041;
  * train row: last <= max_his_len+1 items, right-padded with 4; labels = ids with pad and behaviour tokens -> -100;
    actions = behaviour level per token, pad 100; session_ids pad 0; extended_session_ids pad 0;
  * eval row: last <= max_his_len items, LEFT-padded, target-behaviour token appended; session_ids/ext ids get
    max+1 appended, actions the target level (collator.py:180-201, tasks/test_SMB_decoder.py:105-117).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

PAD, EOS = 4, 8
A0, B0, BEH0, C0, D0 = 14, 270, 526, 529, 785
VOCAB = 1041
N_BEHAVIOR = 3
BEHAVIOR_TOKENS = (526, 527, 528)
BEHAVIOR_LEVEL = (0, 1, 2)            # behaviour b has level b; target behaviour = level 2
BEHAVIOR_PROB = (0.85, 0.12, 0.03)
TOKENS_PER_ITEM = 5


@dataclass
class Catalogue:
    codes: np.ndarray                 # [n_items, 4] codes in [0,256)

    @property
    def n_items(self) -> int:
        return self.codes.shape[0]

    def tokens(self) -> np.ndarray:
        """[n_items, 4] token ids (<a>,<b>,<c>,<d>)."""
        return self.codes + np.array([A0, B0, C0, D0], dtype=np.int64)

    def item_sequences(self, behavior: int) -> np.ndarray:
        """[n_items, 5] = behaviour token + 4 code tokens: the trie keys (tasks/test_SMB_decoder.py:489-494)."""
        t = self.tokens()
        return np.concatenate([np.full((t.shape[0], 1), BEHAVIOR_TOKENS[behavior], dtype=np.int64), t], axis=1)


def make_catalogue(n_items: int = 250_000, seed: int = 1234) -> Catalogue:
    """Unique 4-tuples over 4 codebooks x 256 codes, drawn uniformly without replacement."""
    rng = np.random.default_rng(seed)
    need = n_items
    seen = np.empty((0,), dtype=np.int64)
    while seen.size < need:
        draw = rng.integers(0, 256 ** 4, size=int((need - seen.size) * 1.2) + 16, dtype=np.int64)
        seen = np.unique(np.concatenate([seen, draw]))
    rng.shuffle(seen)
    flat = seen[:need]
    codes = np.stack([(flat >> 24) & 255, (flat >> 16) & 255, (flat >> 8) & 255, flat & 255], axis=1)
    return Catalogue(codes.astype(np.int64))


def _user_history(rng, cat: Catalogue, n_items: int):
    """(item index, behaviour, session id) per interaction."""
    # Zipf(1.05)-like popularity over the catalogue via inverse-CDF on ranks
    u = rng.random(n_items)
    ranks = np.minimum((cat.n_items ** u).astype(np.int64), cat.n_items - 1)
    beh = rng.choice(N_BEHAVIOR, size=n_items, p=BEHAVIOR_PROB)
    new_session = rng.random(n_items) < (1.0 / 8.0)
    new_session[0] = False
    sess = np.cumsum(new_session)
    return ranks, beh, sess


def _tokens_for(cat_tokens, items, beh):
    n = len(items)
    row = np.empty((n, TOKENS_PER_ITEM), dtype=np.int64)
    row[:, 0] = np.asarray(BEHAVIOR_TOKENS)[beh]
    row[:, 1:] = cat_tokens[items]
    return row.reshape(-1)


def _ext_session(sess):
    """remapped session rank * 5 + slot (SMB_dataset.py:206-222)."""
    change = np.concatenate([[True], sess[1:] != sess[:-1]])
    rank = np.cumsum(change) - 1
    return (rank[:, None] * TOKENS_PER_ITEM + np.arange(TOKENS_PER_ITEM)[None, :]).reshape(-1)


def _hist_len(rng, max_items, full_length, median=60.0):
    if full_length:
        return max_items
    n = int(np.clip(np.exp(rng.normal(np.log(median), 0.8)), 2, 400))
    return min(n, max_items)


def train_histories(cat: Catalogue, batch: int, max_his_len: int = 100, seed: int = 0, full_length: bool = False,
                    median_len: float = 60.0) -> list:
    """The raw histories behind `make_train_batch(same arguments)` — (item tokens [n, 4], behaviour [n], session [n]) per
    row, same RNG draws — i.e. the rows of a `collate.PackedSessions` store whose `collate_train` output IS that batch."""
    rng = np.random.default_rng(seed)
    ct = cat.tokens()
    out = []
    for _ in range(batch):
        n = _hist_len(rng, max_his_len + 1, full_length, median_len)
        items, beh, sess = _user_history(rng, cat, n)
        beh[-1] = N_BEHAVIOR - 1 if rng.random() < 0.5 else beh[-1]
        out.append((ct[items], beh, sess))
    return out


def make_train_batch(cat: Catalogue, batch: int, max_his_len: int = 100, seed: int = 0,
                     full_length: bool = False, median_len: float = 60.0) -> dict:
    """DecoderOnlyCollator-shaped training batch (right padding).  int64 CPU tensors."""
    rng = np.random.default_rng(seed)
    ct = cat.tokens()
    rows = []
    for _ in range(batch):
        n = _hist_len(rng, max_his_len + 1, full_length, median_len)
        items, beh, sess = _user_history(rng, cat, n)
        beh[-1] = N_BEHAVIOR - 1 if rng.random() < 0.5 else beh[-1]
        ids = _tokens_for(ct, items, beh)
        rows.append(dict(ids=ids, sess=np.repeat(sess, TOKENS_PER_ITEM), ext=_ext_session(sess),
                         act=np.repeat(np.asarray(BEHAVIOR_LEVEL)[beh], TOKENS_PER_ITEM)))
    L = max(len(r["ids"]) for r in rows)
    out = {k: np.zeros((batch, L), dtype=np.int64) for k in
           ("input_ids", "attention_mask", "labels", "session_ids", "extended_session_ids", "actions")}
    out["input_ids"][:] = PAD
    out["actions"][:] = 100
    for b, r in enumerate(rows):
        n = len(r["ids"])
        out["input_ids"][b, :n] = r["ids"]
        out["attention_mask"][b, :n] = 1
        out["session_ids"][b, :n] = r["sess"]
        out["extended_session_ids"][b, :n] = r["ext"]
        out["actions"][b, :n] = r["act"]
    labels = out["input_ids"].copy()
    labels[labels == PAD] = -100
    for t in BEHAVIOR_TOKENS:
        labels[labels == t] = -100
    out["labels"] = labels
    return {k: torch.from_numpy(v) for k, v in out.items()}


def make_eval_batch(cat: Catalogue, batch: int, max_his_len: int = 100, target_behavior: int = 2, seed: int = 0,
                    full_length: bool = False, median_len: float = 60.0) -> tuple[dict, list[list[tuple]]]:
    """DecoderOnlyTestCollator-shaped batch (left padding, target-behaviour token appended) + per-user target item
    token tuples (1-5 items)."""
    rng = np.random.default_rng(seed)
    ct = cat.tokens()
    rows, targets = [], []
    for _ in range(batch):
        n = _hist_len(rng, max_his_len, full_length, median_len)
        items, beh, sess = _user_history(rng, cat, n)
        ids = _tokens_for(ct, items, beh)
        sess_t = np.repeat(sess, TOKENS_PER_ITEM)
        ext = _ext_session(sess)
        rows.append(dict(
            ids=np.concatenate([ids, [BEHAVIOR_TOKENS[target_behavior]]]),
            sess=np.concatenate([sess_t, [sess_t.max() + 1]]),
            ext=np.concatenate([ext, [ext.max() + 1]]),
            act=np.concatenate([np.repeat(np.asarray(BEHAVIOR_LEVEL)[beh], TOKENS_PER_ITEM),
                                [BEHAVIOR_LEVEL[target_behavior]]])))
        n_t = int(rng.integers(1, 6))
        tgt = rng.integers(0, cat.n_items, size=n_t)
        targets.append([tuple(int(x) for x in ct[i]) for i in tgt])
    L = max(len(r["ids"]) for r in rows)
    out = {k: np.zeros((batch, L), dtype=np.int64) for k in
           ("input_ids", "attention_mask", "session_ids", "extended_session_ids", "actions")}
    out["input_ids"][:] = PAD
    out["actions"][:] = 100
    for b, r in enumerate(rows):
        n = len(r["ids"])
        out["input_ids"][b, L - n:] = r["ids"]
        out["attention_mask"][b, L - n:] = 1
        out["session_ids"][b, L - n:] = r["sess"]
        out["extended_session_ids"][b, L - n:] = r["ext"]
        out["actions"][b, L - n:] = r["act"]
    return {k: torch.from_numpy(v) for k, v in out.items()}, targets


def seeded_state_dict(shapes: dict, seed: int = 42, std: float = 0.05, norm_jitter: float = 0.1) -> dict:
    """Deterministic fp32 weights for parity runs, keyed like the reference's state dict.  Linear/embedding
    tensors ~ N(0, std) (larger than the reference's 0.02 init so logits are not near-uniform and beam scores are
    not near-tied); norm weights = 1 + N(0, norm_jitter).  Keys are visited in sorted order with one CPU
    generator, so the same (shapes, seed) gives the same bits wherever this torch build runs."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(shapes):
        shp = tuple(shapes[k])
        if len(shp) == 1:
            out[k] = 1.0 + norm_jitter * torch.randn(shp, generator=g)
        else:
            out[k] = std * torch.randn(shp, generator=g)
    return out


def make_mb_batch(cat: Catalogue, batch: int, max_his_len: int = 200, behavior_tokens=(526, 527, 528, 1041),
                  seed: int = 0) -> dict:
    """train_MB_decoder-shaped batch (multi-behaviour sequences WITHOUT sessions, tasks/train_MB_decoder.py:317-365;
    BASELINE.json configs[3]: four behaviour types): rows of max_his_len + 1 items, item = behaviour token + 4 code
    tokens, all rows full length; labels mask the behaviour tokens.  int64 CPU tensors."""
    rng = np.random.default_rng(seed)
    ct = cat.tokens()
    n = max_his_len + 1
    ids = np.empty((batch, n, TOKENS_PER_ITEM), dtype=np.int64)
    for r in range(batch):
        u = rng.random(n)
        ranks = np.minimum((cat.n_items ** u).astype(np.int64), cat.n_items - 1)
        ids[r, :, 1:] = ct[ranks]
        ids[r, :, 0] = np.asarray(behavior_tokens)[rng.integers(0, len(behavior_tokens), size=n)]
    ids = torch.from_numpy(ids.reshape(batch, n * TOKENS_PER_ITEM))
    labels = ids.clone()
    for t in behavior_tokens:
        labels[labels == t] = -100
    return {"input_ids": ids, "attention_mask": torch.ones_like(ids), "labels": labels}
