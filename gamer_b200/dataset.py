"""The reference's SMB data files as pre-tokenised arrays: loader, train-time augmentation, splits.

The reference (SeqRec/datasets/SMB_dataset.py) loads per-user JSON lists, formats every history as a string of added
tokens and lets the HF tokenizer turn it back into ids for every batch.  All real tokens are ADDED tokens whose ids are
fixed by construction — `tokenizer.add_tokens(sorted(new_tokens))` on the 14-entry base vocabulary of
config/s2s-models/*/vocab.json (tasks/train_SMB_decoder.py:251, SMB_dataset.py:357-368) — so the id of a token string is
`14 + rank in sorted(new tokens)` and the whole data set becomes a `PackedSessions` store that `collate.py` expands on
the device.  Nothing here runs per batch.

  load_smb_files      <dataset>.SMB.{inter,behavior,session}.json, <dataset><index_file>, <dataset>.behavior_level.json
                      (SMB_dataset.py:72-152): histories, session-based valid / test positions, behaviour levels,
                      token ids, the item catalogue (for the candidate trie, tasks/test_SMB_decoder.py:467-501)
  train_store         SMBExplicitDatasetForDecoder._process_train_data (:585-610) with `augment` (smb_explicit_decoder_4:
                      augment = 4): the user's history up to the validation session plus <= `augment` down-sampled copies
                      (:540-584); the drops replay numpy's global stream seeded with 42, i.e. the reference's own draws
  eval_store          _process_valid_test_data / _process_test_data (:293-343): history before the held-out session and
                      the session's items per behaviour as targets
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field

import numpy as np
import torch

from .collate import PackedSessions

BASE_VOCAB = 14      # config/s2s-models/*/vocab.json: ids 0..13 (pad = 4 <|endoftext|>, eos = 8 <|im_end|>)


@dataclass
class SMBData:
    users: list                      # user keys in file order
    items: list                      # per user: np.int64 [n] item keys (index into `catalogue`)
    behaviors: list                  # per user: np.int16 [n] behaviour index (order of behavior_level.json)
    sessions: list                   # per user: np.int32 [n] session ids normalised to start at 0 (:93-94)
    valid_pos: np.ndarray            # [N] first interaction of the second-to-last session (-1: fewer than 2 sessions)
    test_pos: np.ndarray             # [N] first interaction of the last session
    behavior_names: list
    behavior_level: list             # level per behaviour index
    behavior_tokens: list            # token id of <behavior_x> per behaviour index
    target_behavior: int             # index of the (single) max-level behaviour
    catalogue: np.ndarray            # [n_items, sole_item_len] token ids of every item of the index file
    item_row: dict = field(default_factory=dict)   # item key (str) -> row of `catalogue`
    vocab_size: int = 0

    def item_sequences(self, behavior: int) -> np.ndarray:
        """[n_items, 1 + sole_item_len]: behaviour token + code tokens of every item (the candidate set of the trie)."""
        beh = np.full((self.catalogue.shape[0], 1), self.behavior_tokens[behavior], dtype=np.int64)
        return np.concatenate([beh, self.catalogue], axis=1)


def load_smb_files(data_path: str, dataset: str, index_file: str = ".index.json") -> SMBData:
    root = os.path.join(data_path, dataset)
    rd = lambda suffix: json.load(open(os.path.join(root, dataset + suffix)))
    inters, behaviors, sessions = rd(".SMB.inter.json"), rd(".SMB.behavior.json"), rd(".SMB.session.json")
    indices, level = rd(index_file), rd(".behavior_level.json")
    lens = {len(v) for v in indices.values()}
    if len(lens) != 1:
        raise ValueError(f"All indices must have the same length, but got lengths: {lens}")
    names = list(level.keys())
    top = [b for b in names if level[b] == max(level.values())]
    if len(top) != 1:
        raise ValueError(f"Expected exactly one target behavior with max level, but found {len(top)}: {top}")
    # added-token ids: 14 + rank in the sorted set of index tokens and behaviour tokens
    new_tokens = sorted({t for idx in indices.values() for t in idx} | {f"<behavior_{b}>" for b in names})
    tid = {t: BASE_VOCAB + i for i, t in enumerate(new_tokens)}
    keys = list(indices.keys())
    item_row = {k: i for i, k in enumerate(keys)}
    catalogue = np.asarray([[tid[t] for t in indices[k]] for k in keys], dtype=np.int64)
    b_idx = {b: i for i, b in enumerate(names)}
    users, it, be, se, vpos, tpos = [], [], [], [], [], []
    for uid in inters:
        s = np.asarray(sessions[uid], dtype=np.int64)
        s = s - s.min()
        uniq = np.unique(s)
        tpos.append(int(np.where(s == uniq[-1])[0].min()))
        vpos.append(int(np.where(s == uniq[-2])[0].min()) if len(uniq) >= 2 else -1)
        users.append(uid)
        it.append(np.asarray([item_row[str(i)] for i in inters[uid]], dtype=np.int64))
        be.append(np.asarray([b_idx[b] for b in behaviors[uid]], dtype=np.int16))
        se.append(s.astype(np.int32))
    return SMBData(users, it, be, se, np.asarray(vpos), np.asarray(tpos), names, [int(level[b]) for b in names],
                   [tid[f"<behavior_{b}>"] for b in names], b_idx[top[0]], catalogue, item_row,
                   BASE_VOCAB + len(new_tokens))


def augment_drops(behaviors: np.ndarray, levels, augment: int, rng) -> list:
    """SMBExplicitDatasetForDecoder._augment_interactions (:540-584) for one history: keep-masks of the down-sampled copies.
    For ratio r = 1/augment .. 1 every non-target behaviour b of level l loses int(count_b * r / (l + 1)) interactions,
    drawn without replacement by `rng.choice` in the reference's order (behaviours in behavior_level.json order); a copy
    with fewer than two interactions is skipped."""
    if not augment:
        return []
    top = max(levels)
    pos = {b: np.nonzero(behaviors == b)[0] for b in range(len(levels))}
    masks = []
    for ratio in np.arange(1, augment + 1) / augment:
        keep = np.ones(len(behaviors), dtype=bool)
        for b, lv in enumerate(levels):
            if lv == top or len(pos[b]) == 0:
                continue
            n_drop = int(len(pos[b]) * (ratio / (lv + 1)))
            if n_drop > 0:
                keep[rng.choice(pos[b].tolist(), n_drop, replace=False)] = False
        if keep.sum() >= 2:
            masks.append(keep)
    return masks


def _pack(data: SMBData, rows) -> PackedSessions:
    """rows: iterable of (user index, index array into that user's history)."""
    hist = []
    for u, sel in rows:
        hist.append((data.catalogue[data.items[u][sel]], data.behaviors[u][sel], data.sessions[u][sel]))
    return PackedSessions.from_histories(hist)


def train_store(data: SMBData, augment: int | None = None, seed: int = 42) -> PackedSessions:
    """One training sequence per user (interactions before the validation session; its last item is the target) plus the
    augmented copies.  `set_seed(42)` precedes the loop in the reference (:586): the drops replay that stream."""
    rng = np.random.RandomState(seed)
    rows = []
    for u in range(len(data.users)):
        n = int(data.valid_pos[u])
        if n <= 0:
            continue
        base = np.arange(n)
        rows.append((u, base))
        for keep in augment_drops(data.behaviors[u][:n], data.behavior_level, augment or 0, rng):
            rows.append((u, base[keep]))
    return _pack(data, rows)


def valid_store(data: SMBData) -> PackedSessions:
    """_process_valid_data (:271-291): for every interaction of the validation session, the history before that session
    followed by the interaction as the target (the per-epoch eval_loss set of the training task)."""
    rows = []
    for u in range(len(data.users)):
        pos = int(data.valid_pos[u])
        if pos < 0:
            continue
        for i in range(pos, int(data.test_pos[u])):
            rows.append((u, np.concatenate([np.arange(pos), [i]])))
    return _pack(data, rows)


def eval_store(data: SMBData, mode: str = "test"):
    """History before the held-out session of every user and, per user, {behaviour index: [item token tuples]} of that
    session (test: the last session; valid: the one before it, users with a single session skipped)."""
    rows, targets = [], []
    for u in range(len(data.users)):
        if mode == "test":
            lo, hi = int(data.test_pos[u]), len(data.items[u])
        else:
            if data.valid_pos[u] < 0:
                continue
            lo, hi = int(data.valid_pos[u]), int(data.test_pos[u])
        rows.append((u, np.arange(lo)))
        tg = {}
        for i in range(lo, hi):
            tg.setdefault(int(data.behaviors[u][i]), []).append(tuple(int(t) for t in data.catalogue[data.items[u][i]]))
        targets.append(tg)
    return _pack(data, rows), targets


def write_synthetic_files(data_path: str, dataset: str, n_users: int = 64, n_items: int = 500, seed: int = 0,
                          behaviors=("click", "cart", "buy"), codebook: int = 256) -> None:
    """A small data set in the reference's file format (tests / the CLI's --data_path smoke runs)."""
    rng = np.random.default_rng(seed)
    root = os.path.join(data_path, dataset)
    os.makedirs(root, exist_ok=True)
    idx = {str(i): [f"<{c}_{int(rng.integers(0, codebook))}>" for c in "abcd"] for i in range(n_items)}
    inter, beh, sess, tim = {}, {}, {}, {}
    for u in range(n_users):
        n_s = int(rng.integers(3, 7))
        per = rng.integers(2, 9, size=n_s)
        s = np.repeat(np.arange(n_s) + int(rng.integers(0, 5)), per)
        n = len(s)
        inter[str(u)] = rng.integers(0, n_items, size=n).tolist()
        b = rng.choice(len(behaviors), size=n, p=[0.8, 0.14, 0.06][:len(behaviors)] if len(behaviors) == 3 else None)
        beh[str(u)] = [behaviors[i] for i in b]
        sess[str(u)] = s.tolist()
        tim[str(u)] = [f"2024-01-01 00:{(i // 60) % 60:02d}:{i % 60:02d}" for i in range(n)]
    dump = lambda suffix, obj: json.dump(obj, open(os.path.join(root, dataset + suffix), "w"))
    dump(".SMB.inter.json", inter)
    dump(".SMB.behavior.json", beh)
    dump(".SMB.session.json", sess)
    dump(".SMB.time.json", tim)
    dump(".index.json", idx)
    dump(".behavior_level.json", {b: i for i, b in enumerate(behaviors)})


def synthetic_corpus(n_users: int, n_items: int, max_items: int, seed: int = 0, median_len: float = 60.0,
                     full_length: bool = False) -> SMBData:
    """ShortVideoAD-shaped synthetic users (gamer_b200.synthetic: Zipf item popularity, 85/12/3 % behaviours, a new
    session every ~8 interactions) in the same form the file loader returns, so the tasks run one code path.  Every user
    gets at least three sessions (train / validation / test split by session, SMB_dataset.py:88-105)."""
    from . import synthetic as syn
    rng = np.random.default_rng(seed)
    cat = syn.make_catalogue(n_items, 1234)
    items, beh, sess, vpos, tpos = [], [], [], [], []
    for _ in range(n_users):
        n = max(6, syn._hist_len(rng, max_items, full_length, median_len))
        it, b, s = syn._user_history(rng, cat, n)
        # force two more session boundaries near the end so that valid / test sessions exist
        cut_t = n - int(rng.integers(1, 4))
        cut_v = max(1, cut_t - int(rng.integers(1, 4)))
        s = np.asarray(s, dtype=np.int64)
        s[cut_v:] += 1
        s[cut_t:] += 1
        b = np.asarray(b)
        b[-1] = syn.N_BEHAVIOR - 1 if rng.random() < 0.5 else b[-1]
        uniq = np.unique(s)
        items.append(np.asarray(it, dtype=np.int64))
        beh.append(b.astype(np.int16))
        sess.append((s - s.min()).astype(np.int32))
        tpos.append(int(np.where(s == uniq[-1])[0].min()))
        vpos.append(int(np.where(s == uniq[-2])[0].min()))
    names = [f"behavior_{i}" for i in range(syn.N_BEHAVIOR)]
    return SMBData([str(u) for u in range(n_users)], items, beh, sess, np.asarray(vpos), np.asarray(tpos), names,
                   list(syn.BEHAVIOR_LEVEL), list(syn.BEHAVIOR_TOKENS), syn.N_BEHAVIOR - 1, cat.tokens().astype(np.int64),
                   {str(i): i for i in range(n_items)}, syn.VOCAB)
