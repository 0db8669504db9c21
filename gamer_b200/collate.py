"""Input pipeline for the hot path: batches built from PRE-TOKENISED arrays instead of
re-tokenising prompt strings per batch.

The reference assembles every batch on the CPU from Python strings: `SMBExplicitDatasetForDecoder` formats the history
as text, `DecoderOnlyCollator` / `DecoderOnlyTestCollator` run the HF tokenizer over it and pad
(SeqRec/datasets/SMB_dataset.py:194-234,526-610; SeqRec/datasets/collator.py:47-107,149-207).  All real tokens are
added tokens with fixed ids, so the whole data set reduces to three flat integer arrays + user offsets
(`PackedSessions`, ~7 bytes per interaction).  `collate_train` / `collate_eval` turn a set of users into the tensors the
collators emit — `input_ids, attention_mask, labels, session_ids, extended_session_ids, actions` — with a handful of
vectorised tensor ops when the store lives on the host, and with ONE kernel launch (`gamer_collate_sessions`,
csrc/collate.cu) when it lives on the GPU: pinned host memory -> one small H2D copy -> padding and expansion to 5 tokens
per item on the device.  The tensor-op version below is the specification the kernel is tested against, bit for bit
(tests/test_collate_gpu.py).

Semantics kept (checked against `gamer_b200.synthetic`, whose batches follow the collators):
  * history = the user's last `max_his_len` items (+ the target item when training: `max_his_len + 1` items);
  * item = [<behavior_x>, <a>, <b>, <c>, <d>]; train rows right-padded with pad=4, eval rows LEFT-padded and the
    target-behaviour token appended (collator.py:155,180-196; tasks/test_SMB_decoder.py:105-117);
  * labels = ids with pad and behaviour tokens -> -100 (collator.py:68-78);
  * session_ids per token (pad 0); extended_session_ids = 5 * session rank within the window + slot
    (SMB_dataset.py:206-222); eval appends max+1 to both; actions = behaviour level per token, pad 100 (collator.py:99,201).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

TOKENS_PER_ITEM = 5


@dataclass
class PackedSessions:
    """Flat pre-tokenised interaction store.  User u owns rows offsets[u] .. offsets[u+1]-1, oldest first."""
    item_tokens: torch.Tensor   # [T, 4] int32: the four semantic-code token ids of the item
    behavior: torch.Tensor      # [T] int16: behaviour index 0..n_behavior-1
    session: torch.Tensor       # [T] int32: session id, non-decreasing within a user
    offsets: torch.Tensor       # [N + 1] int64

    @property
    def n_users(self) -> int:
        return self.offsets.numel() - 1

    def to(self, device, non_blocking=False) -> "PackedSessions":
        return PackedSessions(*(t.to(device, non_blocking=non_blocking) for t in
                                (self.item_tokens, self.behavior, self.session, self.offsets)))

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.item_tokens, self.behavior, self.session, self.offsets))

    def pin_memory(self) -> "PackedSessions":
        return PackedSessions(*(t.pin_memory() for t in (self.item_tokens, self.behavior, self.session, self.offsets)))

    @staticmethod
    def from_histories(histories) -> "PackedSessions":
        """histories: iterable of (item_tokens [n,4], behavior [n], session [n]) per user."""
        toks, beh, sess, off = [], [], [], [0]
        for t, b, s in histories:
            toks.append(torch.as_tensor(t, dtype=torch.int32).view(-1, 4))
            beh.append(torch.as_tensor(b, dtype=torch.int16).view(-1))
            sess.append(torch.as_tensor(s, dtype=torch.int32).view(-1))
            off.append(off[-1] + toks[-1].shape[0])
        return PackedSessions(torch.cat(toks), torch.cat(beh), torch.cat(sess), torch.tensor(off, dtype=torch.int64))


def _window(store: PackedSessions, users: torch.Tensor, n_max: int, left_pad: bool, width: int | None = None):
    """Gather the last <= n_max interactions of every user into [B, n] grids.  Returns (row index grid, valid mask).
    `width` (items per row of the grid) defaults to the longest kept history, which costs one host sync; callers that
    know it (fixed-shape batches for CUDA-graph replay) pass it."""
    start, end = store.offsets[users], store.offsets[users + 1]
    n = torch.clamp(end - start, max=n_max)                                    # items kept per user
    width = int(n.max()) if width is None else int(width)
    t = torch.arange(width, device=users.device).unsqueeze(0)                  # [1, width]
    if left_pad:
        valid = t >= (width - n).unsqueeze(1)
        idx = end.unsqueeze(1) - width + t
    else:
        valid = t < n.unsqueeze(1)
        idx = (end - n).unsqueeze(1) + t
    return torch.where(valid, idx, torch.zeros_like(idx)), valid


def _expand(store, idx, valid, behavior_tokens, behavior_level, pad):
    """Item grids [B, n] -> token grids [B, 5 n]: ids, attention mask, session ids, extended session ids, actions."""
    B, n = idx.shape
    dev = idx.device
    beh = store.behavior[idx].long()
    beh_tok = torch.as_tensor(behavior_tokens, dtype=torch.int64, device=dev)[beh]
    level = torch.as_tensor(behavior_level, dtype=torch.int64, device=dev)[beh]
    ids = torch.cat([beh_tok.unsqueeze(-1), store.item_tokens[idx].long()], dim=-1)                  # [B, n, 5]
    sess = store.session[idx].long()
    # session rank inside the window: 0 for the first kept item, +1 at every change of session id
    change = torch.zeros_like(valid)
    change[:, 1:] = (sess[:, 1:] != sess[:, :-1]) & valid[:, 1:] & valid[:, :-1]
    rank = torch.cumsum(change.long(), dim=1)
    slot = torch.arange(TOKENS_PER_ITEM, device=dev).view(1, 1, -1)
    v5 = valid.unsqueeze(-1).expand(B, n, TOKENS_PER_ITEM)
    out = {
        "input_ids": torch.where(v5, ids, torch.full_like(ids, pad)),
        "attention_mask": v5.long(),
        "session_ids": torch.where(v5, sess.unsqueeze(-1).expand_as(ids), torch.zeros_like(ids)),
        "extended_session_ids": torch.where(v5, rank.unsqueeze(-1) * TOKENS_PER_ITEM + slot, torch.zeros_like(ids)),
        "actions": torch.where(v5, level.unsqueeze(-1).expand_as(ids), torch.full_like(ids, 100)),
    }
    return {k: v.reshape(B, n * TOKENS_PER_ITEM).contiguous() for k, v in out.items()}


_LUT_CACHE: dict = {}


def _device_collate(store, users, n_max, width, left_pad, behavior_tokens, behavior_level, pad, target_behavior, labels):
    """The whole batch in one launch (store on the GPU).  `width=None` costs one host sync for the longest kept history."""
    from . import kernels as K
    dev = store.offsets.device
    if width is None:
        n = torch.clamp(store.offsets[users + 1] - store.offsets[users], max=n_max)
        width = int(n.max())
    key = (tuple(int(t) for t in behavior_tokens), tuple(int(t) for t in behavior_level), str(dev))
    lut = _LUT_CACHE.get(key)
    if lut is None:
        lut = (torch.tensor(key[0], dtype=torch.int64, device=dev), torch.tensor(key[1], dtype=torch.int64, device=dev))
        if not torch.cuda.is_current_stream_capturing():
            _LUT_CACHE[key] = lut
    return K.collate_sessions(store, users, n_max, int(width), left_pad, lut[0], lut[1], pad, target_behavior, labels)


def collate_train(store: PackedSessions, users, max_his_len: int, behavior_tokens, behavior_level, pad: int = 4,
                  width: int | None = None) -> dict:
    """DecoderOnlyCollator (train split, only_train_response=False): the last max_his_len + 1 items, right-padded."""
    users = torch.as_tensor(users, dtype=torch.int64, device=store.offsets.device)
    if store.offsets.is_cuda:
        return _device_collate(store, users, max_his_len + 1, width, False, behavior_tokens, behavior_level, pad, -1, True)
    idx, valid = _window(store, users, max_his_len + 1, left_pad=False, width=width)
    out = _expand(store, idx, valid, behavior_tokens, behavior_level, pad)
    labels = out["input_ids"].clone()
    ignore = labels == pad
    for t in behavior_tokens:
        ignore |= labels == t
    out["labels"] = labels.masked_fill(ignore, -100)
    return out


def collate_eval(store: PackedSessions, users, max_his_len: int, target_behavior: int, behavior_tokens, behavior_level,
                 pad: int = 4, width: int | None = None) -> dict:
    """DecoderOnlyTestCollator + the target-behaviour append of test_SMB_decoder.py:105-117: the last max_his_len items,
    LEFT-padded, then one more column holding the target behaviour token (session max+1, extended max+1, its level)."""
    users = torch.as_tensor(users, dtype=torch.int64, device=store.offsets.device)
    if store.offsets.is_cuda:
        return _device_collate(store, users, max_his_len, width, True, behavior_tokens, behavior_level, pad,
                               int(target_behavior), False)
    idx, valid = _window(store, users, max_his_len, left_pad=True, width=width)
    out = _expand(store, idx, valid, behavior_tokens, behavior_level, pad)
    B = users.numel()
    dev = users.device
    col = lambda v: torch.full((B, 1), v, dtype=torch.int64, device=dev)
    out["input_ids"] = torch.cat([out["input_ids"], col(behavior_tokens[target_behavior])], 1)
    out["attention_mask"] = torch.cat([out["attention_mask"], col(1)], 1)
    out["session_ids"] = torch.cat([out["session_ids"], out["session_ids"].max(dim=1, keepdim=True)[0] + 1], 1)
    out["extended_session_ids"] = torch.cat([out["extended_session_ids"],
                                             out["extended_session_ids"].max(dim=1, keepdim=True)[0] + 1], 1)
    out["actions"] = torch.cat([out["actions"], col(behavior_level[target_behavior])], 1)
    return out
