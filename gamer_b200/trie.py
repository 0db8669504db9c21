"""Candidate prefix trie: the reference's Python API (SeqRec/generation/trie.py:5-104) plus a flat CSR form that lives
in HBM and is walked by the fused beam-step kernel (gamer_b200/csrc/decode.cu).

`Trie`, `prefix_allowed_tokens_fn` and `prefix_allowed_tokens_fn_by_last_token` keep the reference's names, arguments
and return values, so `tasks/test_SMB_decoder.py:467-501` works unchanged; the callables they return additionally
carry `.trie` / `.last_token_set`, which `generate()` uses to run the constraint on the GPU instead of calling back into
Python once per beam per step.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Iterable

import numpy as np
import torch


@dataclass
class FlatTrie:
    """CSR trie: node n has children [child_start[n], child_start[n+1]) — tokens ascending in child_tok, node ids in
    child_node.  Node 0 is the root."""
    child_start: torch.Tensor   # int32 [n_nodes + 1]
    child_tok: torch.Tensor     # int32 [n_edges]
    child_node: torch.Tensor    # int32 [n_edges]
    max_children: int
    n_nodes: int

    def to(self, device) -> "FlatTrie":
        return FlatTrie(self.child_start.to(device), self.child_tok.to(device), self.child_node.to(device),
                        self.max_children, self.n_nodes)

    def children(self, node: int) -> list[int]:
        s, e = int(self.child_start[node]), int(self.child_start[node + 1])
        return self.child_tok[s:e].tolist()


def flat_from_array(seqs: np.ndarray) -> FlatTrie:
    """Sort-based build for equal-length sequences [n, depth] (every catalogue item is behaviour token + 4 codes):
    O(n log n) numpy work, no Python loop over items."""
    seqs = np.asarray(seqs, dtype=np.int64)
    n, depth = seqs.shape
    order = np.lexsort(tuple(seqs[:, d] for d in reversed(range(depth))))
    s = seqs[order]
    toks, parents = [], []
    prev_ids = np.zeros(n, dtype=np.int64)       # node id of the length-(d) prefix of every row (root = 0)
    next_id = 1
    changed_any = np.zeros(n, dtype=bool)
    for d in range(depth):
        col_changed = np.ones(n, dtype=bool)
        col_changed[1:] = s[1:, d] != s[:-1, d]
        changed_any = changed_any | col_changed   # prefix of length d+1 differs from the previous row's
        changed_any[0] = True
        ids = next_id + np.cumsum(changed_any) - 1
        first = np.nonzero(changed_any)[0]
        toks.append(s[first, d])
        parents.append(prev_ids[first])
        next_id = int(ids[-1]) + 1
        prev_ids = ids
    tok = np.concatenate(toks)
    par = np.concatenate(parents)
    n_nodes = next_id
    counts = np.bincount(par, minlength=n_nodes)
    start = np.zeros(n_nodes + 1, dtype=np.int64)
    np.cumsum(counts, out=start[1:])
    # edges are already grouped by ascending parent id and ascending token (rows are lexicographically sorted)
    child_node = np.arange(1, n_nodes, dtype=np.int64)
    return FlatTrie(torch.from_numpy(start.astype(np.int32)), torch.from_numpy(tok.astype(np.int32)),
                    torch.from_numpy(child_node.astype(np.int32)), int(counts.max()) if n_nodes > 1 else 0, n_nodes)


def flat_from_dict(root: dict) -> FlatTrie:
    """Breadth-first flattening of a nested-dict trie (ragged sequences)."""
    start, tok, node = [0], [], []
    queue = [root]
    n_nodes = 1
    i = 0
    max_children = 0
    while i < len(queue):
        cur = queue[i]
        i += 1
        for t in sorted(cur):
            tok.append(int(t))
            node.append(n_nodes)
            n_nodes += 1
            queue.append(cur[t])
        max_children = max(max_children, len(cur))
        start.append(len(tok))
    return FlatTrie(torch.tensor(start, dtype=torch.int32), torch.tensor(tok, dtype=torch.int32),
                    torch.tensor(node, dtype=torch.int32), max_children, n_nodes)


class Trie:
    """Drop-in for SeqRec.generation.trie.Trie."""

    def __init__(self, sequences: Iterable[Iterable[int]] = ()):
        self.trie_dict: dict[int, dict] = {}
        self.len = 0
        self._flat: FlatTrie | None = None
        self._rect: list | None = []
        for sequence in sequences:
            self.add(sequence)
        self.append_trie: "Trie | None" = None
        self.bos_token_id = None

    def add(self, sequence):
        node = self.trie_dict
        seq = [int(t) for t in sequence]
        for t in seq:
            node = node.setdefault(t, {})
        self.len += 1
        self._flat = None
        if self._rect is not None:
            if self._rect and len(self._rect[0]) != len(seq):
                self._rect = None
            else:
                self._rect.append(seq)

    def get(self, prefix_sequence) -> list[int]:
        node = self.trie_dict
        for i, t in enumerate(prefix_sequence):
            nxt = node.get(int(t))
            if nxt is None:
                return self.append_trie.get(prefix_sequence) if self.append_trie else []
            node = nxt
        out = list(node.keys())
        if self.append_trie and self.bos_token_id in out:
            out.remove(self.bos_token_id)
            out += list(self.append_trie.trie_dict.keys())
        return out

    @staticmethod
    def load_from_dict(trie_dict: dict) -> "Trie":
        t = Trie()
        t.trie_dict = trie_dict
        t._rect = None
        t.len = sum(1 for _ in t)
        return t

    def __iter__(self):
        stack = [([], self.trie_dict)]
        while stack:
            prefix, node = stack.pop()
            if not node:
                yield prefix
            else:
                for tok in reversed(list(node)):
                    stack.append((prefix + [tok], node[tok]))

    def __len__(self) -> int:
        return self.len

    def __getitem__(self, value):
        return self.get(value)

    def flat(self) -> FlatTrie:
        if self.append_trie is not None:
            raise NotImplementedError("append_trie chaining has no flat form (unused by the SMB decoder tasks)")
        if self._flat is None:
            if self._rect:
                self._flat = flat_from_array(np.asarray(self._rect, dtype=np.int64))
            else:
                self._flat = flat_from_dict(self.trie_dict)
        return self._flat


class _PrefixFn:
    """Callable with the reference signature `(batch_id, sentence) -> list[int]` that also exposes its trie."""

    def __init__(self, trie: Trie, last_token_set: set[int] | None):
        self.trie = trie
        self.last_token_set = None if last_token_set is None else set(int(t) for t in last_token_set)

    def __call__(self, batch_id: int, sentence: torch.Tensor) -> list[int]:
        sentence = sentence.tolist()
        if self.last_token_set is None:
            return self.trie.get(sentence)
        index = len(sentence) - 1
        while index >= 0 and sentence[index] not in self.last_token_set:
            index -= 1
        return self.trie.get(sentence[index + 1:])


def prefix_allowed_tokens_fn(candidate_trie: Trie) -> Callable[[int, torch.Tensor], list[int]]:
    return _PrefixFn(candidate_trie, None)


def prefix_allowed_tokens_fn_by_last_token(candidate_trie: Trie, last_token_set: set[int]):
    return _PrefixFn(candidate_trie, last_token_set)
