"""CPU: the drop-in Trie API and its flat CSR form against the oracle's trie and the golden probes from the reference."""
import numpy as np
import torch

from gamer_b200 import synthetic as syn
from gamer_b200.trie import Trie, flat_from_array, flat_from_dict, prefix_allowed_tokens_fn, \
    prefix_allowed_tokens_fn_by_last_token
from oracle import oracle_decode as od
from tests.helpers import load_golden


def _walk(flat, prefix):
    node = 0
    for t in prefix:
        s, e = int(flat.child_start[node]), int(flat.child_start[node + 1])
        toks = flat.child_tok[s:e].tolist()
        if t not in toks:
            return None
        node = int(flat.child_node[s + toks.index(t)])
    return node


def test_trie_matches_reference_probes():
    g = load_golden("decode_qwen3multi_lvl2.pt")
    cat = syn.make_catalogue(g["catalogue_size"], g["catalogue_seed"])
    items = cat.item_sequences(g["target_behavior"])
    trie = Trie(items.tolist())
    assert len(trie) == len(items)
    flat = trie.flat()
    for prefix, allowed in g["trie_probes"]:
        assert sorted(trie.get(prefix)) == allowed
        node = _walk(flat, prefix)
        assert node is not None and sorted(flat.children(node)) == allowed
    assert trie.get([1, 2, 3]) == []
    assert sorted(tuple(s) for s in trie) == sorted(set(tuple(r) for r in items.tolist()))


def test_flat_builders_agree_and_ragged():
    rng = np.random.default_rng(0)
    seqs = rng.integers(0, 7, size=(300, 4))
    a = flat_from_array(seqs)
    t = Trie(seqs.tolist())
    b = flat_from_dict(t.trie_dict)
    assert a.n_nodes == b.n_nodes and a.max_children == b.max_children
    # same language: every prefix has the same children in both forms
    for row in seqs[:50].tolist():
        for n in range(5):
            na, nb = _walk(a, row[:n]), _walk(b, row[:n])
            assert sorted(a.children(na)) == sorted(b.children(nb)) == sorted(t.get(row[:n]))
    ragged = Trie([[1, 2, 3], [1, 2], [4]])
    f = ragged.flat()
    assert sorted(f.children(0)) == [1, 4] and f.children(_walk(f, [1, 2])) == [3]


def test_prefix_fn_semantics():
    items = [[526, 20, 300, 600, 900], [526, 20, 301, 601, 901], [527, 21, 300, 600, 900]]
    trie = Trie(items)
    last = {900, 901, 4}
    fn = prefix_allowed_tokens_fn_by_last_token(trie, last)
    sent = torch.tensor([4, 4, 526, 20, 300, 600, 900, 526])
    assert sorted(fn(0, sent)) == [20]
    assert sorted(fn(0, torch.cat([sent, torch.tensor([20])]))) == [300, 301]
    assert fn(0, torch.cat([sent, torch.tensor([99])])) == []
    tree = od.PrefixTree(items)
    assert sorted(od.allowed_by_last_token(tree, last, sent.tolist())) == [20]
    assert sorted(prefix_allowed_tokens_fn(trie)(0, torch.tensor([526]))) == [20]
