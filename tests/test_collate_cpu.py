"""CPU: the pre-tokenised input pipeline (gamer_b200/collate.py) reproduces (a) the batches of gamer_b200.synthetic, which
follow the reference collators, tensor for tensor, and (b) — in the build container — the reference's own
`_generate_session_ids / _generate_extended_session_ids / _generate_actions` (SeqRec/datasets/SMB_dataset.py:194-234)
on random histories longer than max_his_len (truncation to the last items, train vs test mode)."""
import ast
import types

import numpy as np
import pytest
import torch

from gamer_b200 import collate as C
from gamer_b200 import synthetic as syn

BT, BL = syn.BEHAVIOR_TOKENS, syn.BEHAVIOR_LEVEL


def _replay_train_histories(cat, batch, max_his_len, seed, median_len=60.0):
    """The same RNG draws as synthetic.make_train_batch, returned as raw histories."""
    rng = np.random.default_rng(seed)
    ct = cat.tokens()
    out = []
    for _ in range(batch):
        n = syn._hist_len(rng, max_his_len + 1, False, median_len)
        items, beh, sess = syn._user_history(rng, cat, n)
        beh[-1] = syn.N_BEHAVIOR - 1 if rng.random() < 0.5 else beh[-1]
        out.append((ct[items], beh, sess))
    return out


def test_train_batches_equal_synthetic():
    cat = syn.make_catalogue(3000, 1)
    for seed, mhl in ((3, 12), (11, 40)):
        want = syn.make_train_batch(cat, 9, max_his_len=mhl, seed=seed, median_len=8)
        store = C.PackedSessions.from_histories(_replay_train_histories(cat, 9, mhl, seed, median_len=8))
        got = C.collate_train(store, range(9), mhl, BT, BL, pad=syn.PAD)
        assert set(got) == set(want)
        for k in want:
            assert torch.equal(got[k], want[k]), k


def test_eval_batches_equal_synthetic():
    cat = syn.make_catalogue(3000, 1)
    seed, mhl, tgt = 21, 14, 1
    want, _ = syn.make_eval_batch(cat, 7, max_his_len=mhl, target_behavior=tgt, seed=seed, median_len=7)
    rng = np.random.default_rng(seed)
    ct = cat.tokens()
    hist = []
    for _ in range(7):
        n = syn._hist_len(rng, mhl, False, 7)
        items, beh, sess = syn._user_history(rng, cat, n)
        hist.append((ct[items], beh, sess))
        rng.integers(0, cat.n_items, size=int(rng.integers(1, 6)))          # the target draws of make_eval_batch
    got = C.collate_eval(C.PackedSessions.from_histories(hist), range(7), mhl, tgt, BT, BL, pad=syn.PAD)
    for k in want:
        assert torch.equal(got[k], want[k]), k


def test_truncation_keeps_the_last_items_and_subsets_of_users():
    cat = syn.make_catalogue(500, 2)
    rng = np.random.default_rng(0)
    ct = cat.tokens()
    hist = []
    for n in (3, 30, 17, 1):
        items, beh, sess = syn._user_history(rng, cat, n)
        hist.append((ct[items], beh, sess))
    store = C.PackedSessions.from_histories(hist)
    got = C.collate_train(store, [1, 3], 9, BT, BL)
    assert got["input_ids"].shape == (2, 50)                                  # 10 items x 5 tokens
    t, b, s = hist[1]
    assert got["input_ids"][0].view(10, 5)[:, 1:].tolist() == t[-10:].tolist()
    assert got["session_ids"][0].view(10, 5)[:, 0].tolist() == s[-10:].tolist()
    assert got["attention_mask"][1].sum().item() == 5 and got["labels"][1, 5:].eq(-100).all()
    assert got["extended_session_ids"][0, :5].tolist() == [0, 1, 2, 3, 4]      # ranks restart inside the window


@pytest.mark.needs_reference
@pytest.mark.parametrize("mode", ["train", "test"])
def test_session_and_action_arrays_match_reference_functions(mode):
    src = open("/root/reference/SeqRec/datasets/SMB_dataset.py").read()
    fns = {}
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.FunctionDef) and node.name in ("_generate_session_ids", "_generate_extended_session_ids",
                                                               "_generate_actions"):
            ns = {}
            exec(compile(ast.Module([node], []), "ref", "exec"), ns)           # the reference's own function bodies
            fns[node.name] = ns[node.name]
    assert len(fns) == 3
    mhl = 6
    stub = types.SimpleNamespace(max_his_len=mhl, mode=mode, token_count=lambda: 5, behavior_level={0: 0, 1: 1, 2: 2})
    rng = np.random.default_rng(5)
    cat = syn.make_catalogue(300, 3)
    ct = cat.tokens()
    for n in (2, 7, 25):
        items, beh, sess = syn._user_history(rng, cat, n)
        sess = sess + 3                                                       # raw ids need not start at zero
        store = C.PackedSessions.from_histories([(ct[items], beh, sess)])
        if mode == "train":
            got = C.collate_train(store, [0], mhl, BT, BL)
            cut = slice(None)
        else:
            got = C.collate_eval(store, [0], mhl, 2, BT, BL)
            cut = slice(None, -1)                                             # without the appended target column
        assert got["session_ids"][0, cut].tolist() == fns["_generate_session_ids"](stub, sess.tolist())
        assert got["extended_session_ids"][0, cut].tolist() == fns["_generate_extended_session_ids"](stub, sess.tolist())
        assert got["actions"][0, cut].tolist() == fns["_generate_actions"](stub, beh.tolist())
