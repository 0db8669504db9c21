"""CPU: the reference's data files -> PackedSessions -> device collate, pinned to samples frozen from the UNMODIFIED
`SMBExplicitDatasetForDecoder(augment=4)` (oracle/make_dataset_golden.py): token ids, the 4x augmentation's draws,
session / extended-session / action arrays, test histories and per-behaviour targets — all bit-exact."""
import json
import os
import re

import numpy as np
import torch

from gamer_b200 import collate
from gamer_b200 import dataset as ds

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataset_smb.json")
TOK = re.compile(r"<[^>]+>")


def _setup(tmp_path):
    g = json.load(open(GOLDEN))
    ds.write_synthetic_files(str(tmp_path), "toy", **g["spec"])
    data = ds.load_smb_files(str(tmp_path), "toy", ".index.json")
    tid = {t: ds.BASE_VOCAB + i for i, t in enumerate(g["new_tokens"])}
    return g, data, tid


def test_token_ids_follow_the_sorted_added_tokens(tmp_path):
    g, data, tid = _setup(tmp_path)
    assert data.vocab_size == ds.BASE_VOCAB + len(g["new_tokens"])
    assert data.behavior_tokens == [tid[f"<behavior_{b}>"] for b in data.behavior_names]
    idx = json.load(open(os.path.join(tmp_path, "toy", "toy.index.json")))
    for k, toks in list(idx.items())[:50]:
        assert data.catalogue[data.item_row[k]].tolist() == [tid[t] for t in toks]
    seqs = data.item_sequences(data.target_behavior)
    assert seqs.shape == (len(idx), 5) and (seqs[:, 0] == data.behavior_tokens[data.target_behavior]).all()


def test_train_samples_with_4x_augmentation_match_reference(tmp_path):
    g, data, tid = _setup(tmp_path)
    store = ds.train_store(data, augment=g["augment"])
    assert store.n_users == len(g["train"])          # same users kept, same copies skipped (< 2 interactions)
    batch = collate.collate_train(store, torch.arange(store.n_users), g["max_his_len"], data.behavior_tokens,
                                  data.behavior_level)
    for i, s in enumerate(g["train"]):
        ids = [tid[t] for t in TOK.findall(s["inters"] + s["item"])]
        n = len(ids)
        assert batch["input_ids"][i, :n].tolist() == ids, i
        assert int(batch["attention_mask"][i].sum()) == n
        assert (batch["input_ids"][i, n:] == 4).all() and (batch["actions"][i, n:] == 100).all()
        assert batch["session_ids"][i, :n].tolist() == s["session_ids"], i
        assert batch["extended_session_ids"][i, :n].tolist() == s["extended_session_ids"], i
        assert batch["actions"][i, :n].tolist() == s["actions"], i
        lab = batch["labels"][i, :n].tolist()
        assert lab == [(-100 if t in data.behavior_tokens else t) for t in ids]


def test_test_split_histories_and_targets_match_reference(tmp_path):
    g, data, tid = _setup(tmp_path)
    store, targets = ds.eval_store(data, "test")
    assert store.n_users == len(g["test"])
    tb = data.target_behavior
    batch = collate.collate_eval(store, torch.arange(store.n_users), g["max_his_len"], tb, data.behavior_tokens,
                                 data.behavior_level)
    for i, s in enumerate(g["test"]):
        ids = [tid[t] for t in TOK.findall(s["inters"])]
        n = len(ids)
        row = batch["input_ids"][i]
        assert row[-1] == data.behavior_tokens[tb]
        assert row[-1 - n:-1].tolist() == ids, i                       # left padded, target-behaviour token appended
        assert (row[:-1 - n] == 4).all()
        assert batch["session_ids"][i, -1 - n:-1].tolist() == s["session_ids"], i
        assert batch["extended_session_ids"][i, -1 - n:-1].tolist() == s["extended_session_ids"], i
        assert batch["actions"][i, -1 - n:-1].tolist() == s["actions"], i
        # inference appends max + 1 (tasks/test_SMB_decoder.py:105-117)
        assert int(batch["session_ids"][i, -1]) == max(s["session_ids"]) + 1
        assert int(batch["extended_session_ids"][i, -1]) == max(s["extended_session_ids"]) + 1
        want = {}
        for item, b in zip(s["item"], s["behavior"]):
            toks = [tid[t] for t in TOK.findall(item)]
            want.setdefault(data.behavior_names.index(b), []).append(tuple(toks[1:]))
        assert targets[i] == want, i


def test_augment_drops_rates():
    rng = np.random.RandomState(0)
    beh = np.asarray([0] * 40 + [1] * 8 + [2] * 4, dtype=np.int16)
    masks = ds.augment_drops(beh, [0, 1, 2], 4, rng)
    assert len(masks) == 4
    for r, keep in zip((0.25, 0.5, 0.75, 1.0), masks):
        assert int((~keep[:40]).sum()) == int(40 * r) and int((~keep[40:48]).sum()) == int(8 * (r / 2))
        assert keep[48:].all()                                          # the target behaviour is never dropped
