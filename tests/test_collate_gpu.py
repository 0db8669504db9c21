"""GPU: `gamer_collate_sessions` (one launch, csrc/collate.cu) against the tensor-op specification of
gamer_b200/collate.py run on the host — bit for bit, train and eval modes, histories shorter and longer than the window,
empty users, user subsets, explicit and derived widths."""
import numpy as np
import pytest
import torch

from gamer_b200 import collate as C
from gamer_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
BT, BL = syn.BEHAVIOR_TOKENS, syn.BEHAVIOR_LEVEL


def _store(n_users, max_len, seed, empty=()):
    rng = np.random.default_rng(seed)
    hist = []
    for u in range(n_users):
        n = 0 if u in empty else int(rng.integers(1, max_len + 1))
        toks = rng.integers(14, 1038, size=(n, 4))
        beh = rng.integers(0, len(BT), size=n)
        sess = np.cumsum(rng.random(n) < 0.3).astype(np.int64) + int(rng.integers(0, 50))
        hist.append((toks, beh, sess))
    return C.PackedSessions.from_histories(hist)


@pytest.mark.parametrize("mhl,max_len,width_mode", [(20, 40, "fixed"), (7, 30, "auto"), (100, 90, "fixed"), (33, 64, "auto")])
def test_device_collate_equals_host_specification(mhl, max_len, width_mode):
    host = _store(57, max_len, seed=mhl, empty=(3, 56))
    dev = host.to("cuda")
    users = torch.tensor([5, 0, 3, 56, 11, 11, 42, 1, 2, 55, 30])
    for train in (True, False):
        n_max = mhl + 1 if train else mhl
        width = n_max if width_mode == "fixed" else None
        if train:
            want = C.collate_train(host, users, mhl, BT, BL, pad=syn.PAD, width=width)
            got = C.collate_train(dev, users.cuda(), mhl, BT, BL, pad=syn.PAD, width=width)
        else:
            want = C.collate_eval(host, users, mhl, 1, BT, BL, pad=syn.PAD, width=width)
            got = C.collate_eval(dev, users.cuda(), mhl, 1, BT, BL, pad=syn.PAD, width=width)
        assert set(got) == set(want)
        for k in want:
            assert got[k].is_cuda and got[k].dtype == torch.int64 and got[k].shape == want[k].shape, k
            assert torch.equal(got[k].cpu(), want[k]), (train, k)


def test_device_collate_headline_shape_matches_synthetic_batch():
    """1024 full-length rows at max_his_len 100 (the bench's e2e input): equal to the synthetic batch the resident leg uses."""
    cat = syn.make_catalogue(20_000, 5)
    want = syn.make_train_batch(cat, 256, max_his_len=100, seed=9, full_length=True)
    store = C.PackedSessions.from_histories(syn.train_histories(cat, 256, max_his_len=100, seed=9, full_length=True)).to("cuda")
    got = C.collate_train(store, torch.arange(256, device="cuda"), 100, BT, BL, pad=syn.PAD, width=101)
    for k in want:
        assert torch.equal(got[k].cpu(), want[k]), k
