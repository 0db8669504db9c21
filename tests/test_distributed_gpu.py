"""GPU, world_size 2, NCCL (skipped on a single-GPU box): the reduced flat gradient of two ranks that each ran the CUDA
forward + backward on half of a batch equals the single-rank gradient of the whole batch (SURVEY.md §4 tier 4) — eagerly
(per-layer buckets launched from the backward hooks) and through the captured micro-batch graph (one flat all-reduce)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import load_golden

pytestmark = pytest.mark.gpu


def _grad_of(trainer, batch, E):
    shifted = E.shift_labels(batch["labels"])
    inv_norm = (1.0 / (shifted != -100).sum().clamp(min=1).float()).reshape(1)
    trainer.flat_g.zero_()
    loss = trainer.forward_backward(batch, inv_norm, last_micro=True)
    trainer.reducer.wait_all()
    torch.cuda.synchronize()
    return trainer.flat_g.clone() / trainer.world, float(loss)


def _worker(rank, world, port, ret):
    import faulthandler
    faulthandler.dump_traceback_later(90, exit=True)      # a collective that never completes must not hang the box
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from gamer_b200 import engine as E
        from gamer_b200 import synthetic as syn
        from gamer_b200.trainer import NativeTrainer
        from tests.test_model_gpu import build_model
        g = load_golden("train_qwen3multi.pt")
        cat = syn.make_catalogue(2000, 1)
        # full-length rows: every rank normalises by the same label count, so mean-of-rank-means = global mean
        full = syn.make_train_batch(cat, 8, max_his_len=12, seed=5, full_length=True)
        dev = torch.device("cuda", rank)
        res, trainers = {}, []
        for graphs in (False, True):
            m = build_model(g).to(dev).train()
            m.config.dropout_rate = 0.0
            m.config.attention_dropout = 0.0
            tr = NativeTrainer(m, use_cuda_graphs=graphs)
            trainers.append(tr)
            half = {k: v[rank * 4:(rank + 1) * 4].to(dev) for k, v in full.items()}
            reps = 3 if graphs else 1            # graphs: eager, capture + replay, replay
            for _ in range(reps):
                mine, loss = _grad_of(tr, half, E)
            res[graphs] = mine
        assert torch.allclose(res[False], res[True], rtol=1e-5, atol=1e-7), "bucketed (eager) and flat (graph) reductions differ"
        if rank == 0:
            # single-rank gradient of the whole batch on a fresh trainer outside the process group's world
            m = build_model(g).to(dev).train()
            m.config.dropout_rate = 0.0
            m.config.attention_dropout = 0.0
            tr = NativeTrainer(m, use_cuda_graphs=False)
            tr.reducer.world, tr.world = 1, 1
            whole, _ = _grad_of(tr, {k: v.to(dev) for k, v in full.items()}, E)
            rel = ((res[True] - whole).norm() / whole.norm()).item()
            ret["rel"] = rel
            ret["max"] = ((res[True] - whole).abs().max() / whole.abs().max()).item()
        for t in trainers:          # graphs holding NCCL kernels must go before the process group does
            t.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run under gpurun --gpus 2)")
def test_two_rank_nccl_gradient_equals_single_rank():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29611, ret), nprocs=2, join=True)
    # same kernels, same rows; only the fp32 summation order across the two halves differs
    assert ret["rel"] <= 1e-3, dict(ret)
    assert ret["max"] <= 1e-3, dict(ret)
