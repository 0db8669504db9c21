"""CPU, world_size 2, gloo: the host-side data-parallel logic — bucketed all-reduce of the flat gradient buffer,
exact user sharding and the single-tensor metric reduction.  Gradients come from the oracle (test infrastructure)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import load_golden, spec_from_golden, weights_from_golden


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gamer_b200 import engine as E
        from gamer_b200.distributed import BucketReducer, reduce_metric_sums, shard_range
        from oracle import oracle_model as om
        from tests.test_model_gpu import build_model
        torch.set_num_threads(2)
        g = load_golden("train_qwen3multi.pt")
        spec = spec_from_golden(g, g["temperature"])
        arch = build_model(g).arch
        B = g["batch"]["input_ids"].shape[0]
        s, e = shard_range(B, rank, world)
        sub = {k: v[s:e] for k, v in g["batch"].items()}
        # per-rank token-mean loss (HF 4.51 default), gradients averaged over ranks — what DDP does for the reference
        W = weights_from_golden(g, requires_grad=True)
        om.forward(spec, W, **sub)["loss"].backward()
        flat = torch.zeros(E.flat_size(arch))
        G = E.flat_views(arch, flat)
        named = E.unfuse_grads(arch, G)
        for k, v in named.items():
            v.copy_(W[k].grad if W[k].grad is not None else torch.zeros_like(W[k]))
        red = BucketReducer(flat, E.layer_ranges(arch))
        local = flat.clone()
        red.launch_all_reverse()
        red.wait_all()
        bucketed = flat.clone()
        flat.copy_(local)                     # the single-collective path (CUDA-graph mode) reduces to the same sums
        red.launch_flat()
        red.wait_all()
        assert torch.allclose(flat, bucketed, rtol=1e-6, atol=1e-8)
        flat /= world
        if rank == 0:
            # single-process equivalent: mean over ranks of the per-shard mean losses
            W2 = weights_from_golden(g, requires_grad=True)
            total = 0
            for r in range(world):
                s2, e2 = shard_range(B, r, world)
                total = total + om.forward(spec, W2, **{k: v[s2:e2] for k, v in g["batch"].items()})["loss"] / world
            total.backward()
            worst = 0.0
            for k, v in named.items():
                ref = W2[k].grad if W2[k].grad is not None else torch.zeros_like(W2[k])
                worst = max(worst, (v - ref).abs().max().item() / (ref.abs().max().item() + 1e-12))
            ret["worst"] = worst
        means, n = reduce_metric_sums({"recall@5": 1.5 + rank, "ndcg@5": 0.5 * (rank + 1)}, 3 + rank)
        ret[f"metrics{rank}"] = (means, n)
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_and_metric_reduction_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret["worst"] < 1e-5, ret["worst"]
    for r in range(world):
        means, n = ret[f"metrics{r}"]
        assert n == 7
        assert abs(means["recall@5"] - (1.5 + 2.5) / 7) < 1e-12 and abs(means["ndcg@5"] - 1.5 / 7) < 1e-12


def test_shard_range_is_exact_partition():
    from gamer_b200.distributed import shard_range
    for n in (0, 1, 7, 256, 1001):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_layer_ranges_cover_flat_buffer():
    from gamer_b200 import engine as E
    from tests.test_model_gpu import build_model
    arch = build_model(load_golden("train_qwen3multi.pt")).arch
    ranges = E.layer_ranges(arch)
    assert ranges[0][1] == 0 and ranges[-1][2] == E.flat_size(arch)
    assert all(ranges[i][2] == ranges[i + 1][1] for i in range(len(ranges) - 1))
    assert [r[0] for r in ranges] == ["model.embed_tokens.weight"] + [f"L{i}" for i in range(arch.n_layers)] + ["model.norm.weight"]
