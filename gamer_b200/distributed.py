"""Data-parallel plumbing (one process per GPU, torch.distributed): the only collectives on the hot path
(SURVEY.md §2.2 / §8(e)).

  * training: the flat fused gradient buffer is all-reduced in per-layer buckets (contiguous ranges, reverse layer
    order), each launched asynchronously as soon as that layer's backward kernels are enqueued — the equivalent of torch
    DDP's bucketed all-reduce that HF Trainer sets up for the reference (tasks/train_SMB_decoder.py:396-428), without
    the per-parameter hooks and the bucket copy (gradients already live in the bucket);
  * evaluation: users are sharded exactly (no DistributedSampler padding duplicates, quirk Q12) and the metric sums are
    reduced ONCE with a single tensor all-reduce instead of four pickled all_gather_object calls per batch
    (tasks/test_SMB_decoder.py:232-241).

Works with any backend: NCCL over NVLink on the GPU box, gloo on CPU for the tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


class BucketReducer:
    """All-reduce (sum) of named contiguous ranges of one flat gradient buffer."""

    def __init__(self, flat_g: torch.Tensor, ranges, group=None):
        self.flat_g, self.group = flat_g, group
        self.ranges = {name: (s, e) for name, s, e in ranges}
        self.order = [name for name, _, _ in ranges]
        self.pending = []
        self.rank, self.world = world_info(group)

    def launch(self, name: str):
        if self.world == 1:
            return
        s, e = self.ranges[name]
        self.pending.append(dist.all_reduce(self.flat_g[s:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def launch_flat(self):
        """The whole buffer in one collective (used when the backward ran as a CUDA graph: nothing left to overlap
        with, so per-bucket launches would only add latency)."""
        if self.world == 1:
            return
        self.pending.append(dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def launch_all_reverse(self):
        for name in reversed(self.order):
            self.launch(name)

    def wait_all(self):
        for h in self.pending:
            h.wait()
        self.pending.clear()


def shard_range(n: int, rank: int, world: int):
    """Exact contiguous sharding of n users: sizes differ by at most one, nothing is duplicated."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def reduce_metric_sums(sums: dict, count: int, device=None, group=None):
    """{metric: local sum}, local user count -> ({metric: global mean}, global count) with one all-reduce."""
    names = sorted(sums)
    t = torch.tensor([float(sums[k]) for k in names] + [float(count)], dtype=torch.float64, device=device)
    _, world = world_info(group)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    total = t[-1].item()
    return {k: (t[i].item() / total if total else 0.0) for i, k in enumerate(names)}, int(total)
