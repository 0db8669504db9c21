"""Native training step for the drop-in models: forward + backward + gradient all-reduce + AdamW on flat buffers.

The reference trains through HF `Trainer` (SeqRec/tasks/train_SMB_decoder.py:396-447): torch DDP bucketed all-reduce,
`clip_grad_norm_(1.0)`, `adamw_torch`, cosine schedule with warm-up, gradient accumulation with the loss normalised by
`num_items_in_batch`.  `accelerate` is not installed here, so HF Trainer cannot run; this module is the
loop underneath it, restated B200-first:

  * all master weights live in ONE flat fp32 buffer laid out by `engine.fused_blocks` (q|k|v and gate|up stacked the way
    the GEMMs read them); every HF parameter is a view of it, so state dicts / checkpoints are unchanged;
  * gradients are accumulated by the wgrad kernels straight into a flat fp32 buffer of the same layout; one layer = one
    contiguous range = one NCCL all-reduce bucket, launched (async, NCCL's stream) as soon as that layer's backward
    kernels are enqueued, so the reduction of layer l overlaps the backward of layers l-1..0;
  * one fused AdamW kernel updates p/m/v and emits the bf16 operand copy for the next step's GEMMs;
  * the forward + backward of a micro-batch (~1500 kernel launches, ~11 us of Python/ctypes each) is captured ONCE per
    batch shape into a CUDA graph and replayed: inputs are copied into static buffers, every operand (flat weights,
    transposed copies, gradient buffer) keeps its address, and the dropout masks change per replay through a device-side
    offset word.  The graph of a step's last micro-batch also holds the per-layer bucket all-reduces (NCCL's stream as a
    parallel branch), so the overlap with the backward survives the capture.
"""
from __future__ import annotations

import math
import re

import torch
import torch.distributed as dist

from . import engine as E
from ._cabi import call, ptr
from .distributed import BucketReducer


def _stream():
    return torch.cuda.current_stream().cuda_stream


# Name patterns exempt from weight decay.  "hf-4.51": the release the reference pins (bias / layernorm / rmsnorm in the
# qualified name: input_layernorm and post_attention_layernorm are exempt, q_norm / k_norm / model.norm are decayed);
# "hf-5": the list of the transformers 5.x installed here (every *norm weight exempt).
DECAY_EXEMPT = {
    "hf-4.51": (r"bias", r"layernorm", r"rmsnorm"),
    "hf-5": (r"bias", r"layernorm", r"rmsnorm", r"(?:^|\.)norm(?:$|\.)", r"_norm(?:$|\.)"),
}
HP_SLOTS = 8          # pinned staging slots for the per-step scalars (see optimizer_step)


class NativeTrainer:
    BATCH_KEYS = ("input_ids", "attention_mask", "labels", "actions", "session_ids", "extended_session_ids")

    def __init__(self, model, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, max_grad_norm=1.0,
                 warmup_steps=0, total_steps=None, process_group=None, min_lr_ratio=0.0, use_cuda_graphs=True,
                 max_graphs=4, decay_rule="hf-4.51"):
        self.model = model
        self.arch = model.arch
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("NativeTrainer needs the model on a CUDA device (no CPU path)")
        self.dev = dev
        a = self.arch
        n = E.flat_size(a)
        self.flat_p = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_bf16 = torch.zeros(n, dtype=torch.bfloat16, device=dev)
        # re-point every parameter at its slice of the flat buffer (checkpoint keys and shapes unchanged)
        views = E.unfuse_grads(a, E.flat_views(a, self.flat_p))
        params = dict(model.named_parameters())
        with torch.no_grad():
            for k, p in params.items():
                views[k].copy_(p.data)
                p.data = views[k]
        # HF Trainer.get_decay_parameter_names: weight decay skips parameters whose qualified name matches a pattern
        # list that changed between releases (DECAY_EXEMPT); the reference pins transformers 4.51
        mask = torch.zeros(n, dtype=torch.uint8, device=dev)
        mviews = E.unfuse_grads(a, E.flat_views(a, mask))
        exempt = [re.compile(pat) for pat in DECAY_EXEMPT[decay_rule]]
        for k in params:
            if not any(pat.search(k.lower()) for pat in exempt):
                mviews[k].fill_(1)
        self.decay_mask = mask
        self.G = E.grad_buffers(a, dev, self.flat_g)
        self.ranges = E.layer_ranges(a)
        self.lr, self.betas, self.eps, self.wd, self.max_grad_norm = lr, betas, eps, weight_decay, max_grad_norm
        self.warmup_steps, self.total_steps, self.min_lr_ratio = warmup_steps, total_steps, min_lr_ratio
        self.step_idx = 0
        self.pg = process_group
        self.reducer = BucketReducer(self.flat_g, self.ranges, process_group)
        self.world = self.reducer.world
        self.hp_host = torch.zeros(HP_SLOTS, 4, dtype=torch.float32).pin_memory()
        self.hp_events = [None] * HP_SLOTS
        self.hp = torch.zeros(4, dtype=torch.float32, device=dev)
        self.gnorm_sq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.lut = E.behaviour_lut(a, dev)
        call("gamer_cast_f32_bf16", ptr(self.flat_p), ptr(self.flat_bf16), n, _stream())
        self.pack = E.FlatPack(a, self.flat_bf16, self.flat_p)
        self.use_cuda_graphs = use_cuda_graphs
        self.max_graphs = max_graphs
        self.graphs = {}                   # shape key -> captured micro-batch (fwd + bwd)
        self.seen = {}                     # shape key -> eager runs so far (capture on the second occurrence)
        self.opt_graph, self.opt_eager_runs = None, 0
        self.drop_off = torch.zeros(1, dtype=torch.int32, device=dev)   # device-side dropout offset, +1 per micro-batch
        self.micro_batches = 0

    # ------------------------------------------------------------------------------------------------------------
    def current_lr(self):
        """HF `cosine` schedule with linear warm-up (get_cosine_schedule_with_warmup)."""
        s = self.step_idx
        if self.warmup_steps and s < self.warmup_steps:
            return self.lr * s / max(1, self.warmup_steps)
        if not self.total_steps:
            return self.lr
        prog = (s - self.warmup_steps) / max(1, self.total_steps - self.warmup_steps)
        return self.lr * max(self.min_lr_ratio, 0.5 * (1.0 + math.cos(math.pi * min(1.0, prog))))

    def _bucket_hook(self, layer_idx):
        """Called right after layer `layer_idx`'s backward kernels were enqueued (n_layers = final norm, -1 = embedding):
        start that bucket's all-reduce; NCCL's stream waits on everything enqueued so far and runs beside the rest of
        the backward."""
        a = self.arch
        self.reducer.launch("model.norm.weight" if layer_idx == a.n_layers else
                            ("model.embed_tokens.weight" if layer_idx < 0 else f"L{layer_idx}"))

    def _drop_ctx(self):
        cfg = self.model._drop_config()
        if cfg is None:
            return None
        seed, p_h, p_a = cfg
        return E.DropCtx(seed, 0, p_h, p_a, offset_dev=self.drop_off)

    def _run_micro(self, batch, inv_norm, hook):
        a = self.arch
        ids = batch["input_ids"].contiguous()
        meta = E.make_meta(a, ids, batch.get("attention_mask"), batch.get("actions"), batch.get("session_ids"),
                           batch.get("extended_session_ids"))
        shifted = E.shift_labels(batch["labels"])
        # forward and backward always run together here: the loss comes out of the backward's fused CE pass, and the
        # forward lm_head GEMM + CE (a second pass over the [M, V] logits) is skipped
        _, st = E.loss_forward(a, self.pack, meta, self.lut, ids, shifted, inv_norm, float(self.model.temperature),
                               drop=self._drop_ctx(), defer_loss=True)
        one = torch.ones((), dtype=torch.float32, device=self.dev)
        return E.loss_backward(a, self.pack, st, one, self.G, on_layer_done=hook)

    def _graph_key(self, batch, reduce):
        cfg = self.model._drop_config()
        return (tuple((k, tuple(batch[k].shape), batch[k].dtype) for k in self.BATCH_KEYS if batch.get(k) is not None),
                cfg[1:] if cfg else None, float(self.model.temperature), bool(reduce))

    def _capture(self, key, batch, inv_norm, reduce):
        """Capture forward + backward of one micro-batch shape.  The eager run that precedes every capture (see
        forward_backward) has already initialised the lazily configured kernels.  With `reduce` (the last micro-batch of a
        step on more than one rank) the per-layer bucket all-reduces are captured too: each is launched from the backward
        hook on NCCL's stream — a parallel branch of the graph that runs beside the remaining backward kernels, as DDP's
        bucketed all-reduce does for the reference (tasks/train_SMB_decoder.py:396-428) — and joined before the capture
        ends."""
        static = {k: torch.empty_like(batch[k]) for k in self.BATCH_KEYS if batch.get(k) is not None}
        for k, v in static.items():
            v.copy_(batch[k])
        s_inv = inv_norm.clone()
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        torch.cuda.empty_cache()     # hand the eager run's cached activations back: the graph gets its own pool
        with torch.cuda.graph(g):
            loss = self._run_micro(static, s_inv, self._bucket_hook if reduce else None)
            if reduce:
                self.reducer.wait_all()
        if len(self.graphs) >= self.max_graphs:
            self.graphs.pop(next(iter(self.graphs)))
        self.graphs[key] = (g, static, s_inv, loss)
        return self.graphs[key]

    def forward_backward(self, batch, inv_norm, last_micro=True):
        """Forward + backward of one micro-batch into the flat gradient buffer; returns its loss (device scalar)."""
        self.drop_off.add_(1)
        self.micro_batches += 1
        if not self.use_cuda_graphs:
            return self._run_micro(batch, inv_norm, self._bucket_hook if last_micro else None)
        reduce = last_micro and self.world > 1
        key = self._graph_key(batch, reduce)
        entry = self.graphs.get(key)
        if entry is None:
            self.seen[key] = self.seen.get(key, 0) + 1
            if self.seen[key] < 2:                 # first occurrence of a shape: eager (also the kernels' warm-up)
                # same collectives as the replays (the per-layer buckets, in backward order), so ranks whose graph caches
                # differ — ragged batches give every rank its own shape history — still issue identical NCCL sequences
                return self._run_micro(batch, inv_norm, self._bucket_hook if reduce else None)
            entry = self._capture(key, batch, inv_norm, reduce)      # capture records, it does not execute
        g, static, s_inv, loss = entry
        for k, v in static.items():
            v.copy_(batch[k], non_blocking=True)
        s_inv.copy_(inv_norm)
        g.replay()                                 # (with `reduce`: the bucket all-reduces are part of the graph)
        return loss.clone()

    def close(self):
        """Drop the captured graphs.  Graphs that hold NCCL kernels keep the communicator busy: call this (or delete the
        trainer) before `dist.destroy_process_group()`, which otherwise waits on them forever."""
        self.graphs.clear()
        self.opt_graph = None
        import gc
        gc.collect()
        torch.cuda.synchronize()

    def step(self, batch, micro_batch=None):
        """One optimizer step over `batch` (device tensors, [B, L]); `micro_batch` splits it for gradient accumulation
        with the HF `num_items_in_batch` normalisation (sum of CE over the window / number of label tokens)."""
        B = batch["input_ids"].shape[0]
        mb = B if not micro_batch else min(micro_batch, B)
        shifted_all = E.shift_labels(batch["labels"])
        inv_norm = (1.0 / (shifted_all != -100).sum().clamp(min=1).float()).reshape(1)
        self.flat_g.zero_()
        total = torch.zeros((), dtype=torch.float32, device=self.dev)
        starts = list(range(0, B, mb))
        for i, b0 in enumerate(starts):
            sub = {k: v[b0:b0 + mb] for k, v in batch.items() if torch.is_tensor(v)}
            total = total + self.forward_backward(sub, inv_norm, last_micro=(i == len(starts) - 1))
        self.reducer.wait_all()
        self.optimizer_step()
        return total

    def optimizer_step(self):
        """clip + AdamW + bf16 operand refresh.  Step-dependent scalars (lr, bias corrections) travel through a small
        device tensor, so with CUDA graphs the ~70 launches of this method are captured once and replayed."""
        # HF LambdaLR: update k (1-based) runs with lambda(k-1); Adam's bias corrections use t = k.
        lr = self.current_lr()
        self.step_idx += 1
        t = self.step_idx
        # The host runs ahead of the device (graph replays are asynchronous), so every step stages its scalars in its own
        # pinned slot; a slot is rewritten only after the H2D copy that last read it has completed.
        slot = t % HP_SLOTS
        if self.hp_events[slot] is not None:
            self.hp_events[slot].synchronize()
        row = self.hp_host[slot]
        row[0] = lr
        row[1] = 1.0 - self.betas[0] ** t
        row[2] = 1.0 - self.betas[1] ** t
        self.hp.copy_(row, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.hp_events[slot] = ev
        if not self.use_cuda_graphs:
            self._optimizer_kernels()
        elif self.opt_graph is None:
            if self.opt_eager_runs < 1:            # first call eager: configures the kernels, and is the warm-up
                self.opt_eager_runs += 1
                self._optimizer_kernels()
            else:
                g = torch.cuda.CUDAGraph()
                torch.cuda.synchronize()
                with torch.cuda.graph(g):
                    self._optimizer_kernels()
                self.opt_graph = g
                g.replay()
        else:
            self.opt_graph.replay()
        self.model._pack = None                    # the kernel wrote the weights behind autograd's back

    def _optimizer_kernels(self):
        n = self.flat_p.numel()
        gn = None
        if self.max_grad_norm and self.max_grad_norm > 0:
            self.gnorm_sq.zero_()
            call("gamer_sumsq_accumulate", ptr(self.flat_g), n, ptr(self.gnorm_sq), _stream())
            gn = self.gnorm_sq
        call("gamer_adamw_step", ptr(self.flat_p), ptr(self.flat_g), ptr(self.flat_m), ptr(self.flat_v),
             ptr(self.decay_mask), ptr(self.flat_bf16), n, ptr(self.hp), self.betas[0], self.betas[1], self.eps,
             self.wd, ptr(gn), float(self.max_grad_norm or 0.0), 1.0 / self.world, _stream())
        self.pack.refresh()                        # in place: operands keep their addresses (captured graphs stay valid)
