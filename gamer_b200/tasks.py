"""The two tasks the hot path stays drop-in for, driven by `main.py` with the reference's flag surface
(SeqRec/tasks/train_SMB_decoder.py:23-137,139-449 and SeqRec/tasks/test_SMB_decoder.py:40-64,90-540).

Scope (SURVEY.md §8(b) "CLI" row): the flags, the model construction (config mutation of train_SMB_decoder.py:321-360,
`set_hyper`, `resize_token_embeddings`), the optimisation recipe (AdamW, cosine schedule with warm-up ratio, gradient
accumulation with `num_items_in_batch` normalisation, per-epoch validation loss, best-checkpoint saving, patience) and
the evaluation loop (per-behaviour candidate trie, constrained beam search, hit/recall/ndcg, merged result, results
JSON) are the reference's.  The reference's dataset machinery (string prompts + HF tokenizer over Git-LFS JSON files) is
out of scope: inputs here are the ShortVideoAD-shaped synthetic sessions of
`gamer_b200.synthetic`, produced directly as the tensors the reference's collators emit.  HF `Trainer` (needs
`accelerate`) is replaced by `gamer_b200.trainer.NativeTrainer`.
"""
from __future__ import annotations

import json
import math
import os
import time

import torch
import torch.distributed as dist

from . import modeling
from . import synthetic as syn
from .distributed import reduce_metric_sums, shard_range, world_info
from .ranking import metric_sums, pack_targets, topk_hits
from .trie import flat_from_array, prefix_allowed_tokens_fn_by_last_token

BACKBONES = {
    "Qwen3Multi": modeling.Qwen3MultiWithTemperature,
    "Qwen3SessionMulti": modeling.Qwen3SessionMultiWithTemperature,
    "Qwen3SessionMoe": modeling.Qwen3SessionMoeWithTemperature,      # quirk Q11: accepted here
    "Qwen3Moe": modeling.Qwen3MoeWithTemperature,
}


def _init(seed: int):
    """MultiGPUTask.init (tasks/multi_gpu.py:41-64): seed, device, process group from the torchrun environment."""
    if not torch.cuda.is_available():
        raise RuntimeError("gamer_b200 tasks need a CUDA device (the hot path has no CPU fallback)")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(seed)
    return dev


def _info(msg):
    if world_info()[0] == 0:
        print(f"[gamer_b200] {msg}", flush=True)


def _config(base_model: str, max_his_len: int, model_max_length: int):
    """The runtime mutations of train_SMB_decoder.py:321-360 for the synthetic vocabulary (14 stub ids + 4 x 256 codes +
    3 behaviour tokens = 1041, `synthetic.py`)."""
    from transformers.models.qwen3_moe import Qwen3MoeConfig
    cfg = Qwen3MoeConfig.from_pretrained(base_model)
    cfg.vocab_size = syn.VOCAB
    cfg.num_behavior = syn.N_BEHAVIOR
    cfg.behavior_maps = {str(t): i for i, t in enumerate(syn.BEHAVIOR_TOKENS)}
    cfg.use_behavior_token = True
    cfg.num_positions = syn.TOKENS_PER_ITEM
    cfg.num_experts = syn.TOKENS_PER_ITEM + 1
    cfg.n_positions = max_his_len + 1
    cfg.use_user_token = False
    cfg.model_max_length = max(model_max_length, syn.TOKENS_PER_ITEM * (max_his_len + 1))
    return cfg


def _batches(cat, n_rows, batch, max_his_len, seed):
    out = []
    for i, b0 in enumerate(range(0, n_rows, batch)):
        out.append(syn.make_train_batch(cat, min(batch, n_rows - b0), max_his_len=max_his_len, seed=seed + i))
    return out


_PAD_VALUE = {"input_ids": syn.PAD, "attention_mask": 0, "labels": -100, "actions": 100, "session_ids": 0,
              "extended_session_ids": 0}


def _pad_to_common_length(batch, dev):
    """Every rank must replay / capture the same micro-batch shape and issue the same collectives: right-pad this rank's
    ragged batch to the longest row of any rank (one MAX all-reduce of a scalar per step)."""
    L = torch.tensor([batch["input_ids"].shape[1]], device=dev)
    dist.all_reduce(L, op=dist.ReduceOp.MAX)
    L = int(L)
    out = {}
    for k, v in batch.items():
        if v.shape[1] < L:
            v = torch.nn.functional.pad(v, (0, L - v.shape[1]), value=_PAD_VALUE[k])
        out[k] = v
    return out


@torch.no_grad()
def _valid_loss(model, batches, dev):
    """`--valid_loss` / per-epoch evaluation: mean of the per-batch losses (test_SMB_decoder.py:306-322)."""
    model.eval()
    losses = [float(model(**{k: v.to(dev) for k, v in b.items()}).loss) for b in batches]
    model.train()
    return sum(losses) / max(1, len(losses))


def train_SMB_decoder(seed, backbone, base_model, output_dir, data_path, tasks, dataset, index_file, max_his_len, optim,
                      epochs, learning_rate, per_device_batch_size, gradient_accumulation_steps, logging_step,
                      model_max_length, weight_decay, resume_from_checkpoint, warmup_ratio, lr_scheduler_type,
                      save_and_eval_strategy, save_and_eval_steps, patience, fp16, bf16, deepspeed, temperature,
                      wandb_run_name, synthetic_users=4096, synthetic_items=50_000, **_):
    from .trainer import NativeTrainer
    dev = _init(seed)
    rank, world = world_info()
    if backbone not in BACKBONES:
        raise ValueError(f"backbone {backbone!r}: the B200 hot path implements {sorted(BACKBONES)}")
    if optim != "adamw_torch" or lr_scheduler_type != "cosine":
        raise NotImplementedError("the native loop implements the reference recipe: adamw_torch + cosine schedule")
    _info(f"dataset {dataset!r}: the reference's JSON/tokenizer pipeline is out of scope — training on ShortVideoAD-shaped "
          f"synthetic sessions ({synthetic_users} users, {synthetic_items} items, max_his_len={max_his_len}); bf16 kernels")
    cfg = _config(base_model, max_his_len, model_max_length)
    model = BACKBONES[backbone].from_pretrained(resume_from_checkpoint) if resume_from_checkpoint else BACKBONES[backbone](cfg)
    model.set_hyper(temperature)
    model.resize_token_embeddings(cfg.vocab_size)
    model = model.to(dev).train()
    if world > 1:
        for p in model.parameters():
            dist.broadcast(p.data, src=0)
    cat = syn.make_catalogue(synthetic_items, 1234)
    # equal rows per rank (the remainder of an uneven split is dropped, as DistributedSampler(drop_last) would): every
    # rank then runs the same number of optimizer steps with the same batch sizes — and so the same NCCL sequence
    n_rows = synthetic_users // world
    step_rows = per_device_batch_size * gradient_accumulation_steps
    train = _batches(cat, n_rows, step_rows, max_his_len, seed=10_000 * (rank + 1))
    valid = _batches(cat, max(per_device_batch_size, n_rows // 8), per_device_batch_size, max_his_len, seed=777_000 + rank)
    total_steps = epochs * len(train)
    trainer = NativeTrainer(model, lr=learning_rate, weight_decay=weight_decay, max_grad_norm=1.0,
                            warmup_steps=int(math.ceil(warmup_ratio * total_steps)), total_steps=total_steps)
    if rank == 0:
        os.makedirs(output_dir, exist_ok=True)
        cfg.save_pretrained(output_dir)
    model.config.use_cache = False
    best, bad, step = float("inf"), 0, 0
    for epoch in range(epochs):
        t0, seen = time.time(), 0
        for b in train:
            b = {k: v.to(dev, non_blocking=True) for k, v in b.items()}
            if world > 1:
                b = _pad_to_common_length(b, dev)
            loss = trainer.step(b, micro_batch=per_device_batch_size)
            step += 1
            seen += b["input_ids"].shape[0]
            if step % logging_step == 0:
                _info(f"epoch {epoch} step {step}/{total_steps} loss {float(loss):.4f} lr {trainer.current_lr():.2e}")
        torch.cuda.synchronize()
        dt = time.time() - t0
        vl = torch.tensor([_valid_loss(model, valid, dev)], device=dev)
        if world > 1:
            dist.all_reduce(vl)
            vl /= world
        vl = float(vl)
        _info(f"epoch {epoch}: {seen * world / dt:.0f} samples/s, eval_loss {vl:.4f}")
        if vl < best - 1e-6:                                   # load_best_model_at_end + save_total_limit
            best, bad = vl, 0
            if rank == 0:      # parameters are views of one flat buffer: save independent copies
                model.save_pretrained(output_dir, state_dict={k: v.detach().clone() for k, v in model.state_dict().items()})
        else:
            bad += 1
            if bad >= patience:                                 # EarlyStoppingCallback(patience)
                _info(f"early stop after {patience} evaluations without improvement")
                break
    _info(f"best eval_loss {best:.4f}; checkpoint in {output_dir}")
    if world > 1:
        dist.barrier()
    return best


@torch.no_grad()
def test_SMB_decoder(seed, backbone, base_model, output_dir, data_path, tasks, dataset, index_file, max_his_len,
                     ckpt_path, results_file, test_batch_size, num_beams, metrics, test_task, behaviors, valid_loss,
                     synthetic_users=1024, synthetic_items=50_000, **_):
    dev = _init(seed)
    rank, world = world_info()
    if backbone not in BACKBONES:
        raise ValueError(f"backbone {backbone!r}: the B200 hot path implements {sorted(BACKBONES)}")
    model = BACKBONES[backbone].from_pretrained(ckpt_path).to(dev).eval()
    cat = syn.make_catalogue(synthetic_items, 1234)
    metric_list = metrics.split(",")
    names = [f"behavior_{i}" for i in range(syn.N_BEHAVIOR)]
    wanted = names if not behaviors else [b for b in behaviors if b in names]
    if valid_loss:
        batches = _batches(cat, max(test_batch_size, synthetic_users // 8), test_batch_size, max_his_len, seed=777_000 + rank)
        vl = _valid_loss(model, batches, dev)
        _info(f"Validation loss: {vl:.4f}")
        return vl
    results, merged, total_all = [], {m: 0.0 for m in metric_list}, 0
    for beh_name in wanted:
        beh = names.index(beh_name)
        items = cat.item_sequences(beh)
        fn = prefix_allowed_tokens_fn_by_last_token(flat_from_array(items), set(int(t) for t in items[:, -1]) | {syn.PAD})
        lo, hi = shard_range(synthetic_users, rank, world)      # exact sharding: no duplicated users (quirk Q12)
        sums, count = {m: 0.0 for m in metric_list}, 0
        for b0 in range(lo, hi, test_batch_size):
            n = min(test_batch_size, hi - b0)
            batch, targets = syn.make_eval_batch(cat, n, max_his_len=max_his_len, target_behavior=beh, seed=seed * 1000 + b0)
            out = model.generate(**{k: v.to(dev) for k, v in batch.items()}, max_new_tokens=syn.TOKENS_PER_ITEM - 1,
                                 prefix_allowed_tokens_fn=fn, num_beams=num_beams, num_return_sequences=num_beams,
                                 output_scores=True, return_dict_in_generate=True, early_stopping=True)
            # hit matching and metric sums on the device, on the code-id tuples themselves (ranking.py: the tensor form is
            # checked against the string functions of the reference in tests/test_ranking_device_cpu.py)
            S = syn.TOKENS_PER_ITEM - 1
            gen = out.sequences[:, -S:].view(n, num_beams, S)
            tt, cnt = pack_targets(targets, S, device=dev)
            res = metric_sums(topk_hits(gen, tt), cnt, metric_list)
            for m in metric_list:
                sums[m] = sums[m] + res[m]
            count += n
        means, total = reduce_metric_sums({m: float(v) for m, v in sums.items()}, count, device=dev)   # one sync
        means["eval_type"] = f"Behavior {beh_name}"
        results.append(means)
        for m in metric_list:
            merged[m] += means[m] * total
        total_all += total
        _info(f"Finished testing behavior {beh_name} with {total} samples.")
    for m in merged:
        merged[m] /= max(1, total_all)
    merged["eval_type"] = "Merged Behavior"
    results.append(merged)
    if rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(results_file)), exist_ok=True)
        with open(results_file, "w") as f:
            json.dump(results, f, indent=4)
        _info(f"Results saved to {results_file}.")
    if world > 1:
        dist.barrier()
    return results
