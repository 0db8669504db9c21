"""The two tasks the hot path stays drop-in for, driven by `main.py` with the reference's flag surface
(SeqRec/tasks/train_SMB_decoder.py:23-137,139-449 and SeqRec/tasks/test_SMB_decoder.py:40-64,90-540).

Scope (SURVEY.md §8(b) "CLI" row): the flags, the model construction (config mutation of train_SMB_decoder.py:321-360,
`set_hyper`, `resize_token_embeddings`), the optimisation recipe (AdamW, cosine schedule with warm-up ratio, gradient
accumulation with `num_items_in_batch` normalisation, per-epoch validation loss, best-checkpoint saving, patience) and
the evaluation loop (per-behaviour candidate trie, constrained beam search, hit/recall/ndcg, merged result, results
JSON) are the reference's.

Input pipeline: when `<data_path>/<dataset>/` holds the reference's files
(`<dataset>.SMB.{inter,behavior,session}.json`, the index file, `<dataset>.behavior_level.json`) they are loaded by
`gamer_b200.dataset` — token ids, session splits, the `smb_explicit_decoder_<k>` k-fold augmentation and the candidate
trie exactly as the reference builds them (tests/test_dataset_cpu.py) — otherwise a ShortVideoAD-shaped synthetic corpus
of the same form is generated.  Either way the corpus is a pre-tokenised `PackedSessions` store resident on the GPU and
every batch is assembled there by `gamer_b200.collate` (no strings, no tokenizer, ~7 B per interaction of H2D once).
HF `Trainer` (needs `accelerate`) is replaced by `gamer_b200.trainer.NativeTrainer`.
"""
from __future__ import annotations

import json
import math
import os
import time

import torch
import torch.distributed as dist

from . import collate
from . import dataset as ds
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


def _corpus(data_path, dataset, index_file, max_his_len, synthetic_users, synthetic_items, seed) -> ds.SMBData:
    root = os.path.join(data_path or "", dataset or "")
    if data_path and dataset and os.path.exists(os.path.join(root, dataset + ".SMB.inter.json")):
        data = ds.load_smb_files(data_path, dataset, index_file or ".index.json")
        _info(f"dataset {dataset!r}: {len(data.users)} users, {data.catalogue.shape[0]} items, vocabulary {data.vocab_size} "
              f"from {root}")
        return data
    _info(f"dataset {dataset!r}: no data files under {root!r} — ShortVideoAD-shaped synthetic corpus "
          f"({synthetic_users} users, {synthetic_items} items, max_his_len={max_his_len})")
    return ds.synthetic_corpus(synthetic_users, synthetic_items, max_his_len + 8, seed=seed)


def _augment_of(tasks: str):
    """loading_SMB.py:39-54: `smb_explicit_decoder` -> no augmentation, `smb_explicit_decoder_<k>` -> augment = k."""
    t = (tasks or "smb_explicit_decoder").split(",")[0].lower()
    if not t.startswith("smb_explicit_decoder"):
        raise NotImplementedError(f"task {tasks!r}: the B200 hot path implements smb_explicit_decoder[_<k>]")
    return None if t == "smb_explicit_decoder" else int(t.split("_")[3])


def _config(base_model: str, max_his_len: int, model_max_length: int, data: ds.SMBData):
    """The runtime mutations of train_SMB_decoder.py:321-360 for the corpus' vocabulary (14 base ids + the sorted added
    tokens: 4 x 256 codes + the behaviour tokens = 1041 for three behaviours)."""
    from transformers.models.qwen3_moe import Qwen3MoeConfig
    cfg = Qwen3MoeConfig.from_pretrained(base_model)
    tpi = data.catalogue.shape[1] + 1
    cfg.vocab_size = data.vocab_size
    cfg.num_behavior = len(data.behavior_tokens)
    cfg.behavior_maps = {str(t): i for i, t in enumerate(data.behavior_tokens)}
    cfg.use_behavior_token = True
    cfg.num_positions = tpi
    cfg.num_experts = tpi + 1
    cfg.n_positions = max_his_len + 1
    cfg.use_user_token = False
    cfg.model_max_length = max(model_max_length, tpi * (max_his_len + 1))
    return cfg


def _step_users(n_rows: int, rows_per_step: int, epoch_seed: int):
    """Shuffled user order of one epoch cut into optimizer steps of equal size (the ragged tail is dropped, as
    DistributedSampler(drop_last) would): every rank runs the same number of steps with the same batch sizes."""
    g = torch.Generator().manual_seed(epoch_seed)
    perm = torch.randperm(n_rows, generator=g)
    n_steps = n_rows // rows_per_step
    return [perm[i * rows_per_step:(i + 1) * rows_per_step] for i in range(n_steps)]


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
def _valid_loss(model, store, rank, world, batch_size, max_his_len, data):
    """`--valid_loss` / per-epoch evaluation: mean of the per-batch losses (test_SMB_decoder.py:306-322) over this rank's
    exact shard of the validation samples, batches collated on the device."""
    model.eval()
    lo, hi = shard_range(store.n_users, rank, world)
    losses = []
    for b0 in range(lo, hi, batch_size):
        users = torch.arange(b0, min(hi, b0 + batch_size), device=store.offsets.device)
        b = collate.collate_train(store, users, max_his_len, data.behavior_tokens, data.behavior_level)
        losses.append(float(model(**b).loss))
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
    data = _corpus(data_path, dataset, index_file, max_his_len, synthetic_users, synthetic_items, seed)
    augment = _augment_of(tasks)
    cfg = _config(base_model, max_his_len, model_max_length, data)
    model = BACKBONES[backbone].from_pretrained(resume_from_checkpoint) if resume_from_checkpoint else BACKBONES[backbone](cfg)
    model.set_hyper(temperature)
    model.resize_token_embeddings(cfg.vocab_size)
    model = model.to(dev).train()
    if world > 1:
        for p in model.parameters():
            dist.broadcast(p.data, src=0)
    # the whole corpus lives on the GPU as flat int arrays (pinned host copy -> one H2D); batches are collated there
    train = ds.train_store(data, augment).pin_memory().to(dev, non_blocking=True)
    valid = ds.valid_store(data).pin_memory().to(dev, non_blocking=True)
    _info(f"train sequences {train.n_users} (augment={augment}), validation samples {valid.n_users}")
    step_rows = per_device_batch_size * gradient_accumulation_steps
    steps_per_epoch = train.n_users // (step_rows * world)
    if steps_per_epoch == 0:
        raise ValueError(f"{train.n_users} training sequences do not fill one optimizer step of {step_rows * world} rows")
    total_steps = epochs * steps_per_epoch
    trainer = NativeTrainer(model, lr=learning_rate, weight_decay=weight_decay, max_grad_norm=1.0,
                            warmup_steps=int(math.ceil(warmup_ratio * total_steps)), total_steps=total_steps)
    if rank == 0:
        os.makedirs(output_dir, exist_ok=True)
        cfg.save_pretrained(output_dir)
    model.config.use_cache = False
    best, bad, step = float("inf"), 0, 0
    for epoch in range(epochs):
        t0, seen = time.time(), 0
        # one global shuffle per epoch (same on every rank), rank r takes rows r::world of every step
        for users in _step_users(train.n_users, step_rows * world, seed * 1000 + epoch):
            mine = users[rank::world].to(dev)
            b = collate.collate_train(train, mine, max_his_len, data.behavior_tokens, data.behavior_level)
            if world > 1:
                b = _pad_to_common_length(b, dev)
            loss = trainer.step(b, micro_batch=per_device_batch_size)
            step += 1
            seen += b["input_ids"].shape[0]
            if step % logging_step == 0:
                _info(f"epoch {epoch} step {step}/{total_steps} loss {float(loss):.4f} lr {trainer.current_lr():.2e}")
                model.check_token_ids()                      # the host is synchronised here anyway
        torch.cuda.synchronize()
        dt = time.time() - t0
        vl = torch.tensor([_valid_loss(model, valid, rank, world, per_device_batch_size, max_his_len, data)], device=dev)
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
    trainer.close()
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
    data = _corpus(data_path, dataset, index_file, max_his_len, synthetic_users, synthetic_items, seed)
    if model.config.vocab_size != data.vocab_size:
        raise ValueError(f"checkpoint vocabulary {model.config.vocab_size} != data vocabulary {data.vocab_size}")
    metric_list = metrics.split(",")
    names = list(data.behavior_names)
    wanted = names if not behaviors else [b for b in behaviors if b in names]
    if valid_loss:
        valid = ds.valid_store(data).pin_memory().to(dev, non_blocking=True)
        vl = _valid_loss(model, valid, rank, world, test_batch_size, max_his_len, data)
        _info(f"Validation loss: {vl:.4f}")
        return vl
    mode = "valid" if (test_task or "").lower().startswith("smb_valid") else "test"
    store, targets_all = ds.eval_store(data, mode)
    store = store.pin_memory().to(dev, non_blocking=True)
    S = data.catalogue.shape[1]
    results, merged, total_all = [], {m: 0.0 for m in metric_list}, 0
    for beh_name in wanted:
        beh = names.index(beh_name)
        # the candidate trie of this behaviour: behaviour token + the code tokens of every catalogue item
        # (test_SMB_decoder.py:467-501); users without a target of this behaviour are skipped (:123-135)
        items = data.item_sequences(beh)
        fn = prefix_allowed_tokens_fn_by_last_token(flat_from_array(items), set(int(t) for t in items[:, -1]) | {syn.PAD})
        users_b = [u for u in range(store.n_users) if beh in targets_all[u]]
        lo, hi = shard_range(len(users_b), rank, world)         # exact sharding: no duplicated users (quirk Q12)
        sums, count = {m: 0.0 for m in metric_list}, 0
        for b0 in range(lo, hi, test_batch_size):
            sel = users_b[b0:min(hi, b0 + test_batch_size)]
            n = len(sel)
            batch = collate.collate_eval(store, torch.tensor(sel, device=dev), max_his_len, beh, data.behavior_tokens,
                                         data.behavior_level)
            out = model.generate(**batch, max_new_tokens=S, prefix_allowed_tokens_fn=fn, num_beams=num_beams,
                                 num_return_sequences=num_beams, output_scores=True, return_dict_in_generate=True,
                                 early_stopping=True)
            # hit matching and metric sums on the device, on the code-id tuples themselves (ranking.py: the tensor form is
            # checked against the string functions of the reference in tests/test_ranking_device_cpu.py)
            gen = out.generated.view(n, num_beams, S)
            tt, cnt = pack_targets([targets_all[u][beh] for u in sel], S, device=dev)
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
