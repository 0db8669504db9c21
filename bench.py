#!/usr/bin/env python
"""bench.py — headline benchmark of the GAMER decoder hot path on B200 (contract: see the task brief / DESIGN.md §5).

    python bench.py --gpus 1 --steps 5 --warmup 3                 # our arm (CUDA path through the C-ABI)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1  # reference arm: the oracle port on the host cores

Metric (BASELINE.json): train samples/s of Qwen3Multi smb_explicit_decoder on ShortVideoAD-shaped synthetic sessions,
max_his_len=100 (L = 505 tokens, every row full length), global batch 1024, bf16.  A "step" = one optimizer step over
the global batch: forward + backward (+ gradient all-reduce) + clip + AdamW.  The global batch is fixed as N grows
(strong scaling, as the reference's launcher divides batch_size by the GPU count: scripts/train_SMB_decoder.sh:17).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gamer", choices=["gamer", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "mb_decoder", "long_history"],
                    help="train = the headline (BASELINE.json configs[1] + the configs[2] eval leg); mb_decoder = configs[3] "
                         "(train_MB_decoder: Qwen3Moe, 4 behaviour types, max_his_len 200); long_history = configs[4] "
                         "(Qwen3Multi max_his_len 500, batch sweep 256-4096)")
    ap.add_argument("--global-batch", type=int, default=1024)
    ap.add_argument("--micro-batch", type=int, default=1024,
                    help="rows per forward+backward pass (gradient accumulation over the per-GPU batch); 1024 rows keep "
                         "~72 GB of activations of the 180 GB and run the step as one pass (2.5 %% faster than 2 x 512)")
    ap.add_argument("--workload-micro-batch", type=int, default=None,
                    help="rows per pass of the mb_decoder / long_history workloads (defaults: 512 / 96)")
    ap.add_argument("--max-his-len", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cuda-graphs", action="store_true", help="launch every kernel from the host (A/B switch)")
    ap.add_argument("--no-eval", action="store_true", help="skip the secondary constrained-beam-search evaluation leg")
    ap.add_argument("--eval-users", type=int, default=256)
    ap.add_argument("--eval-iters", type=int, default=5)
    ap.add_argument("--cpu-eval-users", type=int, default=4, help="users of the bounded CPU sample of the evaluation leg")
    ap.add_argument("--cpu-sample", type=int, default=8, help="rows of the bounded CPU-baseline sample")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the stock-PyTorch-on-B200 baseline leg")
    ap.add_argument("--gpu-baseline-rows", type=int, default=64, help="rows per step of the stock-PyTorch GPU baseline")
    return ap.parse_args()


def model_config(max_his_len, vocab=1041):
    from transformers.models.qwen3_moe import Qwen3MoeConfig
    cfg = Qwen3MoeConfig.from_pretrained(os.path.join(ROOT, "config", "s2s-models", "Qwen3Multi"))
    cfg.vocab_size = vocab                                   # the mutations of train_SMB_decoder.py:321-360
    cfg.num_behavior = 3
    cfg.behavior_maps = {"526": 0, "527": 1, "528": 2}
    cfg.use_behavior_token = True
    cfg.num_positions = 5
    cfg.num_experts = 6
    cfg.n_positions = max_his_len + 1
    cfg.use_user_token = False
    cfg.model_max_length = max(1024, 5 * (max_his_len + 1))
    return cfg


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.rows, self.stop, self.index = [], False, index
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1] else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_reference_step_fn(max_his_len, rows, seed=0, device="cpu", autocast=False, sdpa=False):
    """The reference arm / cpu_baseline: the fp32 oracle port (oracle/oracle_model.py, validated against the reference's
    own classes) doing forward + backward + AdamW on the host cores.  Returns (step_fn, rows).

    With device="cuda" the same port is the `gpu_baseline` leg: stock PyTorch kernels on the same B200 (materialised
    [B, L, L] masks, F.scaled_dot_product_attention, the per-expert gather loop) in fp32 or under bf16 autocast — what the
    reference executes on a GPU (Qwen3Multi/model.py:123-143, 573-741; Qwen3Moe/FFN.py:53-72)."""
    import torch
    from gamer_b200 import synthetic as syn
    from oracle import oracle_model as om
    spec = om.Spec(temperature=0.7)
    torch.manual_seed(42)
    from gamer_b200 import modeling
    cfg = model_config(max_his_len)
    m = modeling.Qwen3MultiWithTemperature(cfg)              # parameter container only (CPU): reference init N(0, 0.02)
    W = {k: v.detach().clone().to(device).requires_grad_(True) for k, v in m.state_dict().items() if k != "lm_head.weight"}
    W["lm_head.weight"] = W["model.embed_tokens.weight"]
    params = [v for k, v in W.items() if k != "lm_head.weight"]
    opt = torch.optim.AdamW(params, lr=5e-4, weight_decay=0.01)
    cat = syn.make_catalogue(50_000, 1234)
    batch = syn.make_train_batch(cat, rows, max_his_len=max_his_len, seed=seed, full_length=True)
    batch = {k: v.to(device) for k, v in batch.items()}

    def step():
        om.USE_SDPA = sdpa
        try:
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                out = om.forward(spec, W, **batch)
            out["loss"].backward()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
        finally:
            om.USE_SDPA = False
        return out["loss"].detach()

    return step, rows


def gpu_baseline_leg(args, dev):
    """`gpu_baseline`: the stock-PyTorch path on the same B200, timed with CUDA events on a bounded
    sample of the headline workload, in fp32 and under bf16 autocast.  A reported baseline like `cpu_baseline`: the only
    other place bench.py executes oracle/ code, never on the product path."""
    import torch
    out = {"unit": "samples/s", "kind": "port",
           "note": "oracle port on cuda: materialised [B,L,L] additive masks + F.scaled_dot_product_attention + per-expert "
                   "gather loop (the reference's GPU code path), dropout-free (in the baseline's favour), torch AdamW"}
    rows = args.gpu_baseline_rows
    L = 5 * (args.max_his_len + 1)
    for name, ac in (("fp32", False), ("bf16_autocast", True)):
        try:
            step, _ = cpu_reference_step_fn(args.max_his_len, rows, device=dev, autocast=ac, sdpa=True)
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 3
            e0.record()
            for _ in range(n):
                loss = step()
            e1.record()
            torch.cuda.synchronize()
            out[name] = {"value": rows * n / (e0.elapsed_time(e1) / 1e3), "last_loss": float(loss)}
            del step
        except Exception as e:  # noqa: BLE001 (an OOM of the baseline must not lose the bench line)
            out[name] = {"value": None, "error": f"{type(e).__name__}: {str(e)[:120]}"}
        torch.cuda.empty_cache()
    out["sample"] = f"{rows} full-length rows (L={L}) per step x 3 steps after 2 warm-ups: fwd+bwd+clip+AdamW"
    return out


def eval_config(max_his_len):
    from transformers.models.qwen3_moe import Qwen3MoeConfig
    cfg = Qwen3MoeConfig.from_pretrained(os.path.join(ROOT, "config", "s2s-models", "Qwen3SessionMoe"))
    cfg.vocab_size = 1041
    cfg.num_behavior = 3
    cfg.behavior_maps = {"526": 0, "527": 1, "528": 2}
    cfg.use_behavior_token = True
    cfg.num_positions = 5
    cfg.num_experts = 6
    cfg.n_positions = max_his_len + 1
    cfg.use_user_token = False
    cfg.model_max_length = max(1024, 5 * (max_his_len + 1))
    return cfg


EVAL_BEAMS = 20
EVAL_TRIE_ITEMS = 250_000


def cpu_eval_baseline(args, users):
    """Eval half of the reference arm / cpu_baseline: the oracle port of the constrained beam search
    (oracle/oracle_decode.py — HF `_beam_search` + the Python trie walk, as test_SMB_decoder.py:159-177 drives them) on
    the host cores, on `users` users of the same workload (501-token prompts, 20 beams, 250k-item trie)."""
    import torch
    from gamer_b200 import modeling
    from gamer_b200 import synthetic as syn
    from oracle import oracle_decode as od
    from oracle import oracle_model as om
    cfg = eval_config(args.max_his_len)
    torch.manual_seed(43)
    m = modeling.Qwen3SessionMoeWithTemperature(cfg)
    spec = om.Spec.from_hf_config(cfg, "Qwen3SessionMoe", temperature=1.0)
    W = {k: v.detach().float() for k, v in m.state_dict().items()}
    W["lm_head.weight"] = W["model.embed_tokens.weight"]
    cat = syn.make_catalogue(EVAL_TRIE_ITEMS, 1234)
    items = cat.item_sequences(2)
    tree = od.PrefixTree(items.tolist())
    last = set(int(t) for t in items[:, -1]) | {syn.PAD}
    batch, _ = syn.make_eval_batch(cat, users, max_his_len=args.max_his_len, target_behavior=2, seed=77, full_length=True)
    t0 = time.perf_counter()
    with torch.no_grad():
        od.constrained_beam_search(spec, W, tree, last, batch["input_ids"], batch["attention_mask"],
                                   batch.get("session_ids"), batch.get("extended_session_ids"), batch.get("actions"),
                                   num_beams=EVAL_BEAMS, max_new_tokens=4)
    dt = time.perf_counter() - t0
    return {"value": users / dt, "unit": "users/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": f"{users} users x {EVAL_BEAMS} beams, 501-token prompts, 4 new tokens, {EVAL_TRIE_ITEMS // 1000}k-item trie: "
                      f"oracle port of HF beam search + Python trie (fp32), one call"}


def eval_leg(args, dev, world, rank, barrier):
    """Secondary metric of BASELINE.json (configs[2]): trie-constrained beam-search evaluation of Qwen3SessionMoe,
    ShortVideoAD-shaped synthetic users, max_his_len=100 (prompt = 501 tokens, full length), global batch 256 users
    sharded exactly over the ranks, 20 beams, 4 new tokens, 250k-item candidate trie.  Returns the `eval` object of the
    JSON line: users/s with inputs resident in HBM and end to end from pinned host buffers (H2D of the prompts + D2H of the
    decoded item tuples and scores inside the timed region), its own roofline (dominant kernel of one profiled call) and,
    at N=1, the CPU baseline of the same call."""
    import torch
    import torch.distributed as dist
    from gamer_b200 import _cabi, modeling
    from gamer_b200 import synthetic as syn
    from gamer_b200.distributed import shard_range
    from gamer_b200.trie import flat_from_array, prefix_allowed_tokens_fn_by_last_token
    cfg = eval_config(args.max_his_len)
    torch.manual_seed(43)
    model = modeling.Qwen3SessionMoeWithTemperature(cfg).to(dev).eval()
    cat = syn.make_catalogue(EVAL_TRIE_ITEMS, 1234)
    items = cat.item_sequences(2)
    fn = prefix_allowed_tokens_fn_by_last_token(flat_from_array(items), set(int(t) for t in items[:, -1]) | {syn.PAD})
    users, beams = args.eval_users, EVAL_BEAMS
    lo, hi = shard_range(users, rank, world)
    batch, _ = syn.make_eval_batch(cat, users, max_his_len=args.max_his_len, target_behavior=2, seed=77, full_length=True)
    host = {k: v[lo:hi].contiguous().pin_memory() for k, v in batch.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    n = hi - lo
    gen_host = torch.empty(n * beams, 4, dtype=torch.int64).pin_memory()
    score_host = torch.empty(n * beams, dtype=torch.float32).pin_memory()

    def decode(b):
        return model.generate(**b, max_new_tokens=4, prefix_allowed_tokens_fn=fn, num_beams=beams,
                              num_return_sequences=beams, output_scores=True, return_dict_in_generate=True)

    for _ in range(3):
        decode(resident)
    # one profiled call: CUDA events around every entry point (kernel breakdown + the dominant kernel's launch duration)
    # (launched eagerly: the timed calls replay a CUDA graph, inside which single kernels cannot be timed)
    prof = _cabi.Profile()
    _cabi.set_profile(prof)
    model.config.gamer_decode_graphs = False
    decode(resident)
    torch.cuda.synchronize()
    model.config.gamer_decode_graphs = True
    _cabi.set_profile(None)
    summ = prof.summary()
    iters = args.eval_iters
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = {}
    for name in ("value", "e2e"):
        barrier()
        ev0.record()
        for _ in range(iters):
            if name == "value":
                o = decode(resident)
            else:
                o = decode({k: v.to(dev, non_blocking=True) for k, v in host.items()})
                # the result of an evaluation call: the decoded item tuples (the 4 new tokens of every hypothesis) and
                # their scores; the prompts are the caller's own input and stay where they are
                gen_host.copy_(o.generated, non_blocking=True)
                score_host.copy_(o.sequences_scores, non_blocking=True)
                torch.cuda.current_stream().synchronize()
        ev1.record()
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name] = users * iters / (float(t.item()) / 1e3)
    obj = {"metric": "eval_users_per_s", "value": out["value"], "unit": "users/s",
           "config": {"workload": f"Qwen3SessionMoe trie-constrained beam search (configs[2]), max_his_len={args.max_his_len}, "
                                  f"prompt 501 tokens, {users} users per batch, {beams} beams, 4 new tokens, "
                                  f"{EVAL_TRIE_ITEMS // 1000}k-item trie", "users_per_gpu": n, "iters": iters},
           "e2e": {"value": out["e2e"], "unit": "users/s",
                   "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in host.values())) * world,
                   "d2h_bytes_per_step": int(gen_host.numel() * 8 + score_host.numel() * 4) * world}}
    if rank == 0 and summ:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        total = sum(v["ms"] for v in summ.values()) or 1.0
        name, top = max(summ.items(), key=lambda kv: kv[1]["ms"])
        if top["flops"] > 0:
            ach, peak, unit, bound = top["flops"] / (top["ms"] / 1e3) / 1e12, peaks.get("bf16_tflops_sustained", 1400.0), "TFLOP/s", "tensor"
        else:
            ach, peak, unit, bound = top["bytes"] / (top["ms"] / 1e3) / 1e9, peaks.get("hbm_gbs", 6650.0), "GB/s", "hbm"
        obj["roofline"] = {"bound": bound, "kernel": name, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                           "traffic": None, "share_of_call": top["ms"] / total, "avg_launch_ms": top["ms"] / max(1, top["calls"]),
                           "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"}
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get("eval", {}).get(name)
            if tr and users == 256 and world == 1:
                obj["roofline"]["traffic"] = tr["bytes_per_launch"]
                obj["roofline"]["algorithmic_bytes_per_launch"] = top["bytes"] / max(1, top["calls"])
                obj["roofline"]["traffic_source"] = tr["source"]
        except Exception:
            pass
        if "gamer_attn_decode" in summ and name != "gamer_attn_decode":       # the decode-specific kernel, beside the dominant one
            d = summ["gamer_attn_decode"]
            obj["decode_attention"] = {"tflops": d["flops"] / (d["ms"] / 1e3) / 1e12, "gbs": d["bytes"] / (d["ms"] / 1e3) / 1e9,
                                       "avg_launch_ms": d["ms"] / max(1, d["calls"]), "share_of_call": d["ms"] / total}
        obj["kernel_breakdown"] = {k: {"ms_per_call": v["ms"], "share": v["ms"] / total, "launches": v["calls"]}
                                   for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"])[:10]}
        obj["gpu_launches_per_call"] = prof.launches
    return obj


def _traffic_note(r):
    """DRAM traffic a little BELOW the algorithmic bytes is an L2 effect, not an accounting error: say so in the line."""
    if r.get("traffic") and r.get("algorithmic_bytes_per_launch") and r["traffic"] < r["algorithmic_bytes_per_launch"]:
        r["traffic_note"] = ("DRAM traffic below the algorithmic bytes: part of the activation operand is still resident in "
                             "the 126 MB L2 from the kernel that produced it (dram__bytes counts misses only)")


def _both_sides(r, top, peaks):
    """A kernel with FLOPs AND bytes gets the other side of its roofline too: algorithmic GB/s against the HBM peak, its
    arithmetic intensity and the ridge.  Below the ridge the HBM side is the one that binds (the K = 256 projections)."""
    if top["flops"] > 0 and top["bytes"] > 0 and top["ms"] > 0:
        tens_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        gbs = top["bytes"] / (top["ms"] / 1e3) / 1e9
        r["hbm_side"] = {"achieved_gbs": gbs, "frac": gbs / hbm_peak, "flop_per_byte": top["flops"] / top["bytes"],
                         "ridge_flop_per_byte": tens_peak * 1e12 / (hbm_peak * 1e9),
                         "binding": "hbm" if top["flops"] / top["bytes"] < tens_peak * 1e12 / (hbm_peak * 1e9) else "tensor"}
    return r


def _roofline_of(name, top, step_ms, peaks):
    tens_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    if top["flops"] > 0:
        ach, peak, unit, bound = top["flops"] / (top["ms"] / 1e3) / 1e12, tens_peak, "TFLOP/s", "tensor"
    else:
        ach, peak, unit, bound = top["bytes"] / (top["ms"] / 1e3) / 1e9, hbm_peak, "GB/s", "hbm"
    return _both_sides(
        {"bound": bound, "kernel": name, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak, "traffic": None,
         "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)",
         "share_of_step": top["ms"] / step_ms, "avg_launch_ms": top["ms"] / max(1, top["calls"])}, top, peaks)


def _breakdown_of(summ, peaks):
    total = sum(v["ms"] for v in summ.values()) or 1.0
    out = {"_note": "one untimed step with CUDA events around every entry point (adds launch gaps)"}
    for n, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"])[:12]:
        e = {"ms_per_step": v["ms"], "share": v["ms"] / total, "calls_per_step": v["calls"]}
        if v["flops"]:
            e["tflops"] = v["flops"] / (v["ms"] / 1e3) / 1e12
            e["frac_of_peak"] = e["tflops"] / peaks.get("bf16_tflops_sustained", 1400.0)
        elif v["bytes"]:
            e["gbs"] = v["bytes"] / (v["ms"] / 1e3) / 1e9
            e["frac_of_peak"] = e["gbs"] / peaks.get("hbm_gbs", 6650.0)
        out[n] = e
    return out


def run_secondary(args):
    """`--workload mb_decoder` (BASELINE.json configs[3]) and `--workload long_history` (configs[4]): the same step
    (forward + backward + all-reduce + clip + AdamW through NativeTrainer, bf16, dropout on, CUDA graphs) on the other two
    training shapes, same JSON contract; long_history adds a global-batch sweep 256..4096 and the embedding-gather /
    masked-attention rooflines at L = 2505."""
    import torch
    import torch.distributed as dist
    from gamer_b200 import _cabi, modeling
    from gamer_b200 import synthetic as syn
    from gamer_b200.trainer import NativeTrainer
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cat = syn.make_catalogue(250_000, 1234)
    torch.manual_seed(42)
    if args.workload == "mb_decoder":
        mhl, beh_tokens = 200, (526, 527, 528, 1041)
        cfg = model_config(mhl, vocab=1042)
        cfg.num_behavior = 4
        cfg.behavior_maps = {str(t): i for i, t in enumerate(beh_tokens)}
        model = modeling.Qwen3MoeWithTemperature(cfg)
        batches, mb_rows = [args.global_batch if args.global_batch != 1024 else 512], 512
        make = lambda rows, seed: syn.make_mb_batch(cat, rows, max_his_len=mhl, behavior_tokens=beh_tokens, seed=seed)
        what = ("Qwen3Moe multi-behaviour decoder train (configs[3]: train_MB_decoder, 4 behaviour types, no sessions), "
                f"max_his_len={mhl}")
    else:
        mhl = 500
        cfg = model_config(mhl)
        model = modeling.Qwen3MultiWithTemperature(cfg)
        # 96 rows x 3 kv heads = 288 (sequence, kv head) groups = 2 full waves of the attention backward's 148 persistent
        # CTAs (64 rows = 192 groups left a third of the second wave idle: 914 -> 1 050 samples/s)
        batches, mb_rows = [256, 512, 1024, 2048, 4096], 96
        make = lambda rows, seed: syn.make_train_batch(cat, rows, max_his_len=mhl, seed=seed, full_length=True)
        what = f"Qwen3Multi smb_explicit_decoder train, long-history stress (configs[4]), max_his_len={mhl}"
    L = 5 * (mhl + 1)
    if args.workload_micro_batch:
        mb_rows = args.workload_micro_batch
    model.set_hyper(0.7)
    model = model.to(dev).train()
    if world > 1:
        for p in model.parameters():
            dist.broadcast(p.data, src=0)
    trainer = NativeTrainer(model, lr=5e-4, weight_decay=0.01, max_grad_norm=1.0, warmup_steps=2, total_steps=10_000)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sweep, line = [], None
    for gb in batches:
        assert gb % world == 0
        B_local = gb // world
        mb = min(mb_rows, B_local)
        host = {k: v.pin_memory() for k, v in make(B_local, 1000 * rank + gb).items()}
        resident = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        for _ in range(args.warmup):
            trainer.step(resident, micro_batch=mb)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clocks:
            barrier()
            ev0.record()
            for _ in range(args.steps):
                trainer.step(resident, micro_batch=mb)
            ev1.record()
            barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) / args.steps
        barrier()
        ev0.record()
        for _ in range(args.steps):
            loss = trainer.step({k: v.to(dev, non_blocking=True) for k, v in host.items()}, micro_batch=mb)
            loss_host = float(loss.item())
        ev1.record()
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item()) / args.steps
        point = {"global_batch": gb, "micro_batch": mb, "ms_per_step": ms, "samples_per_s": gb / (ms / 1e3),
                 "tokens_per_s": gb * L / (ms / 1e3), "e2e_samples_per_s": gb / (e2e_ms / 1e3)}
        sweep.append(point)
        if gb == batches[-1]:
            # kernel breakdown + rooflines on one eager step of the last (largest) batch
            graphs_on, trainer.use_cuda_graphs = trainer.use_cuda_graphs, False
            prof = _cabi.Profile()
            _cabi.set_profile(prof)
            trainer.step(resident, micro_batch=mb)
            barrier()
            _cabi.set_profile(None)
            trainer.use_cuda_graphs = graphs_on
            summ = prof.summary()
            name, top = max(summ.items(), key=lambda kv: kv[1]["ms"])
            line = {"metric": "train_samples_per_s", "value": point["samples_per_s"], "unit": "samples/s", "n_gpus": world,
                    "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                    "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                    "config": {"workload": f"{what}, L={L}, global batch {gb}, full-length rows, "
                                           "fwd+bwd+allreduce+clip+AdamW", "global_batch": gb, "per_gpu_batch": B_local,
                               "micro_batch": mb, "seq_len": L, "parallelism": f"dp{world}",
                               "tokens_per_s": point["tokens_per_s"],
                               "dropout": f"on (dropout_rate={cfg.dropout_rate}, attention_dropout={cfg.attention_dropout})",
                               "l2_policy": "inputs and activations exceed the 126 MB L2; no flush needed"},
                    "clocks": clocks.summary(), "gpu_launches": prof.launches * args.steps, "cuda_graphs": bool(graphs_on),
                    "e2e": {"value": point["e2e_samples_per_s"], "unit": "samples/s",
                            "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in host.values())) * world,
                            "d2h_bytes_per_step": 4 * world, "last_loss": loss_host},
                    "roofline": _roofline_of(name, top, ms, peaks), "kernel_breakdown": _breakdown_of(summ, peaks)}
            if args.workload == "long_history":
                line["sweep"] = sweep
                # the two rooflines configs[4] names: embedding gather (HBM) and masked attention (tensor) at L = 2505
                for key, label in (("gamer_embed_route_fwd", "embedding_gather"), ("gamer_embed_bwd", "embedding_grad"),
                                   ("gamer_attn_fwd", "masked_attention_fwd"), ("gamer_attn_bwd", "masked_attention_bwd")):
                    if key in summ:
                        line.setdefault("rooflines", {})[label] = _roofline_of(key, summ[key], ms, peaks)
                # ... and the two embedding kernels alone at one full-size launch (512 rows x 2505 tokens: inside the step
                # they run on 64-row passes, where launch latency shows)
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import embed_bench
                line["embedding_kernels"] = embed_bench.run(batch=512, his=mhl, iters=10, dev=f"cuda:{local}")
        del resident, host
        torch.cuda.empty_cache()
    if rank == 0:
        emit(line)
    if world > 1:
        trainer.close()
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, rows = cpu_reference_step_fn(args.max_his_len, args.cpu_sample)
    for _ in range(max(1, min(args.warmup, 1))):
        step()
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    val = rows / t
    L = 5 * (args.max_his_len + 1)
    line = {"impl": "reference", "metric": "train_samples_per_s", "value": val, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"Qwen3Multi smb_explicit_decoder train, max_his_len={args.max_his_len}, L={L}, "
                                   f"global batch {args.global_batch}, full-length rows",
                       "note": "CPU reference arm: each step is a bounded sample of the workload"},
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": f"{rows} full-length rows (L={L}) per step: fwd+bwd+clip+AdamW, fp32, oracle port of "
                                       f"the reference's PyTorch path"},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if not args.no_eval:
        ev = cpu_eval_baseline(args, args.cpu_eval_users)
        line["eval"] = {"metric": "eval_users_per_s", "value": ev["value"], "unit": "users/s", "cpu_baseline": ev,
                        "config": {"workload": f"Qwen3SessionMoe trie-constrained beam search (configs[2]), "
                                               f"max_his_len={args.max_his_len}, prompt 501 tokens, {EVAL_BEAMS} beams, 4 new "
                                               f"tokens, {EVAL_TRIE_ITEMS // 1000}k-item trie; bounded sample: {ev['sample']}"},
                        "e2e": {"value": ev["value"], "unit": "users/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_REAL_STDOUT = None


def _claim_stdout():
    """stdout carries exactly ONE line (the JSON result): everything libraries print to fd 1 on the way (NCCL's version
    banner, ...) is sent to stderr instead."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse()
    _claim_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.workload != "train":
        run_secondary(args)
        return
    import torch
    import torch.distributed as dist
    from gamer_b200 import _cabi, modeling
    from gamer_b200 import synthetic as syn
    from gamer_b200.trainer import NativeTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert args.global_batch % world == 0
    B_local = args.global_batch // world
    L = 5 * (args.max_his_len + 1)

    torch.manual_seed(42)
    cfg = model_config(args.max_his_len)
    model = modeling.Qwen3MultiWithTemperature(cfg)
    model.set_hyper(0.7)
    model = model.to(dev).train()
    if world > 1:                                            # replicas start identical (DDP broadcasts rank 0's weights)
        for p in model.parameters():
            dist.broadcast(p.data, src=0)
    trainer = NativeTrainer(model, lr=5e-4, weight_decay=0.01, max_grad_norm=1.0, warmup_steps=2,
                            total_steps=args.warmup + 2 * args.steps + 8, use_cuda_graphs=not args.no_cuda_graphs)

    # synthetic ShortVideoAD-shaped batches, every row full length (deterministic work per sample); pinned host copies
    cat = syn.make_catalogue(250_000, 1234)
    n_host = 2
    host = []
    for i in range(n_host):
        b = syn.make_train_batch(cat, B_local, max_his_len=args.max_his_len, seed=1000 * rank + i, full_length=True)
        host.append({k: v.pin_memory() for k, v in b.items()})
    resident = [{k: v.to(dev, non_blocking=True) for k, v in b.items()} for b in host]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())
    mb = min(args.micro_batch, B_local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        trainer.step(resident[i % n_host], micro_batch=mb)
    barrier()

    # ---- untimed pass: per-entry-point CUDA-event timing of one step (kernel breakdown; picks the dominant kernel) -----
    # (eager launches: CUDA events cannot be recorded inside a replayed graph)
    graphs_on = trainer.use_cuda_graphs
    trainer.use_cuda_graphs = False
    prof_all = _cabi.Profile()
    _cabi.set_profile(prof_all)
    trainer.step(resident[0], micro_batch=mb)
    barrier()
    breakdown_all = prof_all.summary()
    dominant = max(breakdown_all.items(), key=lambda kv: kv[1]["ms"])[0]
    # the dominant kernel alone, eager, events around each of its launches only (the roofline's launch duration)
    prof = _cabi.Profile(only={dominant})
    _cabi.set_profile(prof)
    trainer.step(resident[1 % n_host], micro_batch=mb)
    barrier()
    _cabi.set_profile(None)
    summ = prof.summary()
    launches_per_step = prof.launches
    trainer.use_cuda_graphs = graphs_on

    # ---- timed region 1: inputs resident in HBM (value) -----------------------------------------------------------------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        ev0.record()
        t_host = time.perf_counter()
        for i in range(args.steps):
            loss = trainer.step(resident[i % n_host], micro_batch=mb)
        host_ms = (time.perf_counter() - t_host) * 1e3 / args.steps      # host-side enqueue time (no sync inside)
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = args.global_batch * args.steps / (ms_max / 1e3)
    launches = launches_per_step * args.steps     # the graph replays the same kernels the eager step launches

    # ---- timed region 2: end to end through the public API with HOST buffers (H2D of inputs + D2H of the loss) ------
    # The step's inputs are the pre-tokenised rows of the input pipeline (collate.PackedSessions in pinned host memory:
    # 4 code ids + behaviour + session per interaction, 2.3 MB per 1024 rows instead of 24.8 MB of padded int64 tensors);
    # they are copied to the device and expanded there by collate_train into the six [B, L] tensors the model takes.
    from gamer_b200 import collate
    stores = [collate.PackedSessions.from_histories(
        syn.train_histories(cat, B_local, max_his_len=args.max_his_len, seed=1000 * rank + i, full_length=True)).pin_memory()
        for i in range(n_host)]
    h2d_bytes = stores[0].nbytes()
    all_rows = torch.arange(B_local, device=dev)

    def e2e_batch(i):
        st = stores[i % n_host].to(dev, non_blocking=True)
        return collate.collate_train(st, all_rows, args.max_his_len, syn.BEHAVIOR_TOKENS, syn.BEHAVIOR_LEVEL, pad=syn.PAD,
                                     width=args.max_his_len + 1)

    chk = e2e_batch(0)                                       # the device collate reproduces the resident batch bit for bit
    assert all(torch.equal(chk[k], resident[0][k]) for k in resident[0]), "device collate differs from the synthetic batch"
    # every step's loss is read back to the host: an asynchronous D2H copy into pinned memory per step, consumed one step
    # later (the step it belongs to has finished by then), so the host never stalls the next step's enqueue
    loss_pin = torch.zeros(args.steps, dtype=torch.float32).pin_memory()
    loss_ev = [torch.cuda.Event() for _ in range(args.steps)]
    loss_host = float("nan")
    barrier()
    ev0.record()
    for i in range(args.steps):
        loss = trainer.step(e2e_batch(i), micro_batch=mb)
        loss_pin[i:i + 1].copy_(loss.reshape(1), non_blocking=True)
        loss_ev[i].record()
        if i > 0:
            loss_ev[i - 1].synchronize()
            loss_host = float(loss_pin[i - 1])
    loss_ev[-1].synchronize()
    loss_host = float(loss_pin[args.steps - 1])
    ev1.record()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = args.global_batch * args.steps / (float(t.item()) / 1e3)

    eval_obj = None
    graphs_flag = bool(trainer.use_cuda_graphs)
    trainer.close()                                          # graphs that hold NCCL kernels go before the process group
    if not args.no_eval:
        del trainer, resident
        torch.cuda.empty_cache()
        eval_obj = eval_leg(args, dev, world, rank, barrier)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tens_peak = peaks.get("bf16_tflops_sustained", 1400.0)     # kernels timed inside a long step: sustained figure
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
        name, top = dominant, summ[dominant]
        step_ms = ms_max / args.steps
        if top["flops"] > 0:
            ach = top["flops"] / (top["ms"] / 1e3) / 1e12
            roof = {"bound": "tensor", "kernel": name, "achieved": ach, "peak": tens_peak, "unit": "TFLOP/s",
                    "frac": ach / tens_peak, "traffic": None}
        else:
            ach = top["bytes"] / (top["ms"] / 1e3) / 1e9
            roof = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak, "traffic": None}
        try:   # DRAM traffic of the dominant kernel from the committed ncu --set full capture (per launch)
            # captures are keyed by the micro-batch rows they were taken at (512 and 1024)
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(str(mb), {}).get(name)
            if tr:
                roof["traffic"] = tr["bytes_per_launch"]
                roof["traffic_unit"] = "bytes/launch"
                roof["algorithmic_bytes_per_launch"] = top["bytes"] / max(1, top["calls"])
                roof["traffic_source"] = tr["source"]
                _traffic_note(roof)
        except Exception:
            pass
        roof["peak_source"] = peak_src
        _both_sides(roof, top, peaks)
        roof["share_of_step"] = top["ms"] / step_ms
        roof["avg_launch_ms"] = top["ms"] / max(1, top["calls"])
        roof["note"] = ("launch duration: CUDA events around every launch of this entry point during one eager step of "
                        "the same workload, run just before the timed region (the timed steps replay a CUDA graph, inside "
                        "which single kernels cannot be timed)")
        kern = sorted(breakdown_all.items(), key=lambda kv: -kv[1]["ms"])
        total_kernel_ms = sum(v["ms"] for _, v in kern) or 1.0
        breakdown = {"_note": "one untimed step with CUDA events around every entry point (adds launch gaps); attention "
                              "FLOPs are counted on the causal pair count, backward = 2.5 x forward (five GEMMs: S recompute, "
                              "dP, dV, dK, dQ) — multiply gamer_attn_bwd's tflops by 0.8 for SURVEY 8(d)'s 2 x forward convention"}
        for n, v in kern[:12]:
            e = {"ms_per_step": v["ms"], "share": v["ms"] / total_kernel_ms, "calls_per_step": v["calls"]}
            if v["flops"]:
                e["tflops"] = v["flops"] / (v["ms"] / 1e3) / 1e12
                e["frac_of_peak"] = e["tflops"] / tens_peak
            elif v["bytes"]:
                e["gbs"] = v["bytes"] / (v["ms"] / 1e3) / 1e9
                e["frac_of_peak"] = e["gbs"] / hbm_peak
            breakdown[n] = e
        # the other kernels the north star names a roofline target for, each with its like-for-like DRAM traffic when a
        # capture at this micro-batch is committed (profiles/ncu_traffic.json)
        rooflines = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(str(mb), {})
        except Exception:
            traffic = {}
        for n in ("gamer_gemm_bf16_tn", "gamer_gemm_bf16_wgrad", "gamer_attn_fwd", "gamer_attn_bwd", "gamer_embed_route_fwd",
                  "gamer_embed_bwd"):
            if n in breakdown_all and breakdown_all[n]["ms"] > 0:
                r = _roofline_of(n, breakdown_all[n], step_ms, peaks)
                tr = traffic.get(n)
                if tr:
                    r["traffic"] = tr["bytes_per_launch"]
                    r["algorithmic_bytes_per_launch"] = breakdown_all[n]["bytes"] / max(1, breakdown_all[n]["calls"])
                    r["traffic_source"] = tr["source"]
                    _traffic_note(r)
                rooflines[n] = r
        line = {"metric": "train_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"Qwen3Multi smb_explicit_decoder train (configs[1]), max_his_len={args.max_his_len}, "
                                       f"L={L}, global batch {args.global_batch}, full-length rows, fwd+bwd+allreduce+clip+AdamW",
                           "global_batch": args.global_batch, "per_gpu_batch": B_local, "micro_batch": mb, "seq_len": L,
                           "parallelism": f"dp{world}", "tokens_per_s": value * L, "dropout": f"on (train mode: dropout_rate={cfg.dropout_rate}, attention_dropout={cfg.attention_dropout}, "
                                      f"Philox masks regenerated in the backward)",
                           "l2_policy": "inputs and activations (>1 GB per micro-batch) exceed the 126 MB L2; no flush needed"},
                "clocks": clocks.summary(), "gpu_launches": launches, "host_enqueue_ms_per_step": host_ms,
                "cuda_graphs": graphs_flag,
                "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes * world,
                        "d2h_bytes_per_step": 4 * world, "last_loss": loss_host,
                        "note": "inputs: pre-tokenised PackedSessions rows in pinned host memory -> H2D -> collate_train on the "
                                "device; every step's loss is copied D2H asynchronously and consumed by the host one step later"},
                "roofline": roof, "rooflines": rooflines, "kernel_breakdown": breakdown}
        if eval_obj is not None:
            line["eval"] = eval_obj
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            step, rows = cpu_reference_step_fn(args.max_his_len, args.cpu_sample)
            step()
            t0 = time.perf_counter()
            n_cpu = 2
            for _ in range(n_cpu):
                step()
            dt = (time.perf_counter() - t0) / n_cpu
            line["cpu_baseline"] = {"value": rows / dt, "unit": "samples/s", "cores": cores, "kind": "port",
                                    "sample": f"{rows} full-length rows (L={L}) per step x {n_cpu} steps after 1 warm-up: "
                                              f"fwd+bwd+clip+AdamW, fp32 oracle port"}
            if eval_obj is not None:
                line["eval"]["cpu_baseline"] = cpu_eval_baseline(args, args.cpu_eval_users)
        if world == 1 and not args.no_gpu_baseline:
            line["gpu_baseline"] = gpu_baseline_leg(args, dev)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
