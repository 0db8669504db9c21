"""CPU: main.py keeps the reference's flag surface for the two hot-path tasks (names, types and defaults frozen from
/root/reference/SeqRec/utils/parse.py:10-58, tasks/train_SMB_decoder.py:23-137, tasks/test_SMB_decoder.py:40-64)."""
import main as cli

GLOBAL = {"seed": 42, "backbone": "TIGER", "base_model": "./config/s2s-models/TIGER", "output_dir": "./checkpoint/decoder",
          "data_path": "./data", "tasks": "seqrec", "dataset": "Instruments", "index_file": ".index.json", "max_his_len": 20}
TRAIN = {"optim": "adamw_torch", "epochs": 200, "learning_rate": 5e-4, "per_device_batch_size": 256,
         "gradient_accumulation_steps": 2, "logging_step": 30, "model_max_length": 1024, "weight_decay": 0.01,
         "resume_from_checkpoint": None, "warmup_ratio": 0.1, "lr_scheduler_type": "cosine",
         "save_and_eval_strategy": "epoch", "save_and_eval_steps": 1000, "patience": 20, "fp16": False, "bf16": False,
         "deepspeed": None, "temperature": 1.0, "wandb_run_name": "default"}
TEST = {"ckpt_path": "./checkpoint", "results_file": "./results/test.json", "test_batch_size": 16, "num_beams": 20,
        "metrics": "hit@1,hit@5,hit@10,recall@1,recall@5,recall@10,ndcg@5,ndcg@10", "test_task": "SeqRec",
        "behaviors": None, "valid_loss": False}


def test_defaults_match_reference():
    p = cli.build_parser()
    a = vars(p.parse_args(["train_SMB_decoder"]))
    for k, v in {**GLOBAL, **TRAIN}.items():
        assert a[k] == v, k
    b = vars(p.parse_args(["test_SMB_decoder"]))
    for k, v in {**GLOBAL, **TEST}.items():
        assert b[k] == v, k


def test_launcher_style_invocation_parses():
    # the flags scripts/train_SMB_decoder.sh:123-153 and scripts/test_SMB_decoder.sh pass
    p = cli.build_parser()
    a, unknown = p.parse_known_args(
        "train_SMB_decoder --seed 42 --backbone Qwen3Multi --base_model ./config/s2s-models/Qwen3Multi --output_dir ckpt "
        "--data_path ./data --tasks smb_explicit_decoder_4 --dataset ShortVideoAD --index_file .index.json --max_his_len 100 "
        "--epochs 200 --learning_rate 5e-4 --per_device_batch_size 128 --gradient_accumulation_steps 4 --temperature 0.7 "
        "--patience 20 --wandb_run_name x --not_a_flag 1".split())
    assert a.backbone == "Qwen3Multi" and a.max_his_len == 100 and a.temperature == 0.7 and unknown == ["--not_a_flag", "1"]
    b = p.parse_args("test_SMB_decoder --backbone Qwen3Multi --ckpt_path ckpt --test_batch_size 32 --num_beams 20 "
                     "--behaviors behavior_2 behavior_1 --results_file r.json".split())
    assert b.behaviors == ["behavior_2", "behavior_1"] and b.num_beams == 20
