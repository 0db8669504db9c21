#!/usr/bin/env python
"""Entry point with the reference's CLI surface for the two hot-path tasks (reference: main.py:11-49,
SeqRec/utils/parse.py:10-58, SeqRec/tasks/train_SMB_decoder.py:23-137, SeqRec/tasks/test_SMB_decoder.py:40-64):

    python main.py train_SMB_decoder --backbone Qwen3Multi --base_model ./config/s2s-models/Qwen3Multi ...
    torchrun --nproc_per_node=8 main.py train_SMB_decoder ...
    python main.py test_SMB_decoder --backbone Qwen3Multi --ckpt_path ./checkpoint/decoder ...

Same flag names, types and defaults; unknown arguments only warn, as in the reference.  Two extra flags size the
synthetic ShortVideoAD-shaped data that stands in for the reference's Git-LFS datasets (gamer_b200/tasks.py).
"""
import argparse
import sys


def parse_global_args(parser):
    parser.add_argument("--seed", type=int, default=42)
    parser.add_argument("--backbone", type=str, default="TIGER")
    parser.add_argument("--base_model", type=str, default="./config/s2s-models/TIGER")
    parser.add_argument("--output_dir", type=str, default="./checkpoint/decoder")
    return parser


def parse_dataset_args(parser):
    parser.add_argument("--data_path", type=str, default="./data")
    parser.add_argument("--tasks", type=str, default="seqrec")
    parser.add_argument("--dataset", type=str, default="Instruments")
    parser.add_argument("--index_file", type=str, default=".index.json")
    parser.add_argument("--max_his_len", type=int, default=20)
    return parser


def add_train(sub):
    p = sub.add_parser("train_SMB_decoder")
    parse_dataset_args(parse_global_args(p))
    p.add_argument("--optim", type=str, default="adamw_torch")
    p.add_argument("--epochs", type=int, default=200)
    p.add_argument("--learning_rate", type=float, default=5e-4)
    p.add_argument("--per_device_batch_size", type=int, default=256)
    p.add_argument("--gradient_accumulation_steps", type=int, default=2)
    p.add_argument("--logging_step", type=int, default=30)
    p.add_argument("--model_max_length", type=int, default=1024)
    p.add_argument("--weight_decay", type=float, default=0.01)
    p.add_argument("--resume_from_checkpoint", type=str, default=None)
    p.add_argument("--warmup_ratio", type=float, default=0.1)
    p.add_argument("--lr_scheduler_type", type=str, default="cosine")
    p.add_argument("--save_and_eval_strategy", type=str, default="epoch")
    p.add_argument("--save_and_eval_steps", type=int, default=1000)
    p.add_argument("--patience", type=int, default=20)
    p.add_argument("--fp16", action="store_true", default=False)
    p.add_argument("--bf16", action="store_true", default=False)
    p.add_argument("--deepspeed", type=str, default=None)
    p.add_argument("--temperature", type=float, default=1.0)
    p.add_argument("--wandb_run_name", type=str, default="default")
    p.add_argument("--synthetic_users", type=int, default=4096, help="(gamer_b200) synthetic users per epoch")
    p.add_argument("--synthetic_items", type=int, default=50_000, help="(gamer_b200) synthetic catalogue size")


def add_test(sub):
    p = sub.add_parser("test_SMB_decoder")
    parse_dataset_args(parse_global_args(p))
    p.add_argument("--ckpt_path", type=str, default="./checkpoint")
    p.add_argument("--results_file", type=str, default="./results/test.json")
    p.add_argument("--test_batch_size", type=int, default=16)
    p.add_argument("--num_beams", type=int, default=20)
    p.add_argument("--metrics", type=str, default="hit@1,hit@5,hit@10,recall@1,recall@5,recall@10,ndcg@5,ndcg@10")
    p.add_argument("--test_task", type=str, default="SeqRec")
    p.add_argument("--behaviors", type=str, nargs="+", default=None)
    p.add_argument("--valid_loss", action="store_true")
    p.add_argument("--synthetic_users", type=int, default=1024, help="(gamer_b200) synthetic test users per behaviour")
    p.add_argument("--synthetic_items", type=int, default=50_000, help="(gamer_b200) synthetic catalogue size")


def build_parser():
    parser = argparse.ArgumentParser()
    sub = parser.add_subparsers(dest="pipeline", title="Available pipelines", description="Choose a pipeline to run",
                                help="Which pipeline to run", required=True)
    add_train(sub)
    add_test(sub)
    return parser


def main(argv=None):
    args, unknown = build_parser().parse_known_args(argv)
    if unknown:
        print(f"[gamer_b200] warning: unknown args: {unknown}", file=sys.stderr)
    name = args.pipeline
    del args.pipeline
    from gamer_b200 import tasks
    return getattr(tasks, name)(**vars(args))


if __name__ == "__main__":
    main()
