"""GPU: the two tasks end to end through main.py with the reference's flags — train a few steps on synthetic sessions,
save the best checkpoint, reload it with from_pretrained in the test task, constrained beam search + ranking metrics,
results JSON in the reference's layout (tasks/test_SMB_decoder.py:287-304,534-537)."""
import json
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("backbone", ["Qwen3Multi", "Qwen3SessionMoe"])
def test_train_then_test_tasks(tmp_path, backbone):
    import main as cli
    ckpt, res = str(tmp_path / "ckpt"), str(tmp_path / "results" / "test.json")
    base = os.path.join(ROOT, "config", "s2s-models", backbone)
    best = cli.main(f"train_SMB_decoder --backbone {backbone} --base_model {base} --output_dir {ckpt} --dataset ShortVideoAD "
                    f"--tasks smb_explicit_decoder --max_his_len 10 --epochs 3 --per_device_batch_size 16 "
                    f"--gradient_accumulation_steps 2 --learning_rate 2e-3 --temperature 0.7 --logging_step 1 "
                    f"--synthetic_users 64 --synthetic_items 2000".split())
    assert best == best and best < 7.2, best            # finite, below ln(1041) = 6.95 + margin after 6 steps
    assert os.path.exists(os.path.join(ckpt, "config.json"))
    out = cli.main(f"test_SMB_decoder --backbone {backbone} --ckpt_path {ckpt} --results_file {res} --max_his_len 10 "
                   f"--test_batch_size 16 --num_beams 5 --metrics hit@1,hit@5,recall@5,ndcg@5 --behaviors behavior_2 behavior_0 "
                   f"--synthetic_users 32 --synthetic_items 2000".split())
    saved = json.load(open(res))
    assert [r["eval_type"] for r in saved] == ["Behavior behavior_2", "Behavior behavior_0", "Merged Behavior"]
    for r in saved:
        for m in ("hit@1", "hit@5", "recall@5", "ndcg@5"):
            assert 0.0 <= r[m] <= 1.0
        assert r["hit@1"] <= r["hit@5"]
    assert saved == out
    vl = cli.main(f"test_SMB_decoder --backbone {backbone} --ckpt_path {ckpt} --max_his_len 10 --test_batch_size 16 "
                  f"--valid_loss --synthetic_users 64 --synthetic_items 2000".split())
    assert abs(vl - best) < 0.5


def test_device_ranking_glue_on_generate_output_equals_string_metrics():
    """(f)-3: hit matching and hit / recall / ndcg sums computed on the GPU from a real constrained-beam-search output
    (code-id tuples, targets packed on the device) equal the reference-shaped string functions (evaluation/ranking.py:5-90)
    fed with the same hypotheses as strings — including users whose targets were planted on their own beams."""
    import torch
    from gamer_b200 import ranking as R
    from gamer_b200 import synthetic as syn
    from gamer_b200.trie import flat_from_array, prefix_allowed_tokens_fn_by_last_token
    from tests.helpers import load_golden
    from tests.test_model_gpu import build_model
    g = load_golden("decode_qwen3multi_lvl2.pt")
    m = build_model(g, temperature=1.0).eval()
    cat = syn.make_catalogue(g["catalogue_size"], g["catalogue_seed"])
    items = cat.item_sequences(g["target_behavior"])
    fn = prefix_allowed_tokens_fn_by_last_token(flat_from_array(items), set(int(t) for t in items[:, -1]) | {syn.PAD})
    users, K, S = 12, 8, 4
    batch, targets = syn.make_eval_batch(cat, users, max_his_len=14, target_behavior=g["target_behavior"], seed=9, median_len=7)
    out = m.generate(**{k: v.cuda() for k, v in batch.items()}, max_new_tokens=S, prefix_allowed_tokens_fn=fn, num_beams=K,
                     num_return_sequences=K)
    gen = out.generated.view(users, K, S)
    host = gen.cpu()
    for u in range(0, users, 2):                                 # plant hits at ranks 0, 3 and 7 of every other user
        targets[u] = list(targets[u]) + [tuple(host[u, r].tolist()) for r in (0, 3, 7)]
    metrics = ["hit@1", "hit@5", "recall@1", "recall@5", "recall@10", "ndcg@5", "ndcg@10"]
    tt, cnt = R.pack_targets(targets, S, device="cuda")
    hits = R.topk_hits(gen, tt)
    got = R.metric_sums(hits, cnt, metrics)
    assert hits.is_cuda
    pred = ["".join(f"<{int(t)}>" for t in host[u, k]) for u in range(users) for k in range(K)]
    tstr = [["".join(f"<{t}>" for t in tup) for tup in tl] for tl in targets]
    ref_hits = R.get_topk_results(pred, out.sequences_scores.cpu().tolist(), tstr, K)
    want = R.get_metrics_results(ref_hits, metrics, tstr)
    assert hits.cpu().tolist() == ref_hits
    for name in metrics:
        assert abs(float(got[name]) - want[name]) < 1e-9, (name, float(got[name]), want[name])
    assert want["hit@1"] >= users // 2


def test_tasks_on_reference_format_files_with_augmentation(tmp_path):
    """--data_path/--dataset/--index_file pointing at files in the reference's format: loaded by gamer_b200.dataset
    (ids from the sorted added tokens, session splits), trained with the `smb_explicit_decoder_4` augmentation, evaluated
    with the trie built from the index file."""
    import main as cli
    from gamer_b200 import dataset as ds
    ds.write_synthetic_files(str(tmp_path / "data"), "toy", n_users=96, n_items=400, seed=1)
    data = ds.load_smb_files(str(tmp_path / "data"), "toy")
    ckpt, res = str(tmp_path / "ckpt"), str(tmp_path / "results" / "test.json")
    base = os.path.join(ROOT, "config", "s2s-models", "Qwen3Multi")
    best = cli.main(f"train_SMB_decoder --backbone Qwen3Multi --base_model {base} --output_dir {ckpt} --data_path "
                    f"{tmp_path / 'data'} --dataset toy --index_file .index.json --tasks smb_explicit_decoder_4 "
                    f"--max_his_len 10 --epochs 2 --per_device_batch_size 16 --gradient_accumulation_steps 2 "
                    f"--learning_rate 2e-3 --temperature 0.7 --logging_step 4".split())
    assert best == best and best < 7.5, best
    cfg = json.load(open(os.path.join(ckpt, "config.json")))
    assert cfg["vocab_size"] == data.vocab_size and cfg["num_behavior"] == 3
    out = cli.main(f"test_SMB_decoder --backbone Qwen3Multi --ckpt_path {ckpt} --results_file {res} --data_path "
                   f"{tmp_path / 'data'} --dataset toy --index_file .index.json --max_his_len 10 --test_batch_size 32 "
                   f"--num_beams 5 --metrics hit@1,hit@5,ndcg@5".split())
    assert [r["eval_type"] for r in out] == ["Behavior click", "Behavior cart", "Behavior buy", "Merged Behavior"]
    for r in out:
        assert 0.0 <= r["hit@1"] <= r["hit@5"] <= 1.0
