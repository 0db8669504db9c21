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
