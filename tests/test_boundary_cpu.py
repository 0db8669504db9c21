"""CPU: host-side boundary checks that need no GPU — the C-ABI library loads and exports every symbol the header
declares, the drop-in classes build with the reference's state-dict keys, and the product refuses to run without CUDA."""
import os
import subprocess
import sys

import pytest
import torch

from tests.helpers import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    from gamer_b200 import _cabi
    lib = _cabi.lib()
    names = _cabi.declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    assert lib.gamer_last_error() == b""
    # size helpers are pure host functions: callable without a GPU
    assert lib.gamer_embed_sort_bytes(1000, 1041) == (4 * 1041 + 2 + 1000) * 4
    assert lib.gamer_route_perm_workspace_bytes(7) == 7 * 8 * 2 * 4


@pytest.mark.parametrize("name", ["train_qwen3multi.pt", "train_qwen3sessionmoe.pt", "train_qwen3sessionmulti.pt",
                                  "train_qwen3moe.pt"])
def test_drop_in_classes_hold_reference_state_dict(name):
    from tests.test_model_gpu import build_model
    from gamer_b200 import engine as E
    g = load_golden(name)
    m = build_model(g)
    sd = m.state_dict()
    assert set(sd) == set(g["shapes"])
    assert all(tuple(sd[k].shape) == tuple(g["shapes"][k]) for k in sd)
    assert m.lm_head.weight is m.model.embed_tokens.weight
    assert set(E.param_names(m.arch)) == set(dict(m.named_parameters()))


def test_no_cpu_fallback():
    from tests.test_model_gpu import build_model
    g = load_golden("train_qwen3multi.pt")
    m = build_model(g).cpu()
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(input_ids=g["batch"]["input_ids"], attention_mask=g["batch"]["attention_mask"], actions=g["batch"]["actions"])


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gamer_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_seqrec_facade_paths_in_subprocess():
    """The reference's import paths resolve to the drop-in classes (own process: `SeqRec` must not be the reference's)."""
    code = (
        "from SeqRec.models.generative.Qwen3Multi import Qwen3MultiWithTemperature as A\n"
        "from SeqRec.models.generative.Qwen3SessionMoe import Qwen3SessionMoeWithTemperature as B\n"
        "from SeqRec.models.generative.Qwen3SessionMulti import Qwen3SessionMultiWithTemperature as C\n"
        "from SeqRec.generation.trie import Trie, prefix_allowed_tokens_fn_by_last_token\n"
        "from SeqRec.evaluation.ranking import get_topk_results, get_metrics_results\n"
        "import gamer_b200.modeling as m\n"
        "assert A is m.Qwen3MultiWithTemperature and B is m.Qwen3SessionMoeWithTemperature\n"
        "print('ok')\n")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_ranking_matches_reference_golden():
    from gamer_b200 import ranking
    g = load_golden("decode_qwen3multi_lvl2.pt")
    K = g["num_beams"]
    hits = ranking.get_topk_results(g["rank_pred"], g["sequences_scores"], g["rank_targets"], K)
    assert hits == g["rank_hits"]
    met = ranking.get_metrics_results(hits, g["rank_names"], g["rank_targets"])
    for k, v in g["rank_metrics"].items():
        assert abs(met[k] - v) < 1e-12, k


def test_moe_behavior_only_is_rejected():
    """Qwen3Multi/router.py:31-48: the two-expert routing is not implemented; it must not be silently replaced."""
    import pytest
    from transformers.models.qwen3_moe import Qwen3MoeConfig
    from gamer_b200 import modeling
    cfg = Qwen3MoeConfig.from_pretrained(os.path.join(ROOT, "config", "s2s-models", "Qwen3Multi"))
    cfg.num_positions, cfg.model_max_length, cfg.Moe_behavior_only = 5, 1024, True
    with pytest.raises(NotImplementedError):
        modeling.Qwen3MultiWithTemperature(cfg)
