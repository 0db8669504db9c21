"""Reference path SeqRec/models/generative/Qwen3Moe/model.py -> gamer_b200.modeling."""
from gamer_b200.modeling import Qwen3MoeWithTemperature  # noqa: F401
