"""Reference path SeqRec/models/generative/Qwen3SessionMoe/model.py -> gamer_b200.modeling."""
from gamer_b200.modeling import Qwen3SessionMoeWithTemperature  # noqa: F401
