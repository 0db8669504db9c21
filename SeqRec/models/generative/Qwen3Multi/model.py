"""Reference path SeqRec/models/generative/Qwen3Multi/model.py -> gamer_b200.modeling."""
from gamer_b200.modeling import Qwen3MultiWithTemperature  # noqa: F401
