"""Reference path SeqRec/models/generative/Qwen3SessionMulti/model.py -> gamer_b200.modeling."""
from gamer_b200.modeling import Qwen3SessionMultiWithTemperature  # noqa: F401
