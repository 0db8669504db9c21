from SeqRec.models.generative.Qwen3SessionMulti.model import Qwen3SessionMultiWithTemperature  # noqa: F401
