from SeqRec.models.generative.Qwen3Multi.model import Qwen3MultiWithTemperature  # noqa: F401
