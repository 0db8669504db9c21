from SeqRec.models.generative.Qwen3SessionMoe.model import Qwen3SessionMoeWithTemperature  # noqa: F401
