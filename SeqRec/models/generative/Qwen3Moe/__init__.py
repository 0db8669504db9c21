from SeqRec.models.generative.Qwen3Moe.model import Qwen3MoeWithTemperature  # noqa: F401
