"""Facade with the reference's module paths: `from SeqRec.models.generative.Qwen3Multi import
Qwen3MultiWithTemperature` etc. resolve to the B200-native implementations in gamer_b200."""
