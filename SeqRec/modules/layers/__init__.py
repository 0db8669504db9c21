from SeqRec.modules.layers.transformer import TransformerEncoderLayer, TransformerEncoder  # noqa: F401
