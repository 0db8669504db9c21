"""Reference path SeqRec/modules/layers/transformer.py -> gamer_b200.modules."""
from gamer_b200.modules import (FeedForward, MultiHeadAttention, TransformerEncoder,  # noqa: F401
                                TransformerEncoderLayer)
