"""Reference path SeqRec/generation/trie.py -> gamer_b200.trie."""
from gamer_b200.trie import Trie, prefix_allowed_tokens_fn, prefix_allowed_tokens_fn_by_last_token  # noqa: F401
