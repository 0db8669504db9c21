"""Reference path SeqRec/modules -> gamer_b200.modules (interfaces kept importable; outside the accelerated hot path)."""
