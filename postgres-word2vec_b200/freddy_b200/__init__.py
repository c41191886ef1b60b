"""freddy_b200 — host-side mirror of the FREDDY (postgres-word2vec) search UDFs
over the B200-native engine in ../libfreddy_b200.so."""
from . import _lib
from .engine import Engine, FreddyError, round_through_text

__all__ = ["Engine", "FreddyError", "round_through_text", "_lib"]
