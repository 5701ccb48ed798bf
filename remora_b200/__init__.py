"""remora_b200: B200-native (sm_100a) implementation of Remora's per-chunk modified-base
inference hot path behind the reference's own Python surface
(``model_util.load_model`` / ``inference.call_read_mods`` / ``data_chunks.RemoraRead``).

The compute path is hand-written CUDA reached through a C-ABI shared library
(include/remora_b200.h).  There is no CPU fallback: if the library is missing the ops raise.
"""

__version__ = "0.1.0"


class RemoraError(Exception):
    """Same role as ``remora.RemoraError`` (reference src/remora/__init__.py:4-7)."""
    pass
