"""aligner_b200 -- B200-native (sm_100a) monotonic alignment search and the score
matrices that feed it: the one hot path of xiaozhah/Aligner, behind the
reference's own ``monotonic_align.maximum_path(value, mask)`` API.

    import aligner_b200.monotonic_align as monotonic_align   # or: import monotonic_align (repo root on sys.path)
    path = monotonic_align.maximum_path(neg_cent, attn_mask)

Everything computes in hand-written CUDA reached through a C ABI
(include/aligner_b200.h, libaligner_b200.so); there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (raises ImportError if the CUDA library is not built)
from .monotonic_align import maximum_path, maximum_path_c, maximum_path_lengths, maximum_path_vits  # noqa: F401
from .sharding import balance_shards, lpt_order  # noqa: F401  (multi-GPU / ragged-batch planning)

__version__ = "0.1.0"
