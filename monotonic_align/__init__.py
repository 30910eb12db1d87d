"""``import monotonic_align`` drop-in (reference: monotonic_align/__init__.py).
Re-exports the B200 implementation so code written against the reference package
keeps its import line."""
from aligner_b200.monotonic_align import maximum_path, maximum_path_c, maximum_path_lengths, maximum_path_vits  # noqa: F401
