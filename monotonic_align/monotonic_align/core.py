"""``from monotonic_align.monotonic_align.core import maximum_path_c`` (reference __init__.py:3)."""
from aligner_b200.monotonic_align.monotonic_align.core import maximum_path_c  # noqa: F401
