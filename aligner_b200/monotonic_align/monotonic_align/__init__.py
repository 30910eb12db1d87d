"""Mirrors the reference's inner package (monotonic_align/monotonic_align/, built by its setup.py:11)."""
