"""Relational side of the plugin boundary."""
from .convert import RelConverter  # noqa: F401  (re-exported)
