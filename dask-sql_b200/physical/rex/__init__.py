"""Expression side of the plugin boundary."""
from .convert import RexConverter  # noqa: F401  (re-exported)
