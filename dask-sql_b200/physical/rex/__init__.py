from .convert import RexConverter

__all__ = ["RexConverter"]
