from .convert import RelConverter

__all__ = ["RelConverter"]
