"""dask_sql_b200 — B200-native execution layer for dask-sql's filter -> join -> group-by hot path."""
from . import _lib  # noqa: F401  (fails loudly if libb200sql.so is missing)

__version__ = "0.1.0"
