"""Statements around the path: CREATE TABLE ... WITH / AS, CREATE VIEW ... AS, DROP TABLE
(dask_sql/physical/rel/custom/)."""
from .create_memory_table import CreateMemoryTablePlugin
from .create_table import CreateTablePlugin
from .drop_table import DropTablePlugin

__all__ = [CreateMemoryTablePlugin, CreateTablePlugin, DropTablePlugin]
