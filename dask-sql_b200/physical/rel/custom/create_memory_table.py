"""CREATE TABLE|VIEW <name> AS <query> (dask_sql/physical/rel/custom/create_memory_table.py:14-76).

Equivalent to context.create_table(name, context.sql(query)).  Like the reference, a TABLE is
persisted -- here: executed once, its result columns stay in HBM as a device table -- and a VIEW is
not (the lazy frame is stored and re-executed, fused into whatever query reads it)."""
from ..base import BaseRelPlugin
from ._target import may_create, split_qualified


class CreateMemoryTablePlugin(BaseRelPlugin):
    class_name = ["CreateMemoryTable", "CreateView"]

    def convert(self, rel, context):
        stmt = rel.create_memory_table()
        schema, table = split_qualified(context, stmt.getQualifiedName())
        if not may_create(context, schema, table, stmt.getIfNotExists(), stmt.getOrReplace()):
            return
        result = context._compute_table_from_rel(stmt.getInput())
        context.create_table(table, result, persist=stmt.isTable(), schema_name=schema)
