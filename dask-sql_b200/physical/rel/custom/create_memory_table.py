"""CREATE TABLE|VIEW <name> AS <query> (dask_sql/physical/rel/custom/create_memory_table.py:14-76).

Equivalent to context.create_table(name, context.sql(query)).  Like the reference, a TABLE is
persisted -- here: executed once, its result columns stay in HBM as a device table -- and a VIEW is
not (the lazy frame is stored and re-executed, fused into whatever query reads it)."""
import logging

from ..base import BaseRelPlugin

logger = logging.getLogger(__name__)


class CreateMemoryTablePlugin(BaseRelPlugin):
    class_name = ["CreateMemoryTable", "CreateView"]

    def convert(self, rel, context):
        cmt = rel.create_memory_table()
        qualified = cmt.getQualifiedName()
        *schema_name, table_name = qualified.split(".")
        if len(schema_name) > 1:
            raise RuntimeError(f"Expected unqualified or fully qualified table name, got {qualified}.")
        schema_name = context.schema_name if not schema_name else schema_name[0]
        if schema_name not in context.schema:
            raise RuntimeError(f"A schema with the name {schema_name} is not present.")
        if table_name.lower() in context.schema[schema_name].tables:
            if cmt.getIfNotExists():
                return
            elif not cmt.getOrReplace():
                raise RuntimeError(f"A table with the name {table_name} is already present.")
        input_rel = cmt.getInput()
        persist = cmt.isTable()
        logger.debug(f"Creating new table with name {qualified} and logical plan {input_rel}")
        context.create_table(table_name, context._compute_table_from_rel(input_rel), persist=persist,
                             schema_name=schema_name)
