"""DROP TABLE [IF EXISTS] <name> (dask_sql/physical/rel/custom/drop_table.py)."""
from ..base import BaseRelPlugin
from ._target import split_qualified, table_exists


class DropTablePlugin(BaseRelPlugin):
    class_name = "DropTable"

    def convert(self, rel, context):
        stmt = rel.drop_table()
        schema, table = split_qualified(context, stmt.getQualifiedName())
        if table_exists(context, schema, table):
            context.drop_table(table, schema_name=schema)
        elif not stmt.getIfExists():
            raise RuntimeError(f"A table with the name {stmt.getQualifiedName()} is not present.")
