"""DROP TABLE [IF EXISTS] <name> (dask_sql/physical/rel/custom/drop_table.py)."""
import logging

from ..base import BaseRelPlugin

logger = logging.getLogger(__name__)


class DropTablePlugin(BaseRelPlugin):
    class_name = "DropTable"

    def convert(self, rel, context):
        dt = rel.drop_table()
        qualified = dt.getQualifiedName()
        *schema_name, table_name = qualified.split(".")
        if len(schema_name) > 1:
            raise RuntimeError(f"Expected unqualified or fully qualified table name, got {qualified}.")
        schema_name = context.schema_name if not schema_name else schema_name[0]
        if schema_name not in context.schema or table_name.lower() not in context.schema[schema_name].tables:
            if not dt.getIfExists():
                raise RuntimeError(f"A table with the name {qualified} is not present.")
            return
        context.drop_table(table_name, schema_name=schema_name)
