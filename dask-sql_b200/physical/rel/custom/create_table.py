"""CREATE TABLE <name> WITH (location = ..., format = ..., persist = ..., ...)
(dask_sql/physical/rel/custom/create_table.py:15-88): register a table from a storage location.
The reference hands the location to dask.dataframe.read_<format>; here Context.create_table reads
Parquet / CSV with pyarrow and lays the column chunks out in HBM (or pinned host memory)."""
from ..base import BaseRelPlugin
from ._target import may_create


class CreateTablePlugin(BaseRelPlugin):
    class_name = "CreateTable"

    def convert(self, rel, context):
        stmt = rel.create_table()
        schema, table = stmt.getSchemaName() or context.schema_name, stmt.getTableName()
        if not may_create(context, schema, table, stmt.getIfNotExists(), stmt.getOrReplace()):
            return
        options = dict(stmt.getSQLWithOptions())
        if "location" not in options:
            raise AttributeError("Parameters must include a 'location' parameter.")
        location = options.pop("location")
        fmt = options.pop("format", None)
        context.create_table(table, location, format=fmt.lower() if fmt else None,
                             persist=options.pop("persist", False), schema_name=schema,
                             gpu=options.pop("gpu", False), **options)
