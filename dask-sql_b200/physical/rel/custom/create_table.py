"""CREATE TABLE <name> WITH (location = ..., format = ..., persist = ..., ...)
(dask_sql/physical/rel/custom/create_table.py:15-88): register a table from a storage location.
The reference hands the location to dask.dataframe.read_<format>; here Context.create_table reads
Parquet / CSV with pyarrow and lays the column chunks out in HBM (or pinned host memory)."""
import logging

from ..base import BaseRelPlugin

logger = logging.getLogger(__name__)


class CreateTablePlugin(BaseRelPlugin):
    class_name = "CreateTable"

    def convert(self, rel, context):
        ct = rel.create_table()
        schema_name = ct.getSchemaName() or context.schema_name
        table_name = ct.getTableName()
        if table_name.lower() in context.schema[schema_name].tables:
            if ct.getIfNotExists():
                return
            elif not ct.getOrReplace():
                raise RuntimeError(f"A table with the name {table_name} is already present.")
        kwargs = dict(ct.getSQLWithOptions())
        logger.debug(f"Creating new table with name {table_name} and parameters {kwargs}")
        format = kwargs.pop("format", None)
        if format:
            format = format.lower()
        persist = kwargs.pop("persist", False)
        try:
            location = kwargs.pop("location")
        except KeyError:
            raise AttributeError("Parameters must include a 'location' parameter.")
        gpu = kwargs.pop("gpu", False)
        context.create_table(table_name, location, format=format, persist=persist, schema_name=schema_name,
                             gpu=gpu, **kwargs)
