"""Column reference (dask_sql/physical/rex/core/input_ref.py:13-35)."""
from ..base import BaseRexPlugin


class RexInputRefPlugin(BaseRexPlugin):
    class_name = "InputRef"

    def convert(self, rel, rex, dc, context):
        backend_column_name = dc.column_container.get_backend_by_frontend_index(rex.getIndex())
        return dc.df[backend_column_name]
