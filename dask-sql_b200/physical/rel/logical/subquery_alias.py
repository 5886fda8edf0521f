"""SubqueryAlias: pass-through that re-qualifies column names
(dask_sql/physical/rel/logical/subquery_alias.py)."""
from ....datacontainer import DataContainer
from ..base import BaseRelPlugin


class SubqueryAlias(BaseRelPlugin):
    class_name = "SubqueryAlias"

    def convert(self, rel, context) -> DataContainer:
        (dc,) = self.assert_inputs(rel, 1, context)
        cc = self.fix_column_to_row_type(dc.column_container, rel.getRowType())
        return DataContainer(dc.df, cc)
