"""TableScan: fetch the registered table, apply pushed-down filters, project columns
(dask_sql/physical/rel/logical/table_scan.py:21-119)."""
import logging
import operator
from functools import reduce

from ....datacontainer import DataContainer
from ...rex import RexConverter
from ..base import BaseRelPlugin
from .filter import filter_or_scalar

logger = logging.getLogger(__name__)


class DaskTableScanPlugin(BaseRelPlugin):
    class_name = "TableScan"

    def convert(self, rel, context) -> DataContainer:
        self.assert_inputs(rel, 0)
        table_scan = rel.table_scan()
        dask_table = rel.getTable()
        schema_name, table_name = (n.lower() for n in context.fqn(dask_table))
        dc = context.schema[schema_name].tables[table_name]
        # filters first: their columns need not be projected (table_scan.py:51-52)
        dc = self._apply_filters(table_scan, rel, dc, context)
        dc = self._apply_projections(table_scan, dask_table, dc)
        cc = self.fix_column_to_row_type(dc.column_container, rel.getRowType())
        dc = DataContainer(dc.df, cc)
        return self.fix_dtype_to_row_type(dc, rel.getRowType())

    def _apply_projections(self, table_scan, dask_table, dc):
        df, cc = dc.df, dc.column_container
        if table_scan.containsProjections():
            field_specifications = list(map(cc.get_backend_by_frontend_name, table_scan.getTableScanProjects()))
            df = df[field_specifications]
        else:
            field_specifications = [str(f) for f in dask_table.getRowType().getFieldNames()]
            field_specifications = [f.split(".")[-1] for f in field_specifications]
        cc = cc.limit_to(field_specifications)
        return DataContainer(df, cc)

    def _apply_filters(self, table_scan, rel, dc, context):
        df, cc = dc.df, dc.column_container
        all_filters = table_scan.getFilters()
        if all_filters:
            df_condition = reduce(
                operator.and_,
                [RexConverter.convert(rel, rex, dc, context=context) for rex in all_filters],
            )
            df = filter_or_scalar(df, df_condition)
        return DataContainer(df, cc)
