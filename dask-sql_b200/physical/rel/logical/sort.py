"""ORDER BY (dask_sql/physical/rel/logical/sort.py:12-39; apply_sort, physical/utils/sort.py:9-140):
a 'next' row of the scope, the tail of the real TPC-H Q3."""
from ....datacontainer import DataContainer
from ....utils import new_temporary_column
from ...rex import RexConverter
from ..base import BaseRelPlugin


class DaskSortPlugin(BaseRelPlugin):
    class_name = "Sort"

    def convert(self, rel, context) -> DataContainer:
        (dc,) = self.assert_inputs(rel, 1, context)
        df, cc = dc.df, dc.column_container
        sort_expressions = rel.sort().getCollation()
        sort_columns, extra = [], {}
        for expr in sort_expressions:
            name = expr.column_name(rel)
            if cc.knows(name):
                sort_columns.append(cc.get_backend_by_frontend_name(name))
            else:  # ORDER BY <expression over output columns>
                tmp = new_temporary_column(df)
                extra[tmp] = RexConverter.convert(rel, expr.getSortExpr(), dc, context=context)
                sort_columns.append(tmp)
        if extra:
            df = df.assign(**extra)
        sort_ascending = [expr.isSortAscending() for expr in sort_expressions]
        sort_null_first = [expr.isSortNullsFirst() for expr in sort_expressions]
        df = df.sort_values(sort_columns, ascending=sort_ascending, nulls_first=sort_null_first)
        if extra:
            df = df.drop(columns=list(extra))
        cc = self.fix_column_to_row_type(cc, rel.getRowType())
        return DataContainer(df, cc)
