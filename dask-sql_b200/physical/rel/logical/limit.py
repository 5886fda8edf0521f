"""LIMIT / OFFSET (dask_sql/physical/rel/logical/limit.py:18-113)."""
from ....datacontainer import DataContainer
from ...rex import RexConverter
from ..base import BaseRelPlugin


class DaskLimitPlugin(BaseRelPlugin):
    class_name = "Limit"

    def convert(self, rel, context) -> DataContainer:
        (dc,) = self.assert_inputs(rel, 1, context)
        df, cc = dc.df, dc.column_container

        def value(x):
            if x is None or isinstance(x, int):
                return x
            return RexConverter.convert(rel, x, df, context=context)   # RexType.Literal in DataFusion plans

        limit, offset = value(rel.limit().getFetch()), value(rel.limit().getSkip())
        df = df.limit(fetch=limit, offset=offset or 0)
        cc = self.fix_column_to_row_type(cc, rel.getRowType())
        return DataContainer(df, cc)
