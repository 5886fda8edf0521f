"""WHERE / HAVING not pushed into a scan (dask_sql/physical/rel/logical/filter.py:20-74)."""
import logging

import numpy as np

from ....datacontainer import DataContainer
from ...rex import RexConverter
from ..base import BaseRelPlugin

logger = logging.getLogger(__name__)


def filter_or_scalar(df, filter_condition, add_filters=None):
    """A scalar condition keeps everything or nothing; NULL in a boolean condition is False
    (filter.py:20-45).  The lazy frame records the conjuncts; they are evaluated inside the
    consuming kernel (b2_scan_t terms), never as a separate mask pass unless they are not
    `column <cmp> literal` shaped."""
    if filter_condition is None:
        return df.head(0, compute=False)
    if np.isscalar(filter_condition):
        if not filter_condition:
            logger.warning("Join condition is always false - returning empty dataset")
            return df.head(0, compute=False)
        return df
    filter_condition = filter_condition.fillna(False)
    return df[filter_condition]


class DaskFilterPlugin(BaseRelPlugin):
    class_name = "Filter"

    def convert(self, rel, context) -> DataContainer:
        (dc,) = self.assert_inputs(rel, 1, context)
        df, cc = dc.df, dc.column_container
        condition = rel.filter().getCondition()
        df_condition = RexConverter.convert(rel, condition, dc, context=context)
        df = filter_or_scalar(df, df_condition)
        cc = self.fix_column_to_row_type(cc, rel.getRowType())
        return DataContainer(df, cc)
