"""Projection: re-map references for free, add computed columns lazily
(dask_sql/physical/rel/logical/project.py:17-78)."""
import logging

from ....datacontainer import DataContainer
from ....planner import RexType
from ....utils import is_frame, new_temporary_column
from ...rex import RexConverter
from ..base import BaseRelPlugin

logger = logging.getLogger(__name__)


class DaskProjectPlugin(BaseRelPlugin):
    class_name = "Projection"

    def convert(self, rel, context) -> DataContainer:
        (dc,) = self.assert_inputs(rel, 1, context)
        df, cc = dc.df, dc.column_container
        named_projects = rel.projection().getNamedProjects()
        column_names, new_columns, new_mappings = [], {}, {}
        for key, expr in named_projects:
            key = str(key)
            column_names.append(key)
            if str(expr.getRexType()) == str(RexType.Reference):
                backend_column_name = cc.get_backend_by_frontend_index(expr.getIndex())
                logger.debug(f"Not re-adding the same column {key} (but just referencing it)")
                new_mappings[key] = backend_column_name
            else:
                random_name = new_temporary_column(df)
                value = RexConverter.convert(rel, expr, dc, context=context)
                new_columns[random_name] = value
                logger.debug(f"Adding a new column {key} out of {expr}")
                new_mappings[key] = random_name
        if new_columns:
            df = df.assign(**new_columns)
        for key, backend_column_name in new_mappings.items():
            cc = cc.add(key, backend_column_name)
        cc = cc.limit_to(column_names)
        cc = self.fix_column_to_row_type(cc, rel.getRowType())
        dc = DataContainer(df, cc)
        return self.fix_dtype_to_row_type(dc, rel.getRowType())
