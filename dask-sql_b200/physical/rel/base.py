"""Base class of relational plugins: the drop-in boundary (dask_sql/physical/rel/base.py:16-124).

A plugin declares `class_name` (LogicalPlan node type it serves) and implements
convert(rel, context) -> DataContainer.  Plugins are singletons shared by all queries and must be
stateless; convert() only builds lazy device frames."""
import logging
from typing import Optional

from ...datacontainer import ColumnContainer, DataContainer
from ...mappings import cast_column_type, sql_to_python_type

logger = logging.getLogger(__name__)


class BaseRelPlugin:
    class_name = None

    def convert(self, rel, context) -> DataContainer:
        raise NotImplementedError

    @staticmethod
    def fix_column_to_row_type(cc: ColumnContainer, row_type, join_type: Optional[str] = None) -> ColumnContainer:
        """Blindly rename the columns (already in the right order) to the row type's field names
        (base.py:31-51)."""
        field_names = [str(x) for x in row_type.getFieldNames()]
        if join_type in ("leftsemi", "leftanti"):
            field_names = field_names[: len(cc.columns)]
        logger.debug(f"Renaming {cc.columns} to {field_names}")
        cc = cc.rename_handle_duplicates(from_columns=cc.columns, to_columns=field_names)
        return cc.limit_to(field_names)

    @staticmethod
    def check_columns_from_row_type(df, row_type):
        assert list(df.columns) == [str(x) for x in row_type.getFieldNames()]

    @staticmethod
    def assert_inputs(rel, n: int = 1, context=None):
        """Convert the node's n inputs recursively (base.py:66-86)."""
        input_rels = rel.get_inputs()
        assert len(input_rels) == n
        from .convert import RelConverter
        return [RelConverter.convert(input_rel, context) for input_rel in input_rels]

    @staticmethod
    def fix_dtype_to_row_type(dc: DataContainer, row_type, join_type: Optional[str] = None) -> DataContainer:
        """Cast columns whose type FAMILY differs from the plan's row type (base.py:88-124);
        int64 vs Int64 vs int32 are 'similar' and left alone."""
        df, cc = dc.df, dc.column_container
        field_list = row_type.getFieldList()
        if join_type in ("leftsemi", "leftanti"):
            field_list = field_list[: len(cc.columns)]
        for field in field_list:
            sql_type = field.getType().getSqlType()
            try:
                expected_type = sql_to_python_type(sql_type)
            except NotImplementedError:
                continue
            df_field_name = cc.get_backend_by_frontend_name(str(field.getQualifiedName()))
            if df_field_name in df.columns:
                df = cast_column_type(df, df_field_name, expected_type)
        return DataContainer(df, dc.column_container)
