"""Base class of relational plugins: the drop-in boundary (dask_sql/physical/rel/base.py:16-124).

A plugin declares `class_name` (the LogicalPlan node type(s) it serves) and implements
convert(rel, context) -> DataContainer.  Plugins are singletons shared by all queries and must be
stateless; convert() only builds lazy device frames."""
from typing import Optional

from ...datacontainer import ColumnContainer, DataContainer
from ...mappings import cast_column_type, sql_to_python_type

_SEMI_JOINS = ("leftsemi", "leftanti")      # their row type still lists the right side's fields


class BaseRelPlugin:
    class_name = None

    def convert(self, rel, context) -> DataContainer:
        raise NotImplementedError

    @staticmethod
    def assert_inputs(rel, n: int = 1, context=None):
        """The node's inputs, converted depth-first; their number is part of the contract
        (base.py:66-86)."""
        from .convert import RelConverter

        children = rel.get_inputs()
        assert len(children) == n, f"{rel.get_current_node_type()} expects {n} input(s), got {len(children)}"
        return [RelConverter.convert(child, context) for child in children]

    @staticmethod
    def fix_column_to_row_type(cc: ColumnContainer, row_type, join_type: Optional[str] = None) -> ColumnContainer:
        """Rename the (already correctly ordered) columns to the plan's field names (base.py:31-51)."""
        wanted = [str(name) for name in row_type.getFieldNames()]
        if join_type in _SEMI_JOINS:
            wanted = wanted[: len(cc.columns)]
        return cc.rename_handle_duplicates(from_columns=cc.columns, to_columns=wanted).limit_to(wanted)

    @staticmethod
    def check_columns_from_row_type(df, row_type):
        assert list(df.columns) == [str(name) for name in row_type.getFieldNames()]

    @staticmethod
    def fix_dtype_to_row_type(dc: DataContainer, row_type, join_type: Optional[str] = None) -> DataContainer:
        """Cast a column only when its type FAMILY differs from the plan's (base.py:88-124): int64,
        Int64 and int32 count as the same family and are left alone (SUM(BIGINT) stays int64)."""
        frame, names = dc.df, dc.column_container
        fields = row_type.getFieldList()
        if join_type in _SEMI_JOINS:
            fields = fields[: len(names.columns)]
        for field in fields:
            try:
                target = sql_to_python_type(field.getType().getSqlType())
            except NotImplementedError:
                continue        # a type outside the hot path: leave the column as it is
            backend = names.get_backend_by_frontend_name(str(field.getQualifiedName()))
            if backend in frame.columns:
                frame = cast_column_type(frame, backend, target)
        return DataContainer(frame, names)
