"""Values crossing the plugin boundary (mirrors dask_sql/datacontainer.py:19-231).

ColumnContainer keeps the SQL-facing ("frontend") column names and order separate from the
names the frame really has ("backend"), so plugins re-map references without touching data.
"""
from typing import Dict, List, Optional, Union

ColumnType = Union[str, int]


class ColumnContainer:
    def __init__(self, frontend_columns: List[str], frontend_backend_mapping: Optional[Dict[str, ColumnType]] = None):
        assert all(isinstance(c, str) for c in frontend_columns), "All frontend columns need to be of string type"
        self._frontend_columns = list(frontend_columns)
        self._frontend_backend_mapping = ({c: c for c in self._frontend_columns}
                                          if frontend_backend_mapping is None else frontend_backend_mapping)

    def _copy(self) -> "ColumnContainer":
        return ColumnContainer(list(self._frontend_columns), dict(self._frontend_backend_mapping))

    def limit_to(self, fields: List[str]) -> "ColumnContainer":
        """Keep only `fields`, in that order (datacontainer.py:53-65)."""
        if not fields:
            return self
        assert all(f in self._frontend_backend_mapping for f in fields)
        cc = self._copy()
        cc._frontend_columns = [str(f) for f in fields]
        return cc

    def rename(self, columns: Dict[str, str]) -> "ColumnContainer":
        """Rename frontend columns; order preserved (datacontainer.py:67-85)."""
        cc = self._copy()
        for src, dst in columns.items():
            cc._frontend_backend_mapping[str(dst)] = self._frontend_backend_mapping[str(src)]
        cc._frontend_columns = [str(columns[c]) if c in columns else c for c in self._frontend_columns]
        return cc

    def rename_handle_duplicates(self, from_columns: List[str], to_columns: List[str]) -> "ColumnContainer":
        """rename() that tolerates duplicates in from_columns (datacontainer.py:87-107)."""
        cc = self._copy()
        for src, dst in zip(from_columns, to_columns):
            cc._frontend_backend_mapping[str(dst)] = self._frontend_backend_mapping[str(src)]
        mapping = dict(zip(from_columns, to_columns))
        cc._frontend_columns = [str(mapping.get(c, c)) for c in self._frontend_columns]
        return cc

    def mapping(self):
        return list(self._frontend_backend_mapping.items())

    @property
    def columns(self) -> List[str]:
        return list(self._frontend_columns)

    def add(self, frontend_column: str, backend_column: Optional[str] = None) -> "ColumnContainer":
        cc = self._copy()
        frontend_column = str(frontend_column)
        cc._frontend_backend_mapping[frontend_column] = str(backend_column or frontend_column)
        if frontend_column not in cc._frontend_columns:
            cc._frontend_columns.append(frontend_column)
        return cc

    def get_backend_by_frontend_index(self, index: int) -> str:
        return self._frontend_backend_mapping[self._frontend_columns[index]]

    def get_backend_by_frontend_name(self, column: str) -> str:
        try:
            return self._frontend_backend_mapping[column]
        except KeyError:
            return column

    def make_unique(self, prefix="col") -> "ColumnContainer":
        """<prefix>_<i> for every column (datacontainer.py:161-171)."""
        return self.rename({str(c): f"{prefix}_{i}" for i, c in enumerate(self.columns)})


class Statistics:
    """Row count used by the cost heuristics (datacontainer.py:174-187)."""

    def __init__(self, row_count) -> None:
        self.row_count = row_count

    def __eq__(self, other):
        return isinstance(other, Statistics) and self.row_count == other.row_count


class DataContainer:
    """A lazy frame plus its ColumnContainer (datacontainer.py:190-231)."""

    def __init__(self, df, column_container: ColumnContainer, statistics: Statistics = None, filepath: str = None):
        self.df = df
        self.column_container = column_container
        self.statistics = statistics
        self.filepath = filepath

    def assign(self):
        """Frame with exactly the frontend columns, frontend names (datacontainer.py:217-231)."""
        cc = self.column_container
        df = self.df[[cc._frontend_backend_mapping[c] for c in cc.columns]]
        df.columns = cc.columns
        return df


class SchemaContainer:
    """Per-schema registry of tables / statistics / functions (datacontainer.py:281-290)."""

    def __init__(self, name: str):
        self.__name__ = name
        self.tables: Dict[str, DataContainer] = {}
        self.statistics: Dict[str, Statistics] = {}
        self.experiments: Dict[str, object] = {}
        self.models: Dict[str, object] = {}
        self.functions: Dict[str, object] = {}
        self.function_lists: List[object] = []
        self.filepaths: Dict[str, str] = {}
