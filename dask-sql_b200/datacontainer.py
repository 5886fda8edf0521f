"""Values crossing the plugin boundary.  The public surface is the reference's
(dask_sql/datacontainer.py:19-231 -- a plugin written against dask-sql finds the same classes,
methods and behaviour); the representation behind it is this repo's own.

A ColumnContainer is an immutable view description: `_order` is the tuple of SQL-facing
("frontend") names in output order, `_backing` maps every name that was ever visible -- not only
the ones currently in `_order` -- to the column the frame really has ("backend").  Every operation
returns a new container built by `_derive`; nothing is ever mutated in place, so containers can be
shared between the lazy frames of a plan without defensive copies.
"""
from typing import Dict, Iterable, List, Optional, Tuple, Union

ColumnType = Union[str, int]


class ColumnContainer:
    __slots__ = ("_order", "_backing")

    def __init__(self, frontend_columns: List[str], frontend_backend_mapping: Optional[Dict[str, ColumnType]] = None):
        names = tuple(frontend_columns)
        if not all(isinstance(n, str) for n in names):
            raise AssertionError("All frontend columns need to be of string type")
        self._order: Tuple[str, ...] = names
        self._backing: Dict[str, ColumnType] = (dict(frontend_backend_mapping) if frontend_backend_mapping is not None
                                                else dict(zip(names, names)))

    # -- construction of derived containers -----------------------------------------------------
    def _derive(self, order: Iterable[str], aliases: Iterable[Tuple[str, ColumnType]] = ()) -> "ColumnContainer":
        """New container with `order` visible and `aliases` (name -> backend) added to the known names."""
        out = object.__new__(ColumnContainer)
        out._order = tuple(str(n) for n in order)
        out._backing = {**self._backing, **{str(k): v for k, v in aliases}}
        return out

    def knows(self, name: str) -> bool:
        """Whether `name` was ever a frontend name of this view (visible now or not)."""
        return name in self._backing

    # -- the reference's methods -----------------------------------------------------------------
    @property
    def columns(self) -> List[str]:
        return list(self._order)

    def mapping(self) -> List[Tuple[str, ColumnType]]:
        return list(self._backing.items())

    def limit_to(self, fields: List[str]) -> "ColumnContainer":
        """Show only `fields`, in that order; an empty list means "no restriction" (datacontainer.py:53-65)."""
        if not fields:
            return self
        unknown = [f for f in fields if f not in self._backing]
        if unknown:
            raise AssertionError(f"unknown columns {unknown}")
        return self._derive(fields)

    def rename(self, columns: Dict[str, str]) -> "ColumnContainer":
        """Frontend renames, positions kept; old names stay resolvable (datacontainer.py:67-85)."""
        return self.rename_handle_duplicates(list(columns), list(columns.values()))

    def rename_handle_duplicates(self, from_columns: List[str], to_columns: List[str]) -> "ColumnContainer":
        """rename() given as two parallel lists; if a source name occurs twice its LAST target labels the
        visible column while every target becomes resolvable (datacontainer.py:87-107)."""
        pairs = [(str(s), str(d)) for s, d in zip(from_columns, to_columns)]
        label = dict(pairs)
        return self._derive((label.get(n, n) for n in self._order), ((d, self._backing[s]) for s, d in pairs))

    def add(self, frontend_column: str, backend_column: Optional[str] = None) -> "ColumnContainer":
        """Make `frontend_column` resolve to `backend_column` (default: itself); appended if not visible yet."""
        name = str(frontend_column)
        order = self._order if name in self._order else self._order + (name,)
        return self._derive(order, [(name, str(backend_column or name))])

    def get_backend_by_frontend_index(self, index: int) -> ColumnType:
        return self._backing[self._order[index]]

    def get_backend_by_frontend_name(self, column: str) -> ColumnType:
        """Backend of a frontend name; a name this view never knew is passed through unchanged."""
        return self._backing.get(column, column)

    def make_unique(self, prefix="col") -> "ColumnContainer":
        """Relabel the visible columns <prefix>_0, <prefix>_1, ... (datacontainer.py:161-171)."""
        return self.rename({name: f"{prefix}_{i}" for i, name in enumerate(self._order)})


class Statistics:
    """Row count used by the cost heuristics (datacontainer.py:174-187)."""

    def __init__(self, row_count) -> None:
        self.row_count = row_count

    def __eq__(self, other):
        return isinstance(other, Statistics) and self.row_count == other.row_count

    def __repr__(self):
        return f"Statistics(row_count={self.row_count})"


class DataContainer:
    """A lazy frame together with the ColumnContainer describing how SQL sees it (datacontainer.py:190-231)."""

    def __init__(self, df, column_container: ColumnContainer, statistics: Statistics = None, filepath: str = None):
        self.df = df
        self.column_container = column_container
        self.statistics = statistics
        self.filepath = filepath

    def assign(self):
        """The frame restricted to the visible columns and labelled with their frontend names."""
        cc = self.column_container
        visible = cc.columns
        frame = self.df[[cc.get_backend_by_frontend_name(n) for n in visible]]
        frame.columns = visible
        return frame


class SchemaContainer:
    """What one SQL schema holds (datacontainer.py:281-290): the attribute names are the ones the
    reference's Context and custom plugins read."""

    _REGISTRIES = ("tables", "statistics", "experiments", "models", "functions", "filepaths")

    def __init__(self, name: str):
        self.__name__ = name
        for registry in self._REGISTRIES:
            setattr(self, registry, {})
        self.function_lists: List[object] = []
