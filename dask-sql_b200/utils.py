"""Plugin registry and helpers of the drop-in boundary.

Same public behaviour as dask_sql/utils.py:61-105 (Pluggable: one process-global registry keyed
by subclass; add_plugin(names, plugin, replace=True); get_plugin; get_plugins) so that plugins
written for the reference's RelConverter / RexConverter register unchanged.
"""
import uuid
from collections import defaultdict

from .frame import LazyFrame, LazySeries


class Pluggable:
    """Name -> plugin mapping per subclass (utils.py:61-91 in the reference)."""

    __plugins = defaultdict(dict)
    version = 0

    @classmethod
    def add_plugin(cls, names, plugin, replace=True):
        if isinstance(names, str):
            names = [names]
        registry = Pluggable.__plugins[cls]
        if not replace and all(n in registry for n in names):
            return
        for n in names:
            registry[n] = plugin
        Pluggable.version += 1   # invalidates plans cached by Context.sql

    @classmethod
    def get_plugin(cls, name):
        return Pluggable.__plugins[cls][name]

    @classmethod
    def get_plugins(cls):
        return list(Pluggable.__plugins[cls].values())


class PluginDispatch:
    """What RelConverter and RexConverter have in common: register a plugin class under its
    class_name(s), look the plugin up by a key derived from the plan object, fail with
    NotImplementedError when nobody registered for it (physical/rel/convert.py:32-63,
    physical/rex/convert.py:42-76).  Mixed into two separate Pluggable subclasses because the
    registry is keyed by subclass."""

    kind = "?"

    @classmethod
    def add_plugin_class(cls, plugin_class, replace=True):
        cls.add_plugin(plugin_class.class_name, plugin_class(), replace=replace)

    @classmethod
    def plugin_for(cls, key):
        try:
            return cls.get_plugin(key)
        except KeyError:
            raise NotImplementedError(f"no {cls.kind} plugin is registered for {key!r}") from None


class ParsingException(Exception):
    """SQL could not be parsed / validated (utils.py:94-105)."""

    def __init__(self, sql, validation_exception_string):
        super().__init__(str(validation_exception_string).strip())


class OptimizationException(Exception):
    """The optimizer failed; Context falls back to the unoptimised plan (context.py:858-864)."""

    def __init__(self, exception_string):
        super().__init__(str(exception_string).strip())


def is_frame(x) -> bool:
    """True for frame-like operands (utils.py:19): here the lazy device frames/series."""
    return isinstance(x, (LazyFrame, LazySeries))


def new_temporary_column(df) -> str:
    """A column name not present in df (utils.py:191)."""
    while True:
        name = str(uuid.uuid4())
        if name not in df.columns:
            return name


class LoggableDataFrame:
    """Cheap repr for debug logs (never triggers execution)."""

    def __init__(self, df):
        self.df = df

    def __str__(self):
        df = self.df
        if isinstance(df, LazyFrame):
            return f"LazyFrame(columns={df.columns})"
        if isinstance(df, LazySeries):
            return f"LazySeries(name={df.name})"
        return f"Literal: {df!r}"
