"""Base class of expression plugins (dask_sql/physical/rex/base.py:16-34)."""


class BaseRexPlugin:
    """Converts one REX node into a lazy column or a python scalar."""

    class_name = None

    def convert(self, rel, rex, dc, context):
        raise NotImplementedError
