"""Alias wrapper (dask_sql/physical/rex/core/alias.py)."""
from ....datacontainer import DataContainer
from ..base import BaseRexPlugin
from ..convert import RexConverter


class RexAliasPlugin(BaseRexPlugin):
    class_name = "RexAlias"

    def convert(self, rel, rex, dc, context):
        operands = rex.getOperands()
        assert len(operands) == 1
        value = RexConverter.convert(rel, operands[0], dc, context=context)
        return value.df if isinstance(value, DataContainer) else value
