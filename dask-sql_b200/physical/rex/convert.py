"""Expression dispatcher: PyExpr.getRexType() -> registered plugin
(dask_sql/physical/rex/convert.py:16-76)."""
import logging

from ...utils import LoggableDataFrame, Pluggable, PluginDispatch

log = logging.getLogger(__name__)

# plugin names by expression kind (rex/convert.py:16-22)
PLUGIN_OF_REX_TYPE = dict(Reference="InputRef", Call="RexCall", Literal="RexLiteral", Alias="RexAlias",
                          ScalarSubquery="ScalarSubquery")


class RexConverter(PluginDispatch, Pluggable):
    kind = "expression"

    @classmethod
    def convert(cls, rel, rex, dc, context):
        rex_type = str(rex.getRexType()).rpartition(".")[2]          # "RexType.Call" -> "Call"
        plugin = cls.plugin_for(PLUGIN_OF_REX_TYPE[rex_type])
        value = plugin.convert(rel, rex, dc, context=context)
        if log.isEnabledFor(logging.DEBUG):
            log.debug("%s -> %s via %s", rex, LoggableDataFrame(value), type(plugin).__name__)
        return value
