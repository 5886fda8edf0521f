"""Plan-node dispatcher: LogicalPlan.get_current_node_type() -> registered plugin
(the seam of dask_sql/physical/rel/convert.py:16-63; registering with replace=True is how a
third-party execution layer takes over a node type)."""
import logging

from ...utils import LoggableDataFrame, Pluggable, PluginDispatch

log = logging.getLogger(__name__)


class RelConverter(PluginDispatch, Pluggable):
    kind = "relational"

    @classmethod
    def convert(cls, rel, context):
        plugin = cls.plugin_for(rel.get_current_node_type())
        result = plugin.convert(rel, context=context)
        if log.isEnabledFor(logging.DEBUG):
            log.debug("%s -> %s via %s", rel, LoggableDataFrame(result), type(plugin).__name__)
        return result
