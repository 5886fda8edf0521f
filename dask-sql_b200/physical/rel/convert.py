"""Relational dispatcher (dask_sql/physical/rel/convert.py:16-63): node type -> plugin."""
import logging

from ...utils import LoggableDataFrame, Pluggable

logger = logging.getLogger(__name__)


class RelConverter(Pluggable):
    @classmethod
    def add_plugin_class(cls, plugin_class, replace=True):
        logger.debug(f"Registering REL plugin for {plugin_class.class_name}")
        cls.add_plugin(plugin_class.class_name, plugin_class(), replace=replace)

    @classmethod
    def convert(cls, rel, context):
        node_type = rel.get_current_node_type()
        try:
            plugin_instance = cls.get_plugin(node_type)
        except KeyError:  # pragma: no cover
            raise NotImplementedError(f"No relational conversion for node type {node_type} available (yet).")
        logger.debug(f"Processing REL {rel} using {plugin_instance.__class__.__name__}...")
        df = plugin_instance.convert(rel, context=context)
        logger.debug(f"Processed REL {rel} into {LoggableDataFrame(df)}")
        return df
