"""Expression dispatcher (dask_sql/physical/rex/convert.py:16-76): RexType -> plugin name."""
import logging

from ...utils import LoggableDataFrame, Pluggable

logger = logging.getLogger(__name__)

_REX_TYPE_TO_PLUGIN = {
    "RexType.Reference": "InputRef",
    "RexType.Call": "RexCall",
    "RexType.Literal": "RexLiteral",
    "RexType.Alias": "RexAlias",
    "RexType.ScalarSubquery": "ScalarSubquery",
}


class RexConverter(Pluggable):
    @classmethod
    def add_plugin_class(cls, plugin_class, replace=True):
        logger.debug(f"Registering REX plugin for {plugin_class.class_name}")
        cls.add_plugin(plugin_class.class_name, plugin_class(), replace=replace)

    @classmethod
    def convert(cls, rel, rex, dc, context):
        expr_type = _REX_TYPE_TO_PLUGIN[str(rex.getRexType())]
        try:
            plugin_instance = cls.get_plugin(expr_type)
        except KeyError:  # pragma: no cover
            raise NotImplementedError(f"No conversion for class {expr_type} available (yet).")
        logger.debug(f"Processing REX {rex} using {plugin_instance.__class__.__name__}...")
        df = plugin_instance.convert(rel, rex, dc, context=context)
        logger.debug(f"Processed REX {rex} into {LoggableDataFrame(df)}")
        return df
