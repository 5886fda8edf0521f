"""Expression plugins of the hot path: calls (operators), column references, literals, aliases."""
from . import alias, call, input_ref, literal

RexAliasPlugin = alias.RexAliasPlugin
RexCallPlugin = call.RexCallPlugin
RexInputRefPlugin = input_ref.RexInputRefPlugin
RexLiteralPlugin = literal.RexLiteralPlugin

ALL_PLUGINS = (RexCallPlugin, RexInputRefPlugin, RexLiteralPlugin, RexAliasPlugin)
__all__ = ["RexAliasPlugin", "RexCallPlugin", "RexInputRefPlugin", "RexLiteralPlugin", "ALL_PLUGINS"]
