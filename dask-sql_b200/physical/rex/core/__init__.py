from .alias import RexAliasPlugin
from .call import RexCallPlugin
from .input_ref import RexInputRefPlugin
from .literal import RexLiteralPlugin

__all__ = [RexAliasPlugin, RexCallPlugin, RexInputRefPlugin, RexLiteralPlugin]
