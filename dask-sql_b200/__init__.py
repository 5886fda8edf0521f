"""dask_sql_b200 — B200-native execution layer for dask-sql's filter -> join -> group-by hot path.

Same public surface as the reference for this path: Context (create_table / sql / explain),
RelConverter / BaseRelPlugin, RexConverter / BaseRexPlugin, DataContainer / ColumnContainer.
"""
from . import _lib  # noqa: F401  (fails loudly if libb200sql.so is missing: there is no CPU fallback)
from . import config
from .context import Context
from .datacontainer import ColumnContainer, DataContainer, Statistics
from .physical.rel import RelConverter
from .physical.rel.base import BaseRelPlugin
from .physical.rex import RexConverter
from .physical.rex.base import BaseRexPlugin

__version__ = "0.1.0"
__all__ = ["Context", "config", "RelConverter", "BaseRelPlugin", "RexConverter", "BaseRexPlugin",
           "DataContainer", "ColumnContainer", "Statistics"]
