from .aggregate import DaskAggregatePlugin
from .cross_join import DaskCrossJoinPlugin
from .explain import ExplainPlugin
from .filter import DaskFilterPlugin
from .join import DaskJoinPlugin
from .limit import DaskLimitPlugin
from .project import DaskProjectPlugin
from .sort import DaskSortPlugin
from .subquery_alias import SubqueryAlias
from .table_scan import DaskTableScanPlugin

__all__ = [DaskAggregatePlugin, DaskCrossJoinPlugin, ExplainPlugin, DaskFilterPlugin, DaskJoinPlugin,
           DaskProjectPlugin, SubqueryAlias, DaskTableScanPlugin, DaskSortPlugin, DaskLimitPlugin]
