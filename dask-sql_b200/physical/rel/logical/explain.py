"""EXPLAIN: return the plan text (dask_sql/physical/rel/logical/explain.py)."""
from ..base import BaseRelPlugin


class ExplainPlugin(BaseRelPlugin):
    class_name = "Explain"

    def convert(self, rel, context):
        return "\n".join(rel.explain_node().getExplainString())
