"""CrossJoin (dask_sql/physical/rel/logical/cross_join.py): not a hash-join hot path; refused
loudly rather than executed slowly."""
from ..base import BaseRelPlugin


class DaskCrossJoinPlugin(BaseRelPlugin):
    class_name = "CrossJoin"

    def convert(self, rel, context):
        raise NotImplementedError(
            "CROSS JOIN / joins without an equality key are outside the hash-join hot path of the B200 layer")
