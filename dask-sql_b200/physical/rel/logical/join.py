"""Join: split the condition into equi-keys + residual, hash-join on the keys, filter the rest
(dask_sql/physical/rel/logical/join.py:23-322).

The reference drops NULL keys with two boolean-indexing copies and calls
lhs.merge(rhs, on=common_i, how, broadcast) (join.py:189-248).  Here the merge is recorded lazily;
at compute time it becomes b2_join_build(_dense) + b2_join_count/write (NULL keys are skipped
inside the kernels), or — when an Aggregate sits on top — the fused star pipeline."""
import logging
import operator
from functools import reduce

from .... import config as dask_config
from ....datacontainer import ColumnContainer, DataContainer
from ...rex import RexConverter
from ..base import BaseRelPlugin
from .filter import filter_or_scalar

logger = logging.getLogger(__name__)


class DaskJoinPlugin(BaseRelPlugin):
    class_name = "Join"

    JOIN_TYPE_MAPPING = {
        "INNER": "inner",
        "LEFT": "left",
        "RIGHT": "right",
        "FULL": "outer",
        "LEFTSEMI": "leftsemi",
        "LEFTANTI": "leftanti",
    }

    def convert(self, rel, context) -> DataContainer:
        join = rel.join()
        dc_lhs, dc_rhs = self.assert_inputs(rel, 2, context)
        # unique column names on both sides so SQL's positional references survive the merge
        cc_lhs_renamed = dc_lhs.column_container.make_unique("lhs")
        cc_rhs_renamed = dc_rhs.column_container.make_unique("rhs")
        df_lhs_renamed = DataContainer(dc_lhs.df, cc_lhs_renamed).assign()
        df_rhs_renamed = DataContainer(dc_rhs.df, cc_rhs_renamed).assign()

        join_type = self.JOIN_TYPE_MAPPING[str(join.getJoinType())]
        join_condition = join.getCondition()
        lhs_on, rhs_on, filter_condition = None, None, None
        if join_condition is not None:
            lhs_on, rhs_on, filter_condition = self._split_join_condition(join_condition)
            # indices refer to lhs|rhs side by side: make the rhs ones relative to the rhs frame
            rhs_on = [index - len(df_lhs_renamed.columns) for index in rhs_on]
            assert len(lhs_on) == len(rhs_on)
        if not lhs_on:
            raise NotImplementedError(
                "joins without an equality key (cross joins) are outside the hash-join hot path of the B200 layer")
        df = self._join_on_columns(df_lhs_renamed, df_rhs_renamed, lhs_on, rhs_on, join_type)

        if join_type in ("leftsemi", "leftanti"):
            correct_column_order = list(df_lhs_renamed.columns)
        else:
            correct_column_order = list(df_lhs_renamed.columns) + list(df_rhs_renamed.columns)
        cc = ColumnContainer(df.columns).limit_to(correct_column_order)
        row_type = rel.getRowType()
        field_specifications = [str(f) for f in row_type.getFieldNames()]
        if join_type in ("leftsemi", "leftanti"):
            field_specifications = field_specifications[: len(cc.columns)]
        cc = cc.rename({from_col: to_col for from_col, to_col in zip(cc.columns, field_specifications)})
        cc = self.fix_column_to_row_type(cc, row_type, join_type)
        dc = DataContainer(df, cc)

        if filter_condition:
            # residual (non-equi) part of the ON clause, applied on the join output (join.py:170-181)
            filter_condition = reduce(
                operator.and_,
                [RexConverter.convert(rel, rex, dc, context=context) for rex in filter_condition],
            )
            logger.debug(f"Additionally applying filter {filter_condition}")
            df = filter_or_scalar(df, filter_condition)
            dc = DataContainer(df, cc)
        return self.fix_dtype_to_row_type(dc, rel.getRowType(), join_type)

    def _join_on_columns(self, df_lhs_renamed, df_rhs_renamed, lhs_on, rhs_on, join_type):
        """NULL keys never match (join.py:198-213): the kernels skip them on the build side and
        find no partner for them on the probe side, which yields the same rows as dropping them."""
        lhs_keys = [df_lhs_renamed.columns[i] for i in lhs_on]
        rhs_keys = [df_rhs_renamed.columns[i] for i in rhs_on]
        broadcast = dask_config.get("sql.join.broadcast")
        return df_lhs_renamed.merge(df_rhs_renamed, left_on=lhs_keys, right_on=rhs_keys, how=join_type,
                                    broadcast=broadcast)

    def _split_join_condition(self, join_condition):
        if str(join_condition.getRexType()) in ["RexType.Literal", "RexType.Reference"]:
            return [], [], [join_condition]
        elif not str(join_condition.getRexType()) == "RexType.Call":
            raise NotImplementedError("Can not understand join condition.")
        lhs_on, rhs_on, filter_condition = [], [], []
        try:
            lhs_on, rhs_on, filter_condition_part = self._extract_lhs_rhs(join_condition)
            filter_condition.extend(filter_condition_part)
        except AssertionError:
            filter_condition.append(join_condition)
        if lhs_on and rhs_on:
            return lhs_on, rhs_on, filter_condition
        return [], [], [join_condition]

    def _extract_lhs_rhs(self, rex):
        assert str(rex.getRexType()) == "RexType.Call"
        operator_name = str(rex.getOperatorName())
        assert operator_name in ["=", "AND"]
        operands = rex.getOperands()
        assert len(operands) == 2
        if operator_name == "=":
            operand_lhs, operand_rhs = operands
            if (str(operand_lhs.getRexType()) == "RexType.Reference"
                    and str(operand_rhs.getRexType()) == "RexType.Reference"):
                lhs_index, rhs_index = operand_lhs.getIndex(), operand_rhs.getIndex()
                # the rhs table always comes after the lhs table
                if lhs_index > rhs_index:
                    lhs_index, rhs_index = rhs_index, lhs_index
                return [lhs_index], [rhs_index], []
            raise AssertionError("Invalid join condition")
        lhs_indices, rhs_indices, filter_conditions = [], [], []
        for operand in operands:
            try:
                lhs_index, rhs_index, filter_condition = self._extract_lhs_rhs(operand)
                filter_conditions.extend(filter_condition)
                lhs_indices.extend(lhs_index)
                rhs_indices.extend(rhs_index)
            except AssertionError:
                filter_conditions.append(operand)
        return lhs_indices, rhs_indices, filter_conditions
