"""Aggregate / Distinct: GROUP BY, global aggregates, DISTINCT
(dask_sql/physical/rel/logical/aggregate.py:91-589).

The reference groups once per (filter column, distinct column) bucket with
groupby(by, dropna=False).agg({in: {out: fn}}, split_out, split_every) (aggregate.py:522-589),
using a constant column as key when there is no GROUP BY (aggregate.py:305-306).  Here each
bucket becomes one AggSource; at compute time it runs as b2_scan_agg (no keys) or one of the
b2_groupby_* kernels fused with the pending predicate, or as the fused star pipeline."""
import logging
from collections import defaultdict

from .... import config as dask_config
from ....datacontainer import ColumnContainer, DataContainer
from ....frame import AggSource, LazyFrame
from ....utils import new_temporary_column
from ...rex.convert import RexConverter
from ..base import BaseRelPlugin

logger = logging.getLogger(__name__)


class DaskAggregatePlugin(BaseRelPlugin):
    class_name = ["Aggregate", "Distinct"]

    # SQL aggregate name -> accumulator recipe of the group-by kernels (aggregate.py:117-231,
    # hot-path rows; "sum" is sum(min_count=1), aggregate.py:486-493)
    AGGREGATION_MAPPING = {
        "sum": "sum",
        "$sum0": "sum",
        "avg": "mean",
        "mean": "mean",
        "count": "count",
        "min": "min",
        "max": "max",
        # (count, sum, sum of squares) recipes of aggregate.py:129-231
        "stddev": "stddev_samp", "stddev_samp": "stddev_samp", "stddevsamp": "stddev_samp",
        "stddev_pop": "stddev_pop", "stddevpop": "stddev_pop",
        "variance": "var_samp", "var_samp": "var_samp", "var": "var_samp",
        "var_pop": "var_pop", "variance_pop": "var_pop", "variancepop": "var_pop",
    }
    _MOMENT_FUNCS = ("stddev_samp", "stddev_pop", "var_samp", "var_pop")

    def convert(self, rel, context) -> DataContainer:
        (dc,) = self.assert_inputs(rel, 1, context)
        agg = rel.aggregate()
        df = dc.df
        cc = dc.column_container.make_unique()
        group_exprs = agg.getGroupSets()
        group_columns = (agg.getDistinctColumns() if agg.isDistinctNode()
                         else [group_expr.column_name(rel) for group_expr in group_exprs])
        dc = DataContainer(df, cc)
        if not group_columns:
            logger.debug("Performing full-table aggregation")
        df_agg, output_column_order, cc = self._do_aggregations(rel, dc, group_columns, context)

        def try_get_backend_by_frontend_name(oc):
            try:
                return cc.get_backend_by_frontend_name(oc)
            except KeyError:
                return oc

        backend_output_column_order = [try_get_backend_by_frontend_name(oc) for oc in output_column_order]
        cc = ColumnContainer(df_agg.columns).limit_to(backend_output_column_order)
        cc = self.fix_column_to_row_type(cc, rel.getRowType())
        dc = DataContainer(df_agg, cc)
        return self.fix_dtype_to_row_type(dc, rel.getRowType())

    def _do_aggregations(self, rel, dc, group_columns, context):
        df, cc = dc.df, dc.column_container
        output_column_order = group_columns.copy()
        collected_aggregations, output_column_order, df, cc = self._collect_aggregations(
            rel, df, cc, context, output_column_order)
        groupby_agg_options = dask_config.get("sql.aggregate") or {}
        backend_groups = [cc.get_backend_by_frontend_name(g) for g in group_columns]
        if not collected_aggregations:
            # DISTINCT / GROUP BY without aggregates (aggregate.py:323-332)
            return df[backend_groups].drop_duplicates(**groupby_agg_options), output_column_order, cc

        # the unfiltered bucket first so no group is lost (aggregate.py:334-350)
        df_result = None
        key = (None, None)
        if key in collected_aggregations:
            df_result = self._perform_aggregation(DataContainer(df, cc), None, None,
                                                  collected_aggregations.pop(key), group_columns,
                                                  groupby_agg_options)
        for (filter_column, distinct_column), aggregations in collected_aggregations.items():
            agg_result = self._perform_aggregation(DataContainer(df, cc), filter_column, distinct_column,
                                                   aggregations, group_columns, groupby_agg_options)
            if df_result is None:
                df_result = agg_result
            else:
                # FILTER buckets join the main result on the group keys (the reference assigns by
                # index alignment, aggregate.py:371-373)
                extra = [c for c in agg_result.columns if c not in backend_groups]
                if backend_groups:
                    renamed = agg_result.rename({g: f"__rhs_{g}" for g in backend_groups})
                    merged = df_result.merge(renamed, left_on=backend_groups,
                                             right_on=[f"__rhs_{g}" for g in backend_groups], how="left")
                    df_result = merged[list(df_result.columns) + extra]
                else:
                    raise NotImplementedError("global aggregates with different FILTER clauses")
        return df_result, output_column_order, cc

    def _collect_aggregations(self, rel, df, cc, context, output_column_order):
        """Bucket aggregate calls by (filter column, distinct column) (aggregate.py:377-520)."""
        dc = DataContainer(df, cc)
        agg = rel.aggregate()
        input_rel = rel.get_inputs()[0]
        collected_aggregations = defaultdict(list)
        new_columns = {}
        for expr in agg.getNamedAggCalls():
            assert expr.getExprType() in {"Alias", "AggregateFunction", "AggregateUDF"}, \
                "Do not know how to handle this case!"
            for input_expr in agg.getArgs(expr):
                input_col = input_expr.column_name(input_rel)
                if input_col not in cc._frontend_backend_mapping:
                    random_name = new_temporary_column(df)
                    new_columns[random_name] = RexConverter.convert(input_rel, input_expr, dc, context=context)
                    cc = cc.add(input_col, random_name)
            filter_expr = expr.getFilterExpr()
            if filter_expr is not None:
                filter_col = filter_expr.column_name(input_rel)
                if filter_col not in cc._frontend_backend_mapping:
                    random_name = new_temporary_column(df)
                    new_columns[random_name] = RexConverter.convert(input_rel, filter_expr, dc, context=context)
                    cc = cc.add(filter_col, random_name)
        if new_columns:
            df = df.assign(**new_columns)

        for expr in agg.getNamedAggCalls():
            aggregation_name = agg.getAggregationFuncName(expr).lower()
            inputs = agg.getArgs(expr)
            if len(inputs) == 1:
                input_col = inputs[0].column_name(input_rel)
            elif len(inputs) == 0:
                input_col = None                      # COUNT(*) counts rows
            else:
                raise NotImplementedError("Can not cope with more than one input")
            filter_expr = expr.getFilterExpr()
            filter_backend_col = (cc.get_backend_by_frontend_name(filter_expr.column_name(input_rel))
                                  if filter_expr is not None else None)
            try:
                aggregation_function = self.AGGREGATION_MAPPING[aggregation_name]
            except KeyError:
                raise NotImplementedError(f"Aggregation function {aggregation_name} not implemented (yet).")
            if input_col is None:
                aggregation_function = "size"
            backend_name = cc.get_backend_by_frontend_name(input_col) if input_col is not None else None
            if expr.isDistinctAgg() and backend_name is None:
                raise NotImplementedError("COUNT(DISTINCT *)")
            output_col = expr.toString()
            if filter_backend_col is not None and not expr.isDistinctAgg():
                # agg(x) FILTER (WHERE f)  ==  agg(CASE WHEN f THEN x END): NULLs are skipped by every
                # aggregate, so the filtered aggregate shares the single fused pass of the unfiltered
                # ones instead of the reference's extra groupby per filter bucket (aggregate.py:352-373)
                cond = df[filter_backend_col]
                if backend_name is None:
                    masked, aggregation_function = cond.where(cond), "count"
                else:
                    masked = df[backend_name].where(cond)
                input_col = new_temporary_column(df)
                df = df.assign(**{input_col: masked})
                filter_backend_col = None
            collected_aggregations[(filter_backend_col, backend_name if expr.isDistinctAgg() else None)].append(
                (input_col, output_col, aggregation_function))
            output_column_order.append(output_col)
        return collected_aggregations, output_column_order, df, cc

    def _perform_aggregation(self, dc, filter_column, distinct_column, aggregations, group_columns,
                             groupby_agg_options):
        tmp_df = dc.df
        cc = dc.column_container
        group_columns = [cc.get_backend_by_frontend_name(g) for g in group_columns]
        if filter_column:
            tmp_df = tmp_df[tmp_df[filter_column].fillna(False)]
            logger.debug(f"Filtered by {filter_column} before aggregation.")
        if distinct_column:
            tmp_df = tmp_df.drop_duplicates(subset=(group_columns + [distinct_column]), **groupby_agg_options)
            logger.debug(f"Dropped duplicates from {distinct_column} before aggregation.")
        spec, moments = [], []
        for input_col, output_col, aggregation_f in aggregations:
            backend_in = cc.get_backend_by_frontend_name(input_col) if input_col is not None else None
            if aggregation_f in self._MOMENT_FUNCS:
                # STDDEV / VARIANCE from (count, sum, sum of squares), all accumulated in the same
                # fused pass as the other aggregates (aggregate.py:129-231 uses the same three moments)
                x = tmp_df[backend_in].astype("float64")
                x_name, sq_name = new_temporary_column(tmp_df), new_temporary_column(tmp_df)
                tmp_df = tmp_df.assign(**{x_name: x, sq_name: x * x})
                s, s2, n = (f"{output_col}__{k}" for k in ("s", "s2", "n"))
                spec += [(x_name, s, "sum"), (sq_name, s2, "sum"), (x_name, n, "count")]
                moments.append((output_col, aggregation_f, s, s2, n))
            else:
                spec.append((backend_in, output_col, aggregation_f))
        logger.debug(f"Performing aggregation {spec}")
        frame = LazyFrame(AggSource(tmp_df, group_columns, spec, groupby_agg_options))
        if moments:
            new = {}
            for output_col, f, s, s2, n in moments:
                S, S2, N = frame[s], frame[s2], frame[n]
                var = (S2 / N - (S / N) * (S / N)) if f.endswith("pop") else (S2 - S * S / N) / (N - 1)
                new[output_col] = var.sqrt() if f.startswith("stddev") else var
            frame = frame.assign(**new)[group_columns + [out for _, out, _ in aggregations]]
        return frame
