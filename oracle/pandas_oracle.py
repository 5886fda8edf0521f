"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product (dask-sql_b200/).

CPU restatement of dask-sql's filter -> hash-join -> hash-groupby-aggregate path
(dask-contrib/dask-sql @ f186de3).  dask-sql itself has no arithmetic: its plugins emit
pandas calls that dask runs per partition.  This module makes exactly those pandas calls
(pandas IS the reference's arithmetic, pyproject.toml:32 `pandas>=1.4.0`, no exact pin) and
restates the dask layer around them (dask is not installable here):

    rows split into P contiguous partitions -> per-partition pandas op ->
      join   : the build side is broadcast to every partition (merge(broadcast=True))
      groupby: chunk per partition, then concat <= split_every partials and re-aggregate,
               tree-wise, until one remains (dd.Aggregation chunk/agg, aggregate.py:488-492)

Pinned against the reference's own known-answer tests in tests/test_oracle_golden.py
(fixtures tests/golden/reference_vectors.py, transcribed from /root/reference/tests) and
against sqlite3 (the reference's differential oracle, tests/integration/test_compatibility.py:25-47).
The real dask_sql cannot be imported in this image (needs the Rust planner crate and dask),
so there is no oracle/_ref; see DESIGN.md "Oracle".

Each function cites the reference lines it follows (paths relative to /root/reference).
"""
import operator
from concurrent.futures import ThreadPoolExecutor
from functools import reduce
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import pandas as pd


# ---------------------------------------------------------------------------------------------
# expressions — dask_sql/physical/rex/core/call.py (the operator table OPERATION_MAPPING :1047-1216)
# ---------------------------------------------------------------------------------------------
def rex_reduce(op: Callable, *operands):
    """ReduceOperation.reduce (call.py:140-162): n-ary operators fold left over their operands."""
    return reduce(op, operands) if len(operands) > 1 else op(*operands)


def rex_sql_div(lhs, rhs, output_is_float: bool):
    """SQLDivisionOperator.div (call.py:165-189): true division, truncated (NOT floored) when the
    SQL result type is an integer type."""
    result = lhs / rhs
    return result if output_is_float else np.trunc(result)


def rex_not(x):
    """NotOperation (call.py:348-364)."""
    return ~(x.astype("boolean")) if isinstance(x, pd.Series) else (not x)


def rex_is_null(x):
    """IsNullOperation (call.py:367-383)."""
    return x.isna() if isinstance(x, pd.Series) else bool(pd.isna(x))


def rex_is_true(x):
    """IsTrueOperation (call.py:315-332): NULL -> False."""
    return x.astype("boolean").fillna(False)


def rex_is_false(x):
    """IsFalseOperation (call.py:295-312): NULL -> False."""
    return ~x.astype("boolean").fillna(True)


REX_BINARY = {  # call.py:1050-1062
    "and": operator.and_, "or": operator.or_, ">": operator.gt, ">=": operator.ge, "<": operator.lt,
    "<=": operator.le, "=": operator.eq, "<>": operator.ne, "+": operator.add, "-": operator.sub,
    "*": operator.mul,
}


# ---------------------------------------------------------------------------------------------
# filter  — dask_sql/physical/rel/logical/filter.py:20-45, table_scan.py:80-119
# ---------------------------------------------------------------------------------------------
def filter_or_scalar(df: pd.DataFrame, cond) -> pd.DataFrame:
    """filter.py:20-45: scalar conditions short-circuit; NULL in a boolean mask is False."""
    if np.isscalar(cond):
        return df if cond else df.head(0)
    cond = cond.fillna(False)          # filter.py:39
    return df[cond.astype(bool)]       # filter.py:40


def apply_filters(df: pd.DataFrame, conds: Sequence[Callable[[pd.DataFrame], pd.Series]]) -> pd.DataFrame:
    """table_scan.py:93-110: pushed-down filters are AND-reduced, then applied once."""
    if not conds:
        return df
    cond = reduce(operator.and_, [c(df) for c in conds])
    return filter_or_scalar(df, cond)


# ---------------------------------------------------------------------------------------------
# join  — dask_sql/physical/rel/logical/join.py:189-248
# ---------------------------------------------------------------------------------------------
def join_on_columns(lhs: pd.DataFrame, rhs: pd.DataFrame, lhs_on: Sequence[str], rhs_on: Sequence[str],
                    how: str = "inner") -> pd.DataFrame:
    """join.py:189-248.  Column names of lhs and rhs must already be disjoint (the plugin renames
    them to lhs_i / rhs_i, join.py:65-72).  how in inner|left|right|outer|leftanti|leftsemi."""
    if how in ("inner", "right"):                                   # join.py:202-207
        lhs = lhs[reduce(operator.and_, [~lhs[c].isna() for c in lhs_on])]
    if how in ("inner", "left", "leftanti", "leftsemi"):            # join.py:208-213
        rhs = rhs[reduce(operator.and_, [~rhs[c].isna() for c in rhs_on])]
    added = [f"common_{i}" for i in range(len(lhs_on))]            # join.py:215-226
    lhs_t = lhs.assign(**{a: lhs[c] for a, c in zip(added, lhs_on)})
    rhs_t = rhs.assign(**{a: rhs[c] for a, c in zip(added, rhs_on)})
    if how == "leftanti":                                           # join.py:229-239
        df = lhs_t.merge(rhs_t, on=added, how="left", indicator=True).drop(columns=added)
        return df[df["_merge"] == "left_only"].drop(columns=["_merge"] + list(rhs.columns), errors="ignore")
    if how == "leftsemi":
        # join.py:78-79 degrades CPU leftsemi to inner; SQL semantics (each lhs row once) is what
        # DataFusion plans for IN/EXISTS, so restate it as an inner join on the distinct rhs keys
        keys = rhs_t[added].drop_duplicates()
        return lhs_t.merge(keys, on=added, how="inner").drop(columns=added)
    return lhs_t.merge(rhs_t, on=added, how=how).drop(columns=added)  # join.py:241-246


def broadcast_join(lhs_parts: List[pd.DataFrame], rhs: pd.DataFrame, lhs_on, rhs_on, how="inner",
                   workers: int = 1) -> List[pd.DataFrame]:
    """dask merge(broadcast=True): every probe partition is merged with the whole build side
    (join.py:241-246 with sql.join.broadcast, sql.yaml:9-10)."""
    f = lambda p: join_on_columns(p, rhs, lhs_on, rhs_on, how)
    if workers > 1:
        with ThreadPoolExecutor(workers) as ex:
            return list(ex.map(f, lhs_parts))
    return [f(p) for p in lhs_parts]


# ---------------------------------------------------------------------------------------------
# group-by aggregate — dask_sql/physical/rel/logical/aggregate.py:288-375, 486-495, 522-589
# ---------------------------------------------------------------------------------------------
_CHUNK = {
    "sum": lambda s: s.sum(min_count=1),      # aggregate.py:488-492 custom_sum chunk
    "count": lambda s: s.count(),
    "min": lambda s: s.min(),
    "max": lambda s: s.max(),
}
_COMBINE = {
    "sum": lambda s: s.sum(min_count=1),      # aggregate.py:491 custom_sum agg
    "count": lambda s: s.sum(),
    "min": lambda s: s.min(),
    "max": lambda s: s.max(),
}


def _groupby(df, group_cols):
    if group_cols:
        return df.groupby(by=list(group_cols), dropna=False)                # aggregate.py:575-577
    return df.assign(__const=1).groupby(by=["__const"], dropna=False)       # aggregate.py:305-306,576


def groupby_chunk(df: pd.DataFrame, group_cols: Sequence[str], aggs: Sequence[Tuple[Optional[str], str, str]]):
    """Per-partition partials.  aggs: (input column or None for COUNT(*), output name, fn) with
    fn in sum|count|mean|min|max|size.  mean is carried as (sum, count) like dask's mean."""
    g = _groupby(df, group_cols)
    out = {}
    for col, name, fn in aggs:
        if fn == "size" or col is None:
            out[name] = g.size()
        elif fn == "mean":
            out[name + "__sum"] = g[col].sum()
            out[name + "__count"] = g[col].count()
        else:
            out[name] = _CHUNK[fn](g[col])
    return pd.DataFrame(out) if out else g.size().to_frame("__size")


def groupby_combine(partials: List[pd.DataFrame], group_cols, aggs) -> pd.DataFrame:
    """Concat <= split_every partial frames and re-aggregate (dask tree reduction step)."""
    df = pd.concat(partials)
    levels = list(range(df.index.nlevels))
    g = df.groupby(level=levels, dropna=False)
    out = {}
    for col, name, fn in aggs:
        if fn == "size" or col is None:
            out[name] = g[name].sum()
        elif fn == "mean":
            out[name + "__sum"] = g[name + "__sum"].sum()
            out[name + "__count"] = g[name + "__count"].sum()
        else:
            out[name] = _COMBINE[fn](g[name])
    if not out:
        out["__size"] = g["__size"].sum()
    return pd.DataFrame(out)


def groupby_finalize(df: pd.DataFrame, group_cols, aggs) -> pd.DataFrame:
    out = {}
    for col, name, fn in aggs:
        if fn == "mean" and col is not None:
            out[name] = df[name + "__sum"] / df[name + "__count"]
        else:
            out[name] = df[name]
    res = pd.DataFrame(out, index=df.index)
    res = res.reset_index(drop=not group_cols)                                # aggregate.py:269
    if group_cols:
        res.columns = list(group_cols) + [name for _, name, _ in aggs]
    return res


def groupby_agg(parts: List[pd.DataFrame], group_cols: Sequence[str], aggs, split_every: int = 8,
                workers: int = 1) -> pd.DataFrame:
    """groupby(dropna=False).agg(..., split_every) over partitions (aggregate.py:575-581)."""
    f = lambda p: groupby_chunk(p, group_cols, aggs)
    if workers > 1:
        with ThreadPoolExecutor(workers) as ex:
            partials = list(ex.map(f, parts))
    else:
        partials = [f(p) for p in parts]
    while len(partials) > 1:
        partials = [groupby_combine(partials[i:i + split_every], group_cols, aggs)
                    for i in range(0, len(partials), split_every)]
    if len(partials) == 1 and len(parts) == 1:
        pass
    return groupby_finalize(partials[0], group_cols, aggs)


def split(df: pd.DataFrame, npartitions: int) -> List[pd.DataFrame]:
    """dd.from_pandas(df, npartitions) row-range partitioning (pandaslike.py:38)."""
    n = len(df)
    step = max(1, -(-n // max(1, npartitions)))
    return [df.iloc[i:i + step] for i in range(0, max(n, 1), step)]


# ---------------------------------------------------------------------------------------------
# the BASELINE.json configurations, as the reference would execute them
# ---------------------------------------------------------------------------------------------
def c1_filter_sum(parts: List[pd.DataFrame], workers=1):
    """SELECT SUM(x) FROM t WHERE x > 0  (plan: Aggregate <- TableScan full_filters=[x > 0])."""
    filt = [apply_filters(p, [lambda d: d["x"] > 0])[["x"]] for p in parts]
    return groupby_agg(filt, [], [("x", "SUM(t.x)", "sum")], workers=workers)


def c2_groupby_sum(parts: List[pd.DataFrame], workers=1, split_every=8):
    """SELECT key, SUM(val) FROM t GROUP BY key."""
    return groupby_agg(parts, ["key"], [("val", "SUM(t.val)", "sum")], split_every, workers)


def c3_join(fact_parts: List[pd.DataFrame], dim: pd.DataFrame, workers=1):
    """SELECT f.fk, f.v, d.w FROM fact f JOIN dim d ON f.fk = d.pk (broadcast build side)."""
    outs = broadcast_join(fact_parts, dim, ["fk"], ["pk"], "inner", workers)
    return [o[["fk", "v", "w"]] for o in outs]


def c4_q3(fact_parts: List[pd.DataFrame], dim: pd.DataFrame, workers=1, split_every=8):
    """SELECT d.grp, SUM(f.val) AS rev FROM fact f JOIN dim d ON f.fk = d.pk
       WHERE f.x > 0 AND d.flag < 5 GROUP BY d.grp
    Optimised plan shape: filters pushed into both TableScans, inner join,
    aggregate."""
    dim_f = apply_filters(dim, [lambda d: d["flag"] < 5])[["pk", "grp"]]

    def one(p):
        f = apply_filters(p, [lambda d: d["x"] > 0])[["fk", "val"]]
        j = join_on_columns(f, dim_f, ["fk"], ["pk"], "inner")
        return groupby_chunk(j, ["grp"], [("val", "rev", "sum")])

    if workers > 1:
        with ThreadPoolExecutor(workers) as ex:
            partials = list(ex.map(one, fact_parts))
    else:
        partials = [one(p) for p in fact_parts]
    aggs = [("val", "rev", "sum")]
    while len(partials) > 1:
        partials = [groupby_combine(partials[i:i + split_every], ["grp"], aggs)
                    for i in range(0, len(partials), split_every)]
    return groupby_finalize(partials[0], ["grp"], aggs)


def c5_groupby_sum_avg(parts: List[pd.DataFrame], workers=1, split_every=8):
    """SELECT key, SUM(val), AVG(val) FROM t GROUP BY key."""
    return groupby_agg(parts, ["key"], [("val", "SUM(t.val)", "sum"), ("val", "AVG(t.val)", "mean")],
                       split_every, workers)
