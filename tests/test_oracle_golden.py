"""Pins the oracle (oracle/pandas_oracle.py) to the reference's own known-answer tests
(tests/golden/reference_vectors.py) and to sqlite3, the reference's differential oracle
(tests/integration/test_compatibility.py:25-47).  CPU only."""
import sqlite3

import numpy as np
import pandas as pd
import pytest

from oracle import pandas_oracle as O
from tests.golden import reference_vectors as G


def _cmp(got, exp, float_cols=()):
    got = got.reset_index(drop=True)
    exp = exp.reset_index(drop=True)
    assert len(got) == len(exp)
    if len(got) == 0:
        return
    cols = list(got.columns)
    got = got.sort_values(cols, na_position="last").reset_index(drop=True)
    exp = exp.sort_values(list(exp.columns), na_position="last").reset_index(drop=True)
    for cg, ce in zip(got.columns, exp.columns):
        g = got[cg].to_numpy(dtype=float, na_value=np.nan)
        e = exp[ce].to_numpy(dtype=float, na_value=np.nan)
        if ce in float_cols:
            np.testing.assert_allclose(g, e, rtol=1e-12, equal_nan=True)
        else:
            np.testing.assert_array_equal(g, e)


CASE = {c["name"]: c for c in G.CASES}


def test_filter_cases():
    t = G.tables_of(CASE["filter_lt"])
    df = t["df"]
    _cmp(O.apply_filters(df, [lambda d: d["a"] < 2]), G.expected_of(CASE["filter_lt"], t), ["b"])
    _cmp(O.filter_or_scalar(df, True), G.expected_of(CASE["filter_scalar_true"], t), ["b"])
    _cmp(O.filter_or_scalar(df, False), G.expected_of(CASE["filter_scalar_false"], t))
    _cmp(O.apply_filters(df, [lambda d: d["a"] < 3, lambda d: d["b"] > 1, lambda d: d["b"] < 3]),
         G.expected_of(CASE["filter_complicated"], t), ["b"])
    nan = G.user_table_nan()
    _cmp(O.apply_filters(nan, [lambda d: d["c"] == 3]), CASE["filter_with_nan"]["expected"])


@pytest.mark.parametrize("name,how", [("join_inner", "inner"), ("join_left", "left"), ("join_right", "right"),
                                      ("join_outer", "outer")])
def test_join_cases(name, how):
    u1 = G.user_table_1().rename(columns={"user_id": "lhs_0", "b": "lhs_1"})
    u2 = G.user_table_2().rename(columns={"user_id": "rhs_0", "c": "rhs_1"})
    out = O.join_on_columns(u1, u2, ["lhs_0"], ["rhs_0"], how)
    # SELECT lhs.user_id, lhs.b, rhs.c
    got = out[["lhs_0", "lhs_1", "rhs_1"]]
    got.columns = ["user_id", "b", "c"]
    _cmp(got, CASE[name]["expected"])


def test_join_anti_residual_nullkeys():
    c = CASE["join_left_anti"]
    t = G.tables_of(c)
    l = t["df_1"].rename(columns={"id": "l_id", "a": "l_a"})
    r = t["df_2"].rename(columns={"id": "r_id", "b": "r_b"})
    got = O.join_on_columns(l, r, ["l_id"], ["r_id"], "leftanti")
    got.columns = ["id", "a"]
    _cmp(got, c["expected"])

    c = CASE["join_equi_plus_residual"]
    u1 = G.user_table_1().rename(columns={"user_id": "lu", "b": "lb"})
    u2 = G.user_table_2().rename(columns={"user_id": "ru", "c": "rc"})
    j = O.join_on_columns(u1, u2, ["lu"], ["ru"], "inner")
    j = O.filter_or_scalar(j, j["rc"] - j["lb"] >= 0)
    _cmp(j, c["expected"])

    c = CASE["join_null_keys_never_match"]
    t = G.tables_of(c)
    j = O.join_on_columns(t["df1"], t["df2"], ["a"], ["c"], "inner")
    j = O.filter_or_scalar(j, ~j["b"].isna())
    _cmp(j, c["expected"], ["b", "c"])


def test_groupby_cases():
    u1 = G.user_table_1()
    got = O.groupby_agg(O.split(u1, 3), ["user_id"], [("b", "S", "sum")])
    _cmp(got, CASE["group_by"]["expected"])

    t = G.tables_of(CASE["group_by_multi"])
    got = O.groupby_agg(O.split(t["df"], 3), ["b"], [("a", "s", "sum"), ("a", "av", "mean"), ("a", "c", "count")])
    _cmp(got[["s", "av", "c"]], CASE["group_by_multi"]["expected"], ["av"])

    # SUM(b), SUM(2): the literal is pre-projected as a constant column (aggregate.py:404-420)
    got = O.groupby_agg(O.split(u1.assign(two=2), 3), [], [("b", "S", "sum"), ("two", "X", "sum")])
    _cmp(got, CASE["group_by_all_literals"]["expected"])

    c = CASE["group_by_all_mixed"]
    t = G.tables_of(c)
    df = t["df"].assign(ab=lambda d: d.a + d.b)
    p = O.groupby_agg(O.split(df, 3), [], [("a", "sum_a", "sum"), ("a", "avg_a", "mean"), ("b", "sum_b", "sum"),
                                           ("b", "avg_b", "mean"), ("ab", "mix_2", "sum"), ("ab", "mix_3", "mean")])
    got = pd.DataFrame({"sum_a": p.sum_a, "avg_a": p.avg_a, "sum_b": p.sum_b, "avg_b": p.avg_b,
                        "mix_1": p.sum_a + p.avg_b, "mix_2": p.mix_2, "mix_3": p.mix_3})
    _cmp(got, G.expected_of(c, t), c["float_cols"])

    # FILTER (WHERE user_id = 2): the bucket is filtered first, groups without rows give NaN
    main = O.groupby_agg(O.split(u1, 3), ["user_id"], [("b", "S2", "sum")])
    filt = O.groupby_agg(O.split(u1[u1.user_id == 2], 3), ["user_id"], [("b", "S1", "sum")])
    got = main.merge(filt, on="user_id", how="left")[["user_id", "S1", "S2"]]
    _cmp(got, CASE["group_by_filtered"]["expected"])

    # NULL and inf keys are groups of their own (dropna=False)
    got = O.groupby_agg([G.user_table_nan()], ["c"], [])
    _cmp(got[["c"]].astype("float64"), CASE["group_by_nan"]["expected"])
    got = O.groupby_agg([G.user_table_inf()], ["c"], [])
    _cmp(got[["c"]], CASE["group_by_inf"]["expected"], ["c"])


def test_filter_columns_post_join():
    c = CASE["filter_columns_post_join"]
    t = G.tables_of(c)
    l = t["df"].rename(columns={"a": "l_a", "c": "l_c"})
    r = t["df2"].rename(columns={"b": "r_b", "c": "r_c"})
    j = O.join_on_columns(l, r, ["l_c"], ["r_c"], "inner")
    got = O.groupby_agg([j], ["r_b"], [("l_a", "sum_a", "sum")])[["sum_a", "r_b"]]
    got.columns = ["sum_a", "b"]
    _cmp(got, c["expected"])


def test_tree_reduction_is_partition_invariant():
    rng = np.random.default_rng(0)
    df = pd.DataFrame({"k": rng.integers(0, 50, 5000), "v": rng.random(5000), "i": rng.integers(-9, 9, 5000)})
    aggs = [("v", "s", "sum"), ("v", "m", "mean"), ("i", "c", "count"), ("i", "lo", "min"), ("i", "hi", "max"),
            (None, "n", "size")]
    base = O.groupby_agg([df], ["k"], aggs)
    for nparts, se in [(2, 2), (7, 3), (16, 8), (5, 2)]:
        got = O.groupby_agg(O.split(df, nparts), ["k"], aggs, split_every=se)
        _cmp(got, base, ["s", "m"])


def test_q3_restatement_matches_sqlite():
    """The composed C4 query against the reference's differential oracle (sqlite3)."""
    rng = np.random.default_rng(4)
    nd, nf = 500, 20_000
    dim = pd.DataFrame({"pk": rng.permutation(nd), "flag": rng.integers(0, 10, nd), "grp": rng.integers(0, 40, nd)})
    fact = pd.DataFrame({"fk": rng.integers(0, nd, nf), "x": rng.integers(-100, 100, nf), "val": rng.random(nf)})
    con = sqlite3.connect(":memory:")
    dim.to_sql("dim", con, index=False)
    fact.to_sql("fact", con, index=False)
    exp = pd.read_sql("""SELECT d.grp, SUM(f.val) AS rev FROM fact f JOIN dim d ON f.fk = d.pk
                         WHERE f.x > 0 AND d.flag < 5 GROUP BY d.grp""", con)
    got = O.c4_q3(O.split(fact, 8), dim)
    _cmp(got, exp, ["rev"])
    exp = pd.read_sql("SELECT SUM(x) FROM fact WHERE x > 0", con)
    got = O.c1_filter_sum(O.split(fact[["x"]], 3))
    assert int(got.iloc[0, 0]) == int(exp.iloc[0, 0])


def test_rex_operator_cases():
    """The operator restatement (oracle rex_*; call.py:140-189, 295-383, 1047-1062) against the
    reference's expression tests."""
    c = CASE["rex_operators"]
    t = G.tables_of(c)
    d = t["df"]
    B = O.REX_BINARY
    got = pd.DataFrame({
        "m": O.rex_reduce(B["*"], d.a, d.b), "u": -d.a, "q": O.rex_sql_div(d.a, d.b, True),
        "s": O.rex_reduce(B["+"], d.a, d.b), "d": O.rex_reduce(B["-"], d.a, d.b),
        "e": O.rex_reduce(B["="], d.a, d.b), "g": O.rex_reduce(B[">"], d.a, d.b),
        "ge": O.rex_reduce(B[">="], d.a, d.b), "l": O.rex_reduce(B["<"], d.a, d.b),
        "le": O.rex_reduce(B["<="], d.a, d.b), "n": O.rex_reduce(B["<>"], d.a, d.b)})
    _cmp(got, G.expected_of(c, t), c["float_cols"])

    nan = G.user_table_nan()
    _cmp(pd.DataFrame({"nn": O.rex_not(O.rex_is_null(nan.c)), "n": O.rex_is_null(nan.c)}), CASE["rex_null"]["expected"])

    c = CASE["rex_integer_div"]
    t = G.tables_of(c)
    a = t["df_simple"].a
    got = pd.DataFrame({"a": O.rex_sql_div(1, a, False), "b": O.rex_sql_div(a, 2, False), "c": O.rex_sql_div(1.0, a, True)})
    _cmp(got, G.expected_of(c, t), ["c"])
    # truncation, not floor (the docstring example of SQLDivisionOperator: -1 / 2 = 0)
    assert O.rex_sql_div(pd.Series([-1, -7, 7]), 2, False).tolist() == [-0.0, -3.0, 3.0]

    c = CASE["rex_boolean_operations"]
    t = G.tables_of(c)
    b = t["df"].b
    got = pd.DataFrame({"t": O.rex_is_true(b), "f": O.rex_is_false(b), "nt": O.rex_not(O.rex_is_true(b)),
                        "nf": O.rex_not(O.rex_is_false(b)), "u": O.rex_is_null(b), "nu": O.rex_not(O.rex_is_null(b))})
    _cmp(got, G.expected_of(c, t))
    # n-ary reduce (tests/unit/test_call.py:130-153: and / or / >= / + over several operands)
    s1, s2, s3 = pd.Series([1, 2, 3]), pd.Series([3, 2, 1]), pd.Series([1, 1, 1])
    assert O.rex_reduce(O.REX_BINARY["+"], s1, s2, s3).tolist() == [5, 5, 5]
    assert O.rex_reduce(O.REX_BINARY["and"], s1 > 1, s2 > 1, s3 > 0).tolist() == [False, True, False]
