"""Golden vectors for the filter -> join -> group-by path, TRANSCRIBED from the reference's own
known-answer tests (dask-contrib/dask-sql @ f186de3, /root/reference/tests).  The reference cannot
be imported in this image (its planner is a Rust crate, dask is absent), so these literals — the
inputs and expected outputs its maintainers wrote down — are the pin for the oracle
(tests/test_oracle_golden.py) and for the CUDA path (tests/test_sql_gpu.py).

Every case cites the reference test it comes from.  `expected` is either the literal frame of the
reference test or, where the reference itself compares against a pandas expression over a seeded
fixture, that same pandas expression (kind="pandas").
Regenerate / extend by reading the cited tests; nothing here is produced by our own code.
"""
import numpy as np
import pandas as pd


# ---- fixtures (tests/integration/fixtures.py) -------------------------------------------------
def df_simple():            # fixtures.py:32-33
    return pd.DataFrame({"a": [1, 2, 3], "b": [1.1, 2.2, 3.3]})


def df():                   # fixtures.py:50-57 (seed 42)
    np.random.seed(42)
    return pd.DataFrame({"a": [1.0] * 100 + [2.0] * 200 + [3.0] * 400, "b": 10 * np.random.rand(700)})


def user_table_1():         # fixtures.py:66-67
    return pd.DataFrame({"user_id": [2, 1, 2, 3], "b": [3, 3, 1, 3]})


def user_table_2():         # fixtures.py:71-72
    return pd.DataFrame({"user_id": [1, 1, 2, 4], "c": [1, 2, 3, 4]})


def long_table():           # fixtures.py:76-77
    return pd.DataFrame({"a": [0] * 100 + [1] * 101 + [2] * 103})


def user_table_inf():       # fixtures.py:81-82
    return pd.DataFrame({"c": [3, float("inf"), 1]})


def user_table_nan():       # fixtures.py:86-87
    return pd.DataFrame({"c": [3, pd.NA, 1]}).astype("UInt8")


FIXTURES = {"df_simple": df_simple, "df": df, "user_table_1": user_table_1, "user_table_2": user_table_2,
            "long_table": long_table, "user_table_inf": user_table_inf, "user_table_nan": user_table_nan}

NaN = np.nan

CASES = [
    # ---- filter -------------------------------------------------------------------------------
    dict(name="filter_lt", cite="tests/integration/test_filter.py:13-17", tables=["df"],
         sql="SELECT * FROM df WHERE a < 2",
         expected=lambda t: t["df"][t["df"]["a"] < 2], float_cols=["b"]),
    dict(name="filter_scalar_true", cite="tests/integration/test_filter.py:20-24", tables=["df"],
         sql="SELECT * FROM df WHERE True", expected=lambda t: t["df"], float_cols=["b"]),
    dict(name="filter_scalar_false", cite="tests/integration/test_filter.py:26-29", tables=["df"],
         sql="SELECT * FROM df WHERE False", expected=lambda t: t["df"].head(0)),
    dict(name="filter_scalar_1eq1", cite="tests/integration/test_filter.py:31-34", tables=["df"],
         sql="SELECT * FROM df WHERE (1 = 1)", expected=lambda t: t["df"], float_cols=["b"]),
    dict(name="filter_scalar_1eq0", cite="tests/integration/test_filter.py:36-39", tables=["df"],
         sql="SELECT * FROM df WHERE (1 = 0)", expected=lambda t: t["df"].head(0)),
    dict(name="filter_complicated", cite="tests/integration/test_filter.py:42-49", tables=["df"],
         sql="SELECT * FROM df WHERE a < 3 AND (b > 1 AND b < 3)",
         expected=lambda t: t["df"][(t["df"]["a"] < 3) & ((t["df"]["b"] > 1) & (t["df"]["b"] < 3))],
         float_cols=["b"]),
    dict(name="filter_with_nan", cite="tests/integration/test_filter.py:52-59", tables=["user_table_nan"],
         sql="SELECT * FROM user_table_nan WHERE c = 3", expected=pd.DataFrame({"c": [3]})),
    # ---- join ---------------------------------------------------------------------------------
    dict(name="join_inner", cite="tests/integration/test_join.py:14-43", tables=["user_table_1", "user_table_2"],
         sql="""SELECT lhs.user_id, lhs.b, rhs.c FROM user_table_1 AS lhs
                JOIN user_table_2 AS rhs ON lhs.user_id = rhs.user_id""",
         expected=pd.DataFrame({"user_id": [1, 1, 2, 2], "b": [3, 3, 1, 3], "c": [1, 2, 3, 3]})),
    dict(name="join_outer", cite="tests/integration/test_join.py:46-65", tables=["user_table_1", "user_table_2"],
         sql="""SELECT lhs.user_id, lhs.b, rhs.c FROM user_table_1 AS lhs
                FULL JOIN user_table_2 AS rhs ON lhs.user_id = rhs.user_id""",
         expected=pd.DataFrame({"user_id": [1, 1, 2, 2, 3, NaN], "b": [3, 3, 1, 3, 3, NaN],
                                "c": [1, 2, 3, 3, NaN, 4]})),
    dict(name="join_left", cite="tests/integration/test_join.py:68-87", tables=["user_table_1", "user_table_2"],
         sql="""SELECT lhs.user_id, lhs.b, rhs.c FROM user_table_1 AS lhs
                LEFT JOIN user_table_2 AS rhs ON lhs.user_id = rhs.user_id""",
         expected=pd.DataFrame({"user_id": [1, 1, 2, 2, 3], "b": [3, 3, 1, 3, 3], "c": [1, 2, 3, 3, NaN]})),
    dict(name="join_right", cite="tests/integration/test_join.py:140-159", tables=["user_table_1", "user_table_2"],
         sql="""SELECT lhs.user_id, lhs.b, rhs.c FROM user_table_1 AS lhs
                RIGHT JOIN user_table_2 AS rhs ON lhs.user_id = rhs.user_id""",
         expected=pd.DataFrame({"user_id": [1, 1, 2, 2, NaN], "b": [3, 3, 1, 3, NaN], "c": [1, 2, 3, 3, 4]})),
    dict(name="join_left_anti", cite="tests/integration/test_join.py:90-112 (numeric columns)",
         inline_tables={"df_1": pd.DataFrame({"id": [1, 1, 2, 4], "a": [10, 11, 12, 13]}),
                        "df_2": pd.DataFrame({"id": [2, 1, 2, 3], "b": [20, 21, 22, 23]})},
         sql="SELECT lhs.id, lhs.a FROM df_1 AS lhs LEFT ANTI JOIN df_2 AS rhs ON lhs.id = rhs.id",
         expected=pd.DataFrame({"id": [4], "a": [13]})),
    dict(name="join_equi_plus_residual", cite="tests/integration/test_join.py:205-216",
         tables=["user_table_1", "user_table_2"],
         sql="""SELECT lhs.user_id, lhs.b, rhs.user_id, rhs.c FROM user_table_1 AS lhs
                JOIN user_table_2 AS rhs ON rhs.user_id = lhs.user_id AND rhs.c - lhs.b >= 0""",
         expected=pd.DataFrame({"lhs.user_id": [2, 2], "b": [1, 3], "rhs.user_id": [2, 2], "c": [3, 3]})),
    dict(name="join_null_keys_never_match", cite="tests/integration/test_join.py:260-281 (numeric columns)",
         inline_tables={"df1": pd.DataFrame({"a": [1, 2, 2, 5, 6], "b": [10.0, 11.0, 12.0, NaN, 14.0]}),
                        "df2": pd.DataFrame({"c": [NaN, 3, 2, 5], "d": [20, 21, 22, 23]})},
         sql="SELECT * FROM df1 INNER JOIN df2 ON (a = c AND b IS NOT NULL)",
         expected=pd.DataFrame({"a": [2, 2], "b": [11.0, 12.0], "c": [2.0, 2.0], "d": [22, 22]}),
         float_cols=["b", "c"]),
    dict(name="filter_columns_post_join", cite="tests/integration/test_join.py:442-457",
         inline_tables={"df": pd.DataFrame({"a": [1, 2, 3, 4, 5], "c": [1, None, 2, 2, 2]}),
                        "df2": pd.DataFrame({"b": [1, 1, 2, 2, 3], "c": [2, 2, 2, 2, 2]})},
         sql="SELECT SUM(df.a) as sum_a, df2.b FROM df INNER JOIN df2 ON df.c=df2.c GROUP BY df2.b",
         expected=pd.DataFrame({"sum_a": [24, 24, 12], "b": [1, 2, 3]})),
    # ---- group by -----------------------------------------------------------------------------
    dict(name="group_by", cite="tests/integration/test_groupby.py:25-36", tables=["user_table_1"],
         sql='SELECT user_id, SUM(b) AS "S" FROM user_table_1 GROUP BY user_id',
         expected=pd.DataFrame({"user_id": [1, 2, 3], "S": [3, 4, 3]})),
    dict(name="group_by_multi", cite="tests/integration/test_groupby.py:39-65",
         inline_tables={"df": pd.DataFrame({"a": [1, 2, 3], "b": [1, 1, 2]})},
         sql="SELECT SUM(a) AS s, AVG(a) AS av, COUNT(a) AS c FROM df GROUP BY b",
         expected=pd.DataFrame({"s": [3, 3], "av": [1.5, 3.0], "c": [2, 1]}), float_cols=["av"]),
    dict(name="group_by_all_literals", cite="tests/integration/test_groupby.py:70-80", tables=["user_table_1"],
         sql='SELECT SUM(b) AS "S", SUM(2) AS "X" FROM user_table_1',
         expected=pd.DataFrame({"S": [10], "X": [8]})),
    dict(name="group_by_all_mixed", cite="tests/integration/test_groupby.py:82-107", tables=["df"],
         sql="""SELECT SUM(a) AS sum_a, AVG(a) AS avg_a, SUM(b) AS sum_b, AVG(b) AS avg_b,
                SUM(a)+AVG(b) AS mix_1, SUM(a+b) AS mix_2, AVG(a+b) AS mix_3 FROM df""",
         expected=lambda t: pd.DataFrame({
             "sum_a": [t["df"].a.sum()], "avg_a": [t["df"].a.mean()], "sum_b": [t["df"].b.sum()],
             "avg_b": [t["df"].b.mean()], "mix_1": [t["df"].a.sum() + t["df"].b.mean()],
             "mix_2": [(t["df"].a + t["df"].b).sum()], "mix_3": [(t["df"].a + t["df"].b).mean()]}),
         float_cols=["sum_a", "avg_a", "sum_b", "avg_b", "mix_1", "mix_2", "mix_3"]),
    dict(name="group_by_filtered_global", cite="tests/integration/test_groupby.py:110-121", tables=["user_table_1"],
         sql='SELECT SUM(b) FILTER (WHERE user_id = 2) AS "S1", SUM(b) "S2" FROM user_table_1',
         expected=pd.DataFrame({"S1": [4], "S2": [10]}, dtype="int64")),
    dict(name="group_by_filtered", cite="tests/integration/test_groupby.py:123-141", tables=["user_table_1"],
         sql='SELECT user_id, SUM(b) FILTER (WHERE user_id = 2) AS "S1", SUM(b) "S2" FROM user_table_1 GROUP BY user_id',
         expected=pd.DataFrame({"user_id": [1, 2, 3], "S1": [NaN, 4.0, NaN], "S2": [3, 4, 3]})),
    dict(name="group_by_filtered_only", cite="tests/integration/test_groupby.py:143-151", tables=["user_table_1"],
         sql='SELECT SUM(b) FILTER (WHERE user_id = 2) AS "S1" FROM user_table_1',
         expected=pd.DataFrame({"S1": [4]})),
    dict(name="group_by_nan", cite="tests/integration/test_groupby.py:174-186", tables=["user_table_nan"],
         sql="SELECT c FROM user_table_nan GROUP BY c",
         expected=pd.DataFrame({"c": [3.0, 1.0, NaN]})),
    dict(name="group_by_inf", cite="tests/integration/test_groupby.py:188-202", tables=["user_table_inf"],
         sql="SELECT c FROM user_table_inf GROUP BY c",
         expected=pd.DataFrame({"c": [3.0, 1.0, float("inf")]}), float_cols=["c"]),
]


def _bool_table():          # tests/integration/test_rex.py:457-461
    return pd.DataFrame({"b": pd.array([True, False, pd.NA], dtype="boolean")})


def _operators_expected(t):  # tests/integration/test_rex.py:223-235
    d = t["df"]
    e = pd.DataFrame(index=d.index)
    e["m"], e["u"], e["q"], e["s"], e["d"] = d["a"] * d["b"], -d["a"], d["a"] / d["b"], d["a"] + d["b"], d["a"] - d["b"]
    e["e"], e["g"], e["ge"] = d["a"] == d["b"], d["a"] > d["b"], d["a"] >= d["b"]
    e["l"], e["le"], e["n"] = d["a"] < d["b"], d["a"] <= d["b"], d["a"] != d["b"]
    return e


CASES += [
    # ---- expressions (predicate / pre-projection arithmetic of the path) --------------------
    dict(name="rex_operators", cite="tests/integration/test_rex.py:205-236", tables=["df"],
         sql="SELECT a * b AS m, -a AS u, a / b AS q, a + b AS s, a - b AS d, a = b AS e, a > b AS g, "
             "a >= b AS ge, a < b AS l, a <= b AS le, a <> b AS n FROM df",
         expected=_operators_expected, float_cols=["m", "u", "q", "s", "d"]),
    dict(name="rex_null", cite="tests/integration/test_rex.py:363-377", tables=["user_table_nan"],
         sql="SELECT c IS NOT NULL AS nn, c IS NULL AS n FROM user_table_nan",
         expected=pd.DataFrame({"nn": [True, False, True], "n": [False, True, False]})),
    dict(name="rex_integer_div", cite="tests/integration/test_rex.py:548-568", tables=["df_simple"],
         sql="SELECT 1 / a AS a, a / 2 AS b, 1.0 / a AS c FROM df_simple",
         expected=lambda t: pd.DataFrame({"a": (1 // t["df_simple"].a).astype("Int64"),
                                          "b": (t["df_simple"].a // 2).astype("Int64"),
                                          "c": 1 / t["df_simple"].a}), float_cols=["c"]),
    dict(name="rex_boolean_operations", cite="tests/integration/test_rex.py:456-485",
         inline_tables={"df": _bool_table()},
         sql="SELECT b IS TRUE AS t, b IS FALSE AS f, b IS NOT TRUE AS nt, b IS NOT FALSE AS nf, "
             "b IS UNKNOWN AS u, b IS NOT UNKNOWN AS nu FROM df",
         expected=lambda t: pd.DataFrame({"t": t["df"].b.fillna(False), "f": ~t["df"].b.fillna(True),
                                          "nt": ~t["df"].b.fillna(False), "nf": t["df"].b.fillna(True),
                                          "u": t["df"].b.isna(), "nu": ~t["df"].b.isna()})),
]


def tables_of(case):
    t = {n: FIXTURES[n]() for n in case.get("tables", [])}
    t.update(case.get("inline_tables", {}))
    return t


def expected_of(case, tables):
    e = case["expected"]
    return e(tables) if callable(e) else e
