"""CPU tests of the drop-in boundary: plugin registry, containers, planner shapes, C-ABI symbols,
expression compilation.  No kernel is launched (there is no GPU here and no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- C-ABI -----------------------------------------------------------------------------------
def test_library_loads_and_exports_every_declared_symbol():
    from dask_sql_b200 import _lib
    header = open(os.path.join(ROOT, "include", "b200sql.h")).read()
    declared = sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", header)))
    declared = [d for d in declared if not d.endswith("_t")]
    assert len(declared) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [d for d in declared if not hasattr(lib, d)]
    assert not missing, f"declared in include/b200sql.h but not exported: {missing}"
    assert set(_lib.EXPORTS) <= set(declared)
    assert _lib.version() == 100
    assert _lib.num_tiles(4097) == 2
    # pure host helpers of the ABI
    for x in (0.0, -0.0, 1.5, -2.25, 1e300, -1e-300):
        assert _lib.ordered_to_f64(_lib.f64_to_ordered(x)) == x
    xs = [-3.0, -1.0, -0.0, 0.0, 2.0, 7.5]
    ks = [_lib.f64_to_ordered(x) for x in xs]
    assert ks == sorted(ks)


def test_struct_layouts_match_header_sizes():
    from dask_sql_b200 import _lib as L
    assert ctypes.sizeof(L.Scan) == 16 * 24 + 8 * 32 + 16
    assert ctypes.sizeof(L.Prog) == 64 * 24 + 8


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dask-sql_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# oracle", ""), f


# ---- plugin registry (reference: tests/unit/test_utils.py:33-58) ------------------------------
def test_pluggable_replace_semantics():
    from dask_sql_b200.utils import Pluggable

    class PluginTest1(Pluggable):
        pass

    class PluginTest2(Pluggable):
        pass

    PluginTest1.add_plugin("some_key", "value")
    assert PluginTest1.get_plugin("some_key") == "value"
    assert PluginTest1.get_plugins() == ["value"]
    with pytest.raises(KeyError):
        PluginTest2.get_plugin("some_key")
    PluginTest1.add_plugin("some_key", "value_2")
    assert PluginTest1.get_plugin("some_key") == "value_2"
    PluginTest1.add_plugin("some_key", "value_3", replace=False)
    assert PluginTest1.get_plugin("some_key") == "value_2"
    PluginTest1.add_plugin(["a", "b"], "multi")
    assert PluginTest1.get_plugin("a") == PluginTest1.get_plugin("b") == "multi"


def test_custom_rel_plugin_drops_in():
    from dask_sql_b200 import BaseRelPlugin, Context, RelConverter
    from dask_sql_b200.physical.rel.logical import DaskFilterPlugin

    calls = []

    class MyFilter(DaskFilterPlugin):
        class_name = "Filter"

        def convert(self, rel, context):
            calls.append(rel.get_current_node_type())
            return super().convert(rel, context)

    c = Context()
    RelConverter.add_plugin_class(MyFilter, replace=True)
    try:
        c.create_table("t", pd.DataFrame({"x": [1, 2, 3]}))
        c.sql("SELECT x FROM t WHERE x > 1", config_options={"sql.optimize": False})
        assert calls == ["Filter"]
        assert issubclass(MyFilter, BaseRelPlugin)
    finally:
        RelConverter.add_plugin_class(DaskFilterPlugin, replace=True)


# ---- containers (reference: tests/unit/test_datacontainer.py:4-67) ----------------------------
def test_column_container():
    from dask_sql_b200.datacontainer import ColumnContainer
    c = ColumnContainer(["a", "b", "c"])
    assert c.columns == ["a", "b", "c"]
    assert c.mapping() == [("a", "a"), ("b", "b"), ("c", "c")]
    c2 = c.limit_to(["c", "a"])
    assert c2.columns == ["c", "a"] and c.columns == ["a", "b", "c"]
    c3 = c.rename({"a": "A", "b": "a"})
    assert c3.columns == ["A", "a", "c"]
    assert c3.get_backend_by_frontend_name("A") == "a" and c3.get_backend_by_frontend_name("a") == "b"
    assert c3.get_backend_by_frontend_index(1) == "b"
    c4 = c.add("d").add("e", "a")
    assert c4.columns == ["a", "b", "c", "d", "e"] and c4.get_backend_by_frontend_name("e") == "a"
    c5 = c.make_unique("x")
    assert c5.columns == ["x_0", "x_1", "x_2"] and c5.get_backend_by_frontend_name("x_2") == "c"
    c6 = c.rename_handle_duplicates(["a", "a"], ["p", "q"])
    assert c6.get_backend_by_frontend_name("p") == "a" and c6.get_backend_by_frontend_name("q") == "a"
    # the reference's known answers, including the ORDER of mapping() (tests/unit/test_datacontainer.py:4-67)
    ident = [("a", "a"), ("b", "b"), ("c", "c")]
    assert ColumnContainer(["a", "b", "c"], {"a": "1", "b": "2", "c": "3"}).mapping() == [("a", "1"), ("b", "2"), ("c", "3")]
    assert c2.mapping() == ident and c.mapping() == ident                     # limit_to keeps every name resolvable
    assert c3.mapping() == [("a", "b"), ("b", "b"), ("c", "c"), ("A", "a")] and c.mapping() == ident
    assert c.add("d").mapping() == ident + [("d", "d")]
    assert c.add("d", "D").mapping() == ident + [("d", "D")] and c.add("d", "D").columns == ["a", "b", "c", "d"]
    assert c.add("d", "a").mapping() == ident + [("d", "a")]
    assert c.add("a", "b").columns == ["a", "b", "c"] and c.add("a", "b").mapping() == [("a", "b"), ("b", "b"), ("c", "c")]
    assert c.columns == ["a", "b", "c"] and c.mapping() == ident              # nothing above touched the original
    with pytest.raises(AssertionError):
        ColumnContainer(["a", 1])


# ---- planner shapes ---------------------------------------------------------------------------
def _ctx():
    from dask_sql_b200 import Context
    c = Context()
    rng = np.random.default_rng(0)
    c.create_table("fact", pd.DataFrame({"fk": rng.integers(0, 10, 64), "x": rng.integers(-5, 5, 64),
                                         "val": rng.random(64)}), npartitions=3)
    c.create_table("dim", pd.DataFrame({"pk": np.arange(10), "flag": rng.integers(0, 10, 10),
                                        "grp": rng.integers(0, 3, 10)}))
    return c


def test_c1_plan_pushes_filter_into_scan():
    c = _ctx()
    plan = c.explain("SELECT SUM(x) FROM fact WHERE x > 0")
    # SURVEY 3.2 / tests/integration/test_join.py:482-492 plan text style
    assert plan.splitlines()[-1].strip() == "TableScan: fact projection=[x], full_filters=[fact.x > Int64(0)]"
    assert "Aggregate: groupBy=[[]], aggr=[[SUM(fact.x)]]" in plan
    assert "Filter:" not in plan


def test_q3_plan_shape_and_lazy_graph():
    from dask_sql_b200.frame import AggSource, JoinSource, TableSource
    c = _ctx()
    q = """SELECT d.grp, SUM(f.val) AS rev FROM fact f JOIN dim d ON f.fk = d.pk
           WHERE f.x > 0 AND d.flag < 5 GROUP BY d.grp"""
    plan = c.explain(q)
    assert "Inner Join: f.fk = d.pk" in plan
    assert "full_filters=[fact.x > Int64(0), fact.fk IS NOT NULL]" in plan
    assert "full_filters=[dim.flag < Int64(5), dim.pk IS NOT NULL]" in plan
    lazy = c.sql(q)
    assert lazy.columns == ["grp", "rev"]
    agg = lazy.source
    assert isinstance(agg, AggSource) and [a[2] for a in agg.aggs] == ["sum"]
    join = agg.child.source
    assert isinstance(join, JoinSource) and join.how == "inner"
    assert isinstance(join.left.source, TableSource) and len(join.left.pred) == 2
    assert isinstance(join.right.source, TableSource) and len(join.right.pred) == 2


def test_order_by_limit_and_moment_aggregates_plan():
    from dask_sql_b200.frame import AggSource, LimitSource, SortSource
    c = _ctx()
    q = """SELECT d.grp, SUM(f.val) AS rev, STDDEV(f.val) AS sd FROM fact f JOIN dim d ON f.fk = d.pk
           WHERE f.x > 0 GROUP BY d.grp ORDER BY rev DESC, d.grp LIMIT 10 OFFSET 2"""
    plan = c.explain(q).splitlines()
    assert plan[0] == "Limit: skip=2, fetch=10" and plan[1].strip() == "Sort: rev DESC, d.grp ASC"
    lazy = c.sql(q)
    assert lazy.columns == ["grp", "rev", "sd"]
    lim = lazy.source
    assert isinstance(lim, LimitSource) and (lim.offset, lim.fetch) == (2, 10)
    srt = lim.child.source
    assert isinstance(srt, SortSource) and [(a, nf) for _, a, nf in srt.keys] == [(False, True), (True, False)]
    # STDDEV travels by name in the same AggSource as SUM: the executor expands it into shifted
    # (count, sum, sum of squares) accumulators of the same fused pass
    agg = srt.child.source
    assert isinstance(agg, AggSource)
    assert sorted(f for _, _, f in agg.aggs) == ["stddev_samp", "sum"]
    from dask_sql_b200 import executor as X, _lib as L
    child = agg.child
    aggs = [(child.exprs[i], o, f) for i, o, f in agg.aggs]
    plan = X.AggPlan(aggs, lambda e: True, {repr(aggs[1][0]): 12.5})
    assert [k.op for k in plan.kaggs] == [L.AGG_SUM] * 3 and "sub(" in repr(plan.kaggs[1].expr)
    assert list(plan.second.values()) == [2]


def test_unoptimized_plan_keeps_filter_node():
    c = _ctx()
    lazy = c.sql("SELECT x FROM fact WHERE x > 0", config_options={"sql.optimize": False})
    assert len(lazy.pred) == 1


def test_parse_errors_are_parsing_exceptions():
    from dask_sql_b200.utils import ParsingException
    c = _ctx()
    for bad in ["SELEC x FROM fact", "SELECT nope FROM fact", "SELECT x FROM missing_table",
                "SELECT x, SUM(val) FROM fact", "SELECT x FROM fact WHERE"]:
        with pytest.raises(ParsingException):
            c.sql(bad)


def test_out_of_scope_features_fail_loudly():
    c = _ctx()
    with pytest.raises(NotImplementedError):
        c.sql("SELECT * FROM fact, dim")
    with pytest.raises(Exception):
        c.sql("SELECT UPPER(x) FROM fact")


# ---- expression IR ----------------------------------------------------------------------------
def test_term_extraction_and_program_compilation():
    from dask_sql_b200 import _lib as L
    from dask_sql_b200 import expr as E
    x, y = E.ColRef("x", E.I64), E.ColRef("y", E.F64)
    assert E.as_term(E.binop("gt", x, 0)) == ("x", L.GT, 0)
    assert E.as_term(E.binop("lt", 3, x)) == ("x", L.GT, 3)            # flipped
    assert E.as_term(E.binop("le", y, 2)) == ("y", L.LE, 2.0)
    assert E.as_term(E.binop("gt", x, 0.5)) == ("x", L.GT, 0.5)        # int column, float literal
    assert E.as_term(E.unop("not", E.unop("isnull", x))) == ("x", L.IS_NOT_NULL, 0)
    assert E.as_term(E.binop("gt", E.binop("add", x, 1), 0)) is None   # needs the interpreter
    e = E.fillna(E.binop("and", E.binop("gt", x, 0), E.binop("lt", y, 1.5)), False)
    assert len(E.conjuncts(e)) == 2
    prog = E.compile_expr(E.binop("mul", E.binop("add", x, 1), y), ["x", "y"])
    ops = [prog.code[i].op for i in range(prog.n)]
    assert ops == [L.OP_LOAD, L.OP_CONST_I, L.OP_ADD_I, L.OP_I2F, L.OP_LOAD, L.OP_MUL_F]
    assert prog.out_dtype == L.F64
    assert E.binop("divt", x, 2).dtype == E.I64 and E.binop("truediv", x, 2).dtype == E.F64


def test_host_column_roundtrip_keeps_logical_dtypes():
    import torch
    from dask_sql_b200.device import DeviceColumn, column_to_host
    from dask_sql_b200.table import _host_column
    s = pd.Series(pd.array([1, None, 3, None, 5] * 13, dtype="Int64"))
    hc = _host_column(s, pin=False)
    assert hc.dtype == 0 and hc.valid is not None and hc.logical == "Int64"
    back = column_to_host(DeviceColumn(hc.data, hc.valid, hc.dtype, hc.logical))
    assert pd.Series(back).isna().tolist() == s.isna().tolist()
    assert pd.Series(back).dropna().astype(int).tolist() == s.dropna().astype(int).tolist()
    f = _host_column(pd.Series([1.5, np.nan, 2.5]), pin=False)
    assert f.dtype == 1 and f.valid is None
    b = _host_column(pd.Series([True, False, True]), pin=False)
    assert b.dtype == 2 and b.data.dtype == torch.uint8
    with pytest.raises(NotImplementedError):
        _host_column(pd.Series(["a", "b"]), pin=False)


def test_ddl_statements_plan_like_the_reference():
    """CREATE TABLE AS / CREATE VIEW / CREATE TABLE WITH / DROP TABLE produce plan nodes with the
    accessor surface the reference's custom plugins call (physical/rel/custom/*.py)."""
    from dask_sql_b200.planner import plan_sql

    catalog = lambda schema, table: ("root", [("a", "BIGINT"), ("b", "DOUBLE")]) if table == "x" else None
    p = plan_sql("CREATE OR REPLACE TABLE s.t AS (SELECT a FROM x WHERE a > 1)", catalog)
    assert p.get_current_node_type() == "CreateMemoryTable"
    cmt = p.create_memory_table()
    assert (cmt.getQualifiedName(), cmt.getOrReplace(), cmt.getIfNotExists(), cmt.isTable()) == ("s.t", True, False, True)
    assert cmt.getInput().get_current_node_type() in ("Projection", "TableScan")
    v = plan_sql("create view v as select b from x", catalog)
    assert v.get_current_node_type() == "CreateView" and not v.create_memory_table().isTable()
    ct = plan_sql("CREATE TABLE IF NOT EXISTS t WITH (location = '/d/a.parquet', format = 'parquet', persist = True, "
                  "npartitions = 4)", catalog).create_table()
    assert ct.getTableName() == "t" and ct.getSchemaName() is None and ct.getIfNotExists()
    assert ct.getSQLWithOptions() == {"location": "/d/a.parquet", "format": "parquet", "persist": True, "npartitions": 4}
    dt = plan_sql("DROP TABLE IF EXISTS s.t;", catalog).drop_table()
    assert dt.getQualifiedName() == "s.t" and dt.getIfExists()
    # the words stay usable as identifiers
    assert plan_sql("SELECT a AS view, b AS replace FROM x", catalog).getRowType().getFieldNames() == ["view", "replace"]


def test_arrow_buffers_become_host_columns_without_pandas(tmp_path):
    import pyarrow as pa
    import pyarrow.parquet as pq
    from dask_sql_b200.table import arrow_columns, read_location, _host_column, DeviceTable, ArrowColumn

    n = 203
    vals = np.arange(n)
    t = pa.table({"i": pa.array([None if v % 7 == 0 else int(v) for v in vals], type=pa.int16()),
                  "f": pa.array([None if v % 5 == 0 else v / 2 for v in vals], type=pa.float32()),
                  "b": pa.array([None if v % 11 == 0 else bool(v & 1) for v in vals]),
                  "d": pa.array(vals, type=pa.int64())})
    path = str(tmp_path / "x.parquet")
    pq.write_table(t, path, row_group_size=64)
    back = read_location(path, None, columns=["i", "f", "b", "d"])
    assert back.column("i").num_chunks > 1
    for tab in (t, back, back.slice(5, 150)):
        cols = arrow_columns(tab)
        ref = tab.to_pandas()
        m = len(ref)
        assert isinstance(cols["d"], ArrowColumn) and cols["i"].logical == "Int16" and cols["d"].logical == "int64"
        table = DeviceTable.from_columns(cols, npartitions=3, device=None, persist=False)
        assert table.nrows == m and len(table.partitions) == 3
        got_i, got_valid = [], []
        for part in table.partitions:
            hc = part["i"]
            got_i.append(hc.data.numpy())
            words = hc.valid.numpy().view(np.uint8) if hc.valid is not None else np.full((hc.n + 7) // 8, 255, np.uint8)
            got_valid.append(np.unpackbits(words, bitorder="little")[: hc.n].astype(bool))
        got_i, got_valid = np.concatenate(got_i), np.concatenate(got_valid)
        exp_valid = ~ref["i"].isna().to_numpy()
        np.testing.assert_array_equal(got_valid, exp_valid)
        np.testing.assert_array_equal(got_i[exp_valid], ref["i"].to_numpy(dtype=float)[exp_valid].astype(np.int64))
        f = np.concatenate([p["f"].data.numpy() for p in table.partitions])
        np.testing.assert_array_equal(np.isnan(f), ref["f"].isna().to_numpy())
        assert all(p["f"].valid is None and p["d"].valid is None for p in table.partitions)
    with pytest.raises(NotImplementedError):
        arrow_columns(pa.table({"s": ["a", "b"]}))
    with pytest.raises(AttributeError):
        read_location(path, "orc")


def test_pending_part_resolves_on_first_use_only():
    """executor.PendingPart: a partition whose row count is still on the device behaves like a Part
    once anybody looks at it, and not before (host logic only; the kernels are covered by -m gpu)."""
    from dask_sql_b200 import executor as X

    calls = []

    def thunk():
        calls.append(1)
        return X.Part({"a": "col-a", "b": "col-b"}, 7)

    p = X.PendingPart(thunk)
    assert not p.resolved and calls == []
    q = X.PendingPart(lambda: X.Part({"x": p.resolve()["a"]}, p.n))       # chained (execute()'s projection)
    assert not q.resolved and calls == []
    assert q.n == 7 and q.resolved and p.resolved and calls == [1]
    assert q["x"] == "col-a" and list(q) == ["x"] and "x" in q and len(q) == 1 and q.get("y") is None
    assert p.n == 7 and dict(p) == {"a": "col-a", "b": "col-b"} and list(p.items())[0] == ("a", "col-a")
    assert calls == [1]                                                     # resolved exactly once
    assert X.Part({"k": 1}, 3).resolve().n == 3


def test_sql_to_lazy_frames_without_a_gpu():
    """Planning + plugin conversion is host-only: with host-resident tables (persist=False, the
    reference's default) every statement reaches a LazyFrame / catalog change without touching CUDA."""
    import pandas as pd
    from dask_sql_b200 import Context
    from dask_sql_b200.frame import LazyFrame
    from dask_sql_b200.physical.rel import RelConverter

    c = Context()
    c.create_table("t", pd.DataFrame({"a": [1, 2, 3, 4], "b": [1.5, 2.5, 3.5, 4.5], "k": [1, 1, 2, 2]}))
    for q, cols in [("SELECT a, a + b AS s, b * 2 AS d FROM t WHERE a > 1", ["a", "s", "d"]),
                    ("SELECT k, SUM(b) AS sb FROM t GROUP BY k HAVING SUM(b) > 3", ["k", "sb"]),
                    ("SELECT x.k, y.a FROM t x JOIN t y ON x.a = y.a WHERE x.b + y.b > 3", ["k", "a"]),
                    ("SELECT a, b FROM t ORDER BY b DESC LIMIT 2", ["a", "b"]),
                    ("SELECT COUNT(*) AS n, AVG(b) AS m FROM t WHERE NOT (a = 2)", ["n", "m"])]:
        lf = c.sql(q)
        assert isinstance(lf, LazyFrame) and lf.columns == cols
    assert c.sql("CREATE VIEW v AS SELECT a, b FROM t WHERE a > 2") is None
    assert c.sql("SELECT * FROM v").columns == ["a", "b"]
    c.sql("DROP TABLE v")
    with pytest.raises(RuntimeError, match="not present"):
        c.sql("DROP TABLE v")
    c.sql("DROP TABLE IF EXISTS v")
    with pytest.raises(RuntimeError, match="already present"):
        c.sql("CREATE VIEW t AS SELECT a FROM t")
    c.sql("CREATE VIEW IF NOT EXISTS t AS SELECT a FROM t")
    assert c.sql("SELECT * FROM t").columns == ["a", "b", "k"]          # kept

    class Unknown:
        def get_current_node_type(self):
            return "WindowAggr"

    with pytest.raises(NotImplementedError):
        RelConverter.convert(Unknown(), c)


def test_header_is_plain_c():
    """include/b200sql.h must be consumable by a C compiler on its own (cgo / JNI / ctypes bind it):
    C99, no C++ or torch types in any signature."""
    import shutil
    import subprocess
    import tempfile

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "h.c")
        with open(src, "w") as f:
            f.write('#include "b200sql.h"\nint main(void) { return (int)sizeof(b2_scan_t) * 0; }\n')
        r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only",
                            "-I", os.path.join(root, "include"), src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = open(os.path.join(root, "include", "b200sql.h")).read()
    assert "at::" not in text and "std::" not in text and "template" not in text


def test_partition_plan_heuristics(monkeypatch):
    """executor._partition_plan (host logic): range-partition only group tables far beyond L2 whose
    inputs are plain 8-byte columns; buckets of ~24 MB of table, at most 1024 of them."""
    from types import SimpleNamespace as NS
    from dask_sql_b200 import executor as X, _lib as L

    col = lambda dtype, valid=None: NS(dtype=dtype, valid=valid)
    plan = NS(kaggs=[NS(op=L.AGG_SUM, need_cnt=False)], need_rows=True)          # SUM + AVG of a non-null column
    ctx = NS(cols=[col(L.I64), col(L.F64)])
    work = [(NS(n=10), ctx, 0, [1])]
    monkeypatch.delenv("B200SQL_NO_PARTITION", raising=False)
    monkeypatch.delenv("B200SQL_PARTITION_MIN_BYTES", raising=False)
    monkeypatch.delenv("B200SQL_PARTITION_BUCKET_BYTES", raising=False)
    assert X._partition_plan(1_000_001, plan, work, 200_000_000) is None         # 16 MB of table: stays in L2
    shift, nb = X._partition_plan(100_000_001, plan, work, 500_000_000)          # C5: 1.6 GB of table
    assert (shift, nb) == (20, 96) and ((100_000_001 - 1) >> shift) < nb
    assert X._partition_plan(100_000_001, plan, work, 1_000_000) is None         # too few rows to pay for a pass
    shift, nb = X._partition_plan(1 << 31, plan, work, 1 << 33)
    assert nb <= 1024 and (((1 << 31) - 1) >> shift) < nb                        # bucket count is capped
    nullable = [(NS(n=10), NS(cols=[col(L.I64), col(L.F64, valid=object())]), 0, [1])]
    assert X._partition_plan(100_000_001, plan, nullable, 500_000_000) is None   # bitmap inputs: direct path
    boolean = [(NS(n=10), NS(cols=[col(L.I64), col(L.U8)]), 0, [1])]
    assert X._partition_plan(100_000_001, plan, boolean, 500_000_000) is None
    monkeypatch.setenv("B200SQL_NO_PARTITION", "1")
    assert X._partition_plan(100_000_001, plan, work, 500_000_000) is None


def test_lazy_parquet_table_prunes_row_groups(tmp_path):
    """persist=False Parquet tables read per query: only referenced columns, only the row groups whose
    min/max/null-count statistics admit the pushed-down conjuncts (physical/utils/filter.py:17,
    table_scan.py:80-99).  Host-side logic only: which row groups survive, what schema is reported."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    from dask_sql_b200 import Context, _lib as L
    from dask_sql_b200.table import ParquetTable
    n = 10_000
    rng = np.random.default_rng(0)
    df = pd.DataFrame({"a": np.arange(n), "b": rng.random(n),
                       "c": pd.array(np.where(np.arange(n) % 7 == 0, None, np.arange(n) % 50), dtype="Int64")})
    path = str(tmp_path / "t.parquet")
    pq.write_table(pa.Table.from_pandas(df), path, row_group_size=1000)
    t = ParquetTable(path)
    assert (t.nrows, t.npartitions) == (n, 10) and not t._cache           # nothing read yet
    assert [(nm, lg) for nm, _, lg in t.schema()] == [("a", "int64"), ("b", "float64"), ("c", "Int64")]
    st = t.column_stats("a")
    assert (st.vmin, st.vmax, st.nulls) == (0, n - 1, 0) and not t._cache   # from the footer, not the pages
    assert t.surviving_groups([("a", L.GT, 8500)]) == [8, 9]
    assert t.surviving_groups([("a", L.LT, 1000)]) == [0]
    assert t.surviving_groups([("a", L.EQ, 4242), ("b", L.GT, 0.0)]) == [4]
    assert t.surviving_groups([("c", L.GE, 50)]) == []                   # c < 50 everywhere
    assert len(t.surviving_groups([("c", L.IS_NULL, 0)])) == 10
    parts = t.scan_pruned({"a", "b"}, [("a", L.GE, 9000)])
    assert len(parts) == 1 and set(parts[0]) == {"a", "b"} and parts[0]["a"].n == 1000
    assert t.stats["row_groups_skipped"] == 9 and set(k[1] for k in t._cache) == {"a", "b"}
    parts = t.scan_pruned({"a"}, [("a", L.GT, 10**9)])                  # nothing survives: one empty partition
    assert len(parts) == 1 and parts[0]["a"].n == 0
    # registration is lazy and keeps the location (the reference stores it in DataContainer.filepath)
    c = Context()
    c.create_table("t", path)
    dc = c.schema[c.schema_name].tables["t"]
    assert dc.filepath == path and isinstance(dc.df.source.table, ParquetTable) and dc.df.npartitions == 10
    lazy = c.sql("SELECT SUM(b) AS s FROM t WHERE a >= 9000 AND c > 3")
    assert [repr(p) for p in lazy.source.child.pred][:1]                 # the filter reached the scan
