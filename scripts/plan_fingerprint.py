"""Host-only fingerprint of what the planner + plugins turn a set of queries into (lazy-frame trees,
temporary names normalised).  Run before and after touching planner/ or physical/ and `cmp` the two
JSON files: identical output means the executor sees exactly the same work.
usage: plan_fingerprint.py out.json"""
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pandas as pd, numpy as np
from dask_sql_b200 import Context
from dask_sql_b200.frame import LazyFrame, TableSource, JoinSource, AggSource, SortSource, LimitSource

UUID = re.compile(r"[0-9a-f]{8}-[0-9a-f]{4}-[0-9a-f]{4}-[0-9a-f]{4}-[0-9a-f]{12}")
def norm(s):
    seen = {}
    return UUID.sub(lambda m: seen.setdefault(m.group(0), f"tmp{len(seen)}"), s)
def describe(f):
    src = f.source
    d = {"cols": f.columns, "exprs": {k: repr(v) for k, v in f.exprs.items()}, "pred": [repr(p) for p in f.pred], "src": type(src).__name__}
    if isinstance(src, JoinSource):
        d["how"], d["on"], d["l"], d["r"] = src.how, [src.left_on, src.right_on], describe(src.left), describe(src.right)
    elif isinstance(src, AggSource):
        d["g"], d["aggs"], d["child"] = src.group_cols, [list(map(str, a)) for a in src.aggs], describe(src.child)
    elif isinstance(src, (SortSource, LimitSource)):
        d["child"] = describe(src.child)
        d["extra"] = str(getattr(src, "keys", None)) + str(getattr(src, "fetch", None)) + str(getattr(src, "offset", None))
    elif isinstance(src, TableSource):
        d["table"] = src.table.name
    return d
c = Context()
rng = np.random.default_rng(0)
c.create_table("t", pd.DataFrame({"a": rng.integers(0, 9, 50), "b": rng.random(50), "k": rng.integers(0, 3, 50),
                                  "n": pd.array(rng.integers(0, 5, 50), dtype="Int64"), "f32": rng.random(50).astype(np.float32)}))
c.create_table("u", pd.DataFrame({"k": [0, 1, 2], "w": [1.0, 2.0, 3.0], "flag": [True, False, True]}))
Q = ["SELECT * FROM t", "SELECT a FROM t WHERE b > 0.5 AND a < 4", "SELECT a + 1 AS a1, b FROM t WHERE n IS NOT NULL",
     "SELECT k, SUM(b) AS s, COUNT(*) AS c, AVG(a) AS m FROM t WHERE a > 2 GROUP BY k",
     "SELECT t.k, u.w FROM t JOIN u ON t.k = u.k WHERE t.b > 0.1 AND u.flag",
     "SELECT u.w, SUM(t.b) AS s FROM t JOIN u ON t.k = u.k WHERE t.a > 1 AND u.w < 3 GROUP BY u.w",
     "SELECT t.a, u.w FROM t LEFT JOIN u ON t.k = u.k AND u.w > t.b", "SELECT DISTINCT k, a FROM t WHERE a <> 3",
     "SELECT a, b FROM t ORDER BY a DESC, b LIMIT 5", "SELECT x.a FROM (SELECT a, k FROM t WHERE b < 0.9) x WHERE x.k = 1",
     "SELECT CAST(a AS DOUBLE) AS d, CAST(f32 AS BIGINT) AS i FROM t", "SELECT k FROM t WHERE a IN (1, 2, 3) AND b BETWEEN 0.2 AND 0.8",
     "SELECT t.k FROM t WHERE t.k IN (SELECT k FROM u WHERE w > 1)" , "SELECT SUM(a) FILTER (WHERE b > 0.5) AS s FROM t",
     "WITH q AS (SELECT k, MAX(b) AS mb FROM t GROUP BY k) SELECT q.k, q.mb, u.w FROM q JOIN u ON q.k = u.k"]
out = {}
for q in Q:
    try:
        out[q] = json.loads(norm(json.dumps(describe(c.sql(q)), sort_keys=True)))
    except Exception as e:
        out[q] = f"{type(e).__name__}: {e}"
json.dump(out, open(sys.argv[1], "w"), indent=1, sort_keys=True)
print(sum(isinstance(v, dict) for v in out.values()), "planned of", len(Q))
