"""Run one BASELINE config repeatedly (for launch lists / event timing).  usage: probe_one.py c1|c2|c2i|c3|c4|c5 rows reps [nparts]  (env knobs apply)"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dask_sql_b200.frame import LazyFrame, TableSource, AggSource
from dask_sql_b200.table import DeviceTable
from dask_sql_b200 import executor

cfg, n, reps = sys.argv[1], int(float(sys.argv[2])), int(sys.argv[3])
nparts = int(sys.argv[4]) if len(sys.argv) > 4 else 8
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
g = torch.Generator(device=dev); g.manual_seed(1)
table = lambda cols, p: LazyFrame(TableSource(DeviceTable.from_columns(cols, p, dev, True)))
if cfg == "c1":
    t = table({"x": torch.randint(-2**31, 2**31, (n,), dtype=torch.int64, device=dev, generator=g)}, nparts)
    q = LazyFrame(AggSource(t[t["x"] > 0], [], [("x", "s", "sum")])); bpr = 8
elif cfg == "c2":
    t = table({"key": torch.randint(0, 1_000_000, (n,), dtype=torch.int64, device=dev, generator=g),
               "vf": torch.rand(n, dtype=torch.float64, device=dev, generator=g)}, nparts)
    q = LazyFrame(AggSource(t, ["key"], [("vf", "s", "sum")])); bpr = 16
elif cfg == "c2i":
    t = table({"key": torch.randint(0, 1_000_000, (n,), dtype=torch.int64, device=dev, generator=g),
               "vi": torch.randint(-1000, 1001, (n,), dtype=torch.int64, device=dev, generator=g)}, nparts)
    q = LazyFrame(AggSource(t, ["key"], [("vi", "s", "sum")])); bpr = 16
elif cfg == "c5":
    t = table({"key": torch.randint(0, 100_000_000, (n,), dtype=torch.int64, device=dev, generator=g),
               "val": torch.rand(n, dtype=torch.float64, device=dev, generator=g)}, nparts)
    q = LazyFrame(AggSource(t, ["key"], [("val", "s", "sum"), ("val", "m", "mean")])); bpr = 16
elif cfg == "c3":
    nd = 10_000_000
    f = table({"fk": torch.randint(0, int(nd * 1.25), (n,), dtype=torch.int64, device=dev, generator=g),
               "v": torch.rand(n, dtype=torch.float64, device=dev, generator=g)}, nparts)
    d = table({"pk": torch.randperm(nd, device=dev, generator=g),
               "w": torch.randint(0, 1000, (nd,), dtype=torch.int64, device=dev, generator=g)}, 1)
    q = f.merge(d, left_on=["fk"], right_on=["pk"], how="inner")[["fk", "v", "w"]]; bpr = 35.4
else:
    nd = 10_000_000
    f = table({"fk": torch.randint(0, nd, (n,), dtype=torch.int64, device=dev, generator=g),
               "x": torch.randint(-2**31, 2**31, (n,), dtype=torch.int64, device=dev, generator=g),
               "val": torch.rand(n, dtype=torch.float64, device=dev, generator=g)}, nparts)
    d = table({"pk": torch.randperm(nd, device=dev, generator=g),
               "flag": torch.randint(0, 10, (nd,), dtype=torch.int64, device=dev, generator=g),
               "grp": torch.randint(0, 1_000_000, (nd,), dtype=torch.int64, device=dev, generator=g)}, 1)
    j = f[f["x"] > 0].merge(d[d["flag"] < 5], left_on=["fk"], right_on=["pk"], how="inner")
    q = LazyFrame(AggSource(j, ["grp"], [("val", "rev", "sum")])); bpr = 24
for _ in range(2):
    executor.execute(q)
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter(); e0.record(); executor.execute(q); e1.record(); torch.cuda.synchronize()
    ts.append((e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3))
ts.sort()
d_ms, w_ms = ts[len(ts) // 2]
print(json.dumps({"cfg": cfg, "rows": n, "nparts": nparts, "dev_ms": round(d_ms, 3), "wall_ms": round(w_ms, 3),
                  "GBps": round(n * bpr / d_ms / 1e6, 1), "frac": round(n * bpr / d_ms / 1e6 / 6564.2, 3)}))
