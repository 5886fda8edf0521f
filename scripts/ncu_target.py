"""Small driver for ncu captures: one C1, one C2 (dense, float), one C4 (fused star) query."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dask_sql_b200.frame import LazyFrame, TableSource, AggSource
from dask_sql_b200.table import DeviceTable
from dask_sql_b200 import executor

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
which = sys.argv[2] if len(sys.argv) > 2 else "c1,c2,c4"
g = torch.Generator(device=dev); g.manual_seed(1)


def table(cols, nparts):
    return LazyFrame(TableSource(DeviceTable.from_columns(cols, nparts, dev, True)))


if "c1" in which:
    x = torch.randint(-2**31, 2**31, (n,), dtype=torch.int64, device=dev, generator=g)
    t = table({"x": x}, 1)
    q = LazyFrame(AggSource(t[t["x"] > 0], [], [("x", "s", "sum")]))
    for _ in range(2):
        executor.execute(q)
if "c2" in which:
    key = torch.randint(0, 1_000_000, (n,), dtype=torch.int64, device=dev, generator=g)
    vf = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
    t = table({"key": key, "vf": vf}, 1)
    q = LazyFrame(AggSource(t, ["key"], [("vf", "s", "sum")]))
    for _ in range(2):
        executor.execute(q)
if "c4" in which:
    nd = 10_000_000
    fk = torch.randint(0, nd, (n,), dtype=torch.int64, device=dev, generator=g)
    x = torch.randint(-2**31, 2**31, (n,), dtype=torch.int64, device=dev, generator=g)
    val = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
    pk = torch.randperm(nd, device=dev, generator=g)
    flag = torch.randint(0, 10, (nd,), dtype=torch.int64, device=dev, generator=g)
    grp = torch.randint(0, 1_000_000, (nd,), dtype=torch.int64, device=dev, generator=g)
    f = table({"fk": fk, "x": x, "val": val}, 1)
    d = table({"pk": pk, "flag": flag, "grp": grp}, 1)
    j = f[f["x"] > 0].merge(d[d["flag"] < 5], left_on=["fk"], right_on=["pk"], how="inner")
    q = LazyFrame(AggSource(j, ["grp"], [("val", "rev", "sum")]))
    for _ in range(2):
        executor.execute(q)
torch.cuda.synchronize()
print("done")
