"""ncu driver for the non-headline kernels: C1 (scan_agg), C2 (dense group-by), C3 (join), filter select."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dask_sql_b200.frame import LazyFrame, TableSource, AggSource
from dask_sql_b200.table import DeviceTable
from dask_sql_b200 import executor

dev = torch.device("cuda", 0); torch.cuda.set_device(0)
n = 125_000_000
g = torch.Generator(device=dev); g.manual_seed(1)
table = lambda cols: LazyFrame(TableSource(DeviceTable.from_columns(cols, 1, dev, True)))
x = torch.randint(-2**31, 2**31, (n,), dtype=torch.int64, device=dev, generator=g)
t = table({"x": x})
executor.execute(LazyFrame(AggSource(t[t["x"] > 0], [], [("x", "s", "sum")])))
key = torch.randint(0, 1_000_000, (n,), dtype=torch.int64, device=dev, generator=g)
vf = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
t2 = table({"key": key, "vf": vf})
for _ in range(2):
    executor.execute(LazyFrame(AggSource(t2, ["key"], [("vf", "s", "sum")])))
nd = 10_000_000
pk = torch.randperm(nd, device=dev, generator=g)
w = torch.randint(0, 1000, (nd,), dtype=torch.int64, device=dev, generator=g)
fk = torch.randint(0, int(nd * 1.25), (n,), dtype=torch.int64, device=dev, generator=g)
f3, d3 = table({"fk": fk, "v": vf}), table({"pk": pk, "w": w})
executor.execute(f3.merge(d3, left_on=["fk"], right_on=["pk"], how="inner")[["fk", "v", "w"]])
t4 = table({"x": x, "v": vf})
executor.execute(t4[t4["x"] > 0])
torch.cuda.synchronize()
print("done")
