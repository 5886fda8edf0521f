"""Quick device-timed probe of the BASELINE configs at frame level (not the bench contract)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dask_sql_b200.frame import LazyFrame, TableSource, AggSource
from dask_sql_b200.table import DeviceTable
from dask_sql_b200 import executor

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
PEAK = 6564.2e9
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0


def table(cols, nparts):
    return LazyFrame(TableSource(DeviceTable.from_columns(cols, nparts, dev, True)))


def timeit(name, fn, bytes_, rows, reps=5):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append((e0.elapsed_time(e1) * 1e-3, time.perf_counter() - w0))
    t = sorted(x[0] for x in ts)[len(ts) // 2]
    w = sorted(x[1] for x in ts)[len(ts) // 2]
    print(json.dumps({"cfg": name, "dev_ms": round(t * 1e3, 3), "wall_ms": round(w * 1e3, 3),
                      "GBps": round(bytes_ / t / 1e9, 1), "frac": round(bytes_ / t / PEAK, 3),
                      "Grows_s": round(rows / t / 1e9, 2)}), flush=True)


g = torch.Generator(device=dev); g.manual_seed(1)
# C1
for n in (10_000_000, int(1_000_000_000 * scale)):
    x = torch.randint(-2**31, 2**31, (n,), dtype=torch.int64, device=dev, generator=g)
    t = table({"x": x}, 1 if n <= 10_000_000 else 8)
    q = LazyFrame(AggSource(t[t["x"] > 0], [], [("x", "s", "sum")]))
    timeit(f"C1 n={n}", lambda: executor.execute(q), n * 8, n)
    del x, t, q
torch.cuda.empty_cache()
# C2
n = int(200_000_000 * scale)
key = torch.randint(0, 1_000_000, (n,), dtype=torch.int64, device=dev, generator=g)
vi = torch.randint(-1000, 1001, (n,), dtype=torch.int64, device=dev, generator=g)
vf = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
t = table({"key": key, "vi": vi, "vf": vf}, 8)
q = LazyFrame(AggSource(t, ["key"], [("vi", "s", "sum")]))
timeit("C2 int dense", lambda: executor.execute(q), n * 16, n)
q = LazyFrame(AggSource(t, ["key"], [("vf", "s", "sum")]))
timeit("C2 float dense", lambda: executor.execute(q), n * 16, n)
executor.DENSE_MAX_SLOTS = 0
q = LazyFrame(AggSource(t, ["key"], [("vf", "s", "sum")]))
timeit("C2 float hash1", lambda: executor.execute(q), n * 16, n)
executor.DENSE_MAX_SLOTS = 1 << 27
del key, vi, vf, t, q
torch.cuda.empty_cache()
# C4
n = int(1_000_000_000 * scale); nd = 10_000_000
fk = torch.randint(0, nd, (n,), dtype=torch.int64, device=dev, generator=g)
x = torch.randint(-2**31, 2**31, (n,), dtype=torch.int64, device=dev, generator=g)
val = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
pk = torch.randperm(nd, device=dev, generator=g)
flag = torch.randint(0, 10, (nd,), dtype=torch.int64, device=dev, generator=g)
grp = torch.randint(0, 1_000_000, (nd,), dtype=torch.int64, device=dev, generator=g)
f = table({"fk": fk, "x": x, "val": val}, 8)
d = table({"pk": pk, "flag": flag, "grp": grp}, 1)
j = f[f["x"] > 0].merge(d[d["flag"] < 5], left_on=["fk"], right_on=["pk"], how="inner")
q = LazyFrame(AggSource(j, ["grp"], [("val", "rev", "sum")]))
timeit("C4 q3 fused", lambda: executor.execute(q), n * 24 + nd * 24, n)
print("stats", executor.stats)
del x, val
torch.cuda.empty_cache()
# C3 join materialised
w = torch.randint(0, 1000, (nd,), dtype=torch.int64, device=dev, generator=g)
v = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
fk3 = torch.randint(0, int(nd * 1.25), (n,), dtype=torch.int64, device=dev, generator=g)
f3 = table({"fk": fk3, "v": v}, 8)
d3 = table({"pk": pk, "w": w}, 1)
q = f3.merge(d3, left_on=["fk"], right_on=["pk"], how="inner")[["fk", "v", "w"]]
timeit("C3 join mat", lambda: executor.execute(q), n * 16 + nd * 16 + int(0.8 * n) * 24, n, reps=3)
