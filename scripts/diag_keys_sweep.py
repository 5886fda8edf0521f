"""Throughput of b2_groupby_dense (float SUM, one accumulator array) vs number of distinct keys:
where does the L2-resident atomic rate end?  usage: diag_keys_sweep.py rows"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dask_sql_b200 import _lib as L, device as D
from dask_sql_b200.device import DeviceColumn, I64, F64

n = int(float(sys.argv[1]))
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
g = torch.Generator(device=dev); g.manual_seed(5)
val = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
for nkeys in (1 << 20, 1 << 21, 1 << 22, 1 << 23, 1 << 24, 1 << 25, 1 << 26):
    key = torch.randint(0, nkeys, (n,), dtype=torch.int64, device=dev, generator=g)
    table = D.GroupTable(dev, nkeys + 1, [(1, L.AGG_SUM)], [F64], [False], False, False)
    sc = D.make_scan([DeviceColumn(key, None, I64), DeviceColumn(val, None, F64)], [], n)
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); D.groupby_dense(sc, 0, 0, table); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"keys {nkeys:>9d} table {nkeys * 8 >> 20:4d} MB  {ms:7.3f} ms  {n / ms / 1e6:7.1f} G rows/s")
