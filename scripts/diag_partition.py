"""Diagnostic: is the output of b2_range_partition bucket-ordered, and what does the dense group-by
cost on ordered vs unordered input?  usage: diag_partition.py rows nkeys"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dask_sql_b200 import _lib as L, device as D
from dask_sql_b200.device import DeviceColumn, I64, F64

n, nkeys = int(float(sys.argv[1])), int(float(sys.argv[2]))
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
g = torch.Generator(device=dev); g.manual_seed(5)
key = torch.randint(0, nkeys, (n,), dtype=torch.int64, device=dev, generator=g)
val = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
nslots, kmin = nkeys + 1, 0
shift = 20
nb = ((nslots - 1) >> shift) + 1
scan = D.make_scan([DeviceColumn(key, None, I64), DeviceColumn(val, None, F64)], [], n)
ws = torch.zeros(L.range_partition_ws_bytes(nb) // 8, dtype=torch.int64, device=dev)
ok = torch.full((n,), kmin + nslots + 1, dtype=torch.int64, device=dev)
ov = torch.empty(n, dtype=torch.float64, device=dev)
cc = (C.c_int32 * 1)(1); oc = (C.c_void_p * 1)(ov.data_ptr())
L.range_partition(C.byref(scan), 0, kmin, nslots, shift, nb, 1, cc, D.ptr(ok), oc, D.ptr(ws), D.stream_ptr())
torch.cuda.synchronize()
b = (ok - kmin) >> shift
print("buckets", nb, "rows written", int(ws[nb].item()), "monotonic", bool((b[1:] >= b[:-1]).all().item()),
      "sum check", abs(float(val.sum().item()) - float(ov.sum().item())) < 1e-6 * n)


def run_ticket(kcol, vcol, label):
    table = D.GroupTable(dev, nslots + 1, [(1, L.AGG_SUM)], [F64], [False], True, False)
    sc = D.make_scan([DeviceColumn(kcol, None, I64), DeviceColumn(vcol, None, F64)], [], n)
    for _ in range(2):
        ticket = torch.zeros(1, dtype=torch.int64, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.groupby_dense_ordered(C.byref(sc), 0, kmin, table.nslots, table.aggs, len(table.specs), C.byref(table.state),
                                D.ptr(ticket), D.stream_ptr())
        e1.record(); torch.cuda.synchronize()
    print(label, "dense group-by (ticket order) ms", round(e0.elapsed_time(e1), 3))


def run(kcol, vcol, label):
    acc = torch.zeros(nslots + 1, dtype=torch.float64, device=dev)
    rows = torch.zeros(nslots + 1, dtype=torch.int64, device=dev)
    table = D.GroupTable(dev, nslots + 1, [(1, L.AGG_SUM)], [F64], [False], True, False)
    sc = D.make_scan([DeviceColumn(kcol, None, I64), DeviceColumn(vcol, None, F64)], [], n)
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); D.groupby_dense(sc, 0, kmin, table); e1.record(); torch.cuda.synchronize()
    print(label, "dense group-by ms", round(e0.elapsed_time(e1), 3))


sk, idx = torch.sort(key)
sv = val[idx]
run(key, val, "unordered")
run(ok, ov, "bucket-ordered, fixed tile stride")
run_ticket(ok, ov, "bucket-ordered")
run(sk, sv, "fully sorted")
