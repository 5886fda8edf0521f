"""Per-step timing of the C2 query (GROUP BY key, SUM(val); 200M rows, 1M keys): where does the
'compact' phase's time go, step by step?  One GPU."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dask_sql_b200 import Context, executor

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(2)
n, nkeys = int(2e8), 1_000_000
kind = sys.argv[1] if len(sys.argv) > 1 else "float"
key = torch.randint(0, nkeys, (n,), dtype=torch.int64, device=dev, generator=g)
v = torch.rand(n, dtype=torch.float64, device=dev, generator=g) if kind == "float" else \
    torch.randint(-1000, 1001, (n,), dtype=torch.int64, device=dev, generator=g)
c = Context()
c.create_table("t", {"key": key, "val": v}, persist=True, npartitions=8)
q = "SELECT key, SUM(val) AS s FROM t GROUP BY key"
for mode in ("resolve_each", "run_ahead"):
    for _ in range(3):
        for p in executor.execute(c.sql(q), top=True):
            p.resolve()
    torch.cuda.synchronize()
    rows = []
    steps = 8
    executor.prefill_timing_events(steps * 64)
    recs = []
    for s in range(steps):
        executor.kernel_events, executor.phase_events = [], []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        parts = executor.execute(c.sql(q), top=True)
        e1.record()
        host = time.perf_counter() - t0
        if mode == "resolve_each":
            for p in parts:
                p.resolve()
        recs.append((e0, e1, host, executor.kernel_events, executor.phase_events))
    torch.cuda.synchronize()
    for s, (e0, e1, host, kev, pev) in enumerate(recs):
        ph = {}
        for name, a, b in pev:
            ph[name] = ph.get(name, 0.0) + a.elapsed_time(b)
        k = sum(a.elapsed_time(b) for _, _, a, b in kev)
        print(f"{kind} {mode} step {s}: gpu {e0.elapsed_time(e1):.3f} ms host {host*1e3:.3f} ms kernels {k:.3f} phases "
              + " ".join(f"{n_}={t:.3f}" for n_, t in ph.items()), flush=True)
executor.kernel_events = executor.phase_events = None
