"""Host-side cost of one bench step (Context.sql + execute) on a small table: cProfile of 300 steps.
usage: pyprofile_step.py [rows]"""
import cProfile, pstats, sys, os, time, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dask_sql_b200 import Context, executor
import bench

rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 8_000_000
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
g = torch.Generator(device=dev); g.manual_seed(4)
nd = bench.DIM_ROWS
c = Context()
c.create_table("fact", {"fk": torch.randint(0, nd, (rows,), dtype=torch.int64, device=dev, generator=g),
                        "x": torch.randint(-2**31, 2**31, (rows,), dtype=torch.int64, device=dev, generator=g),
                        "val": torch.rand(rows, dtype=torch.float64, device=dev, generator=g)}, persist=True, npartitions=8)
c.create_table("dim", {"pk": torch.randperm(nd, device=dev, generator=g),
                       "flag": torch.randint(0, 10, (nd,), dtype=torch.int64, device=dev, generator=g),
                       "grp": torch.randint(0, bench.N_GROUPS, (nd,), dtype=torch.int64, device=dev, generator=g)}, persist=True)
step = lambda: executor.execute(c.sql(bench.QUERY))
for _ in range(5):
    parts = step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300):
    parts = step()
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"host issue time/step {t_issue / 300 * 1e3:.3f} ms   wall/step incl. GPU drain {t_all / 300 * 1e3:.3f} ms   groups {parts[0].n}")
pr = cProfile.Profile(); pr.enable()
for _ in range(300):
    parts = step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(25); print(s.getvalue()[:5000])
