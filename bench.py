#!/usr/bin/env python
"""bench.py — rows/s of the Q3-shaped filter -> join -> group-by (BASELINE.json configs[3], "C4")
on N B200s, plus roofline of the dominant kernel and the CPU baseline.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU path (pandas restatement)

Workload (seed 4): fact(fk int64, x int64, val float64) with --rows rows in total
(default 1e9, strong scaling: rows/N per GPU, 8 partitions per GPU), dim(pk, flag, grp) 10M rows,
1M groups;  SELECT d.grp, SUM(f.val) AS rev FROM fact f JOIN dim d ON f.fk = d.pk
            WHERE f.x > 0 AND d.flag < 5 GROUP BY d.grp

value : fact/dim resident in HBM, step = Context.sql(Q) (plan + plugins) + execution, result
        left on the device; timed with CUDA events, barrier + synchronize on both sides, max
        over ranks.  Inputs (24 GB) are far larger than L2 (126 MB), so no explicit L2 flush.
e2e   : same query through the public API on HOST (pinned) tables: every step copies the
        referenced fact/dim columns host->device and the result device->host (pandas).
verified_full_size (N=1): the 1e9-row result checked through size-independent properties -- sum of
        the group sums == masked sum of val (1e-9 relative), group count, key uniqueness.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

QUERY = ("SELECT d.grp, SUM(f.val) AS rev FROM fact f JOIN dim d ON f.fk = d.pk "
         "WHERE f.x > 0 AND d.flag < 5 GROUP BY d.grp")
DIM_ROWS = 10_000_000
N_GROUPS = 1_000_000
BYTES_PER_FACT_ROW = 24       # fk + x + val, each read once
TOTAL_PARTITIONS = 8      # BASELINE configs: "8 partitions"; spread over the GPUs of the run


def parts_per_gpu(world):
    return max(1, TOTAL_PARTITIONS // world)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=float, default=float(os.environ.get("B200SQL_BENCH_ROWS", 1e9)))
    ap.add_argument("--cpu-sample-rows", type=float, default=float(os.environ.get("B200SQL_CPU_SAMPLE_ROWS", 16e6)))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's path (oracle = pandas restatement; the only place bench.py runs it)
# ---------------------------------------------------------------------------------------------
def cpu_tables(rows, seed=4):
    import numpy as np
    import pandas as pd
    rng = np.random.default_rng(seed)
    dim = pd.DataFrame({"pk": rng.permutation(DIM_ROWS).astype(np.int64),
                        "flag": rng.integers(0, 10, DIM_ROWS), "grp": rng.integers(0, N_GROUPS, DIM_ROWS)})
    fact = pd.DataFrame({"fk": rng.integers(0, DIM_ROWS, rows), "x": rng.integers(-2**31, 2**31, rows),
                         "val": rng.random(rows)})
    return fact, dim


def cpu_step(fact_parts, dim, workers):
    from oracle import pandas_oracle as O
    t0 = time.perf_counter()
    out = O.c4_q3(fact_parts, dim, workers=workers)
    return time.perf_counter() - t0, len(out)


def run_cpu_baseline(sample_rows, steps=1, warmup=0):
    from oracle import pandas_oracle as O
    cores = os.cpu_count() or 1
    fact, dim = cpu_tables(int(sample_rows))
    parts = O.split(fact, max(TOTAL_PARTITIONS, cores))
    for _ in range(warmup):
        cpu_step(parts, dim, cores)
    ts = [cpu_step(parts, dim, cores)[0] for _ in range(max(1, steps))]
    t = sorted(ts)[len(ts) // 2]
    return {"value": sample_rows / t, "unit": "rows/s", "cores": cores, "kind": "port",
            "sample": f"{int(sample_rows)} fact rows of the same workload x {DIM_ROWS} dim rows, "
                      f"{len(parts)} partitions on a {cores}-thread pool, pandas restatement of "
                      "table_scan.py/join.py/aggregate.py (oracle/pandas_oracle.py c4_q3)",
            "seconds_per_step": t}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = run_cpu_baseline(args.cpu_sample_rows, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "rows/s on Q3-shaped filter->join->groupby", "value": base["value"],
        "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": base["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n):
    return {"workload": "C4: TPC-H-Q3-shaped filter->join->groupby (BASELINE.json configs[3])",
            "fact_rows_total": int(args.rows), "fact_rows_per_gpu": int(args.rows) // n, "dim_rows": DIM_ROWS,
            "groups": N_GROUPS, "partitions_per_gpu": parts_per_gpu(n), "query": QUERY,
            "l2": "inputs (24 B/row x rows) >> 126 MB L2, no flush needed",
            "planning": "Context.sql() is called every step; its plan (not its result) is served from the "
                        "prepared-statement cache after the first call; build side, lookup and group table "
                        "are rebuilt every step",
            "parallelism": f"fact sharded over {n} GPU(s); dim broadcast from rank 0 (NCCL); "
                           "dense partial aggregates all-reduced (NCCL)" if n > 1 else "single GPU"}


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi-equivalent clock / throttle-reason samples (NVML) taken only while `active`:
    the thread and NVML are brought up before the warm-up so that the first sample falls inside
    the timed region, which is only tens of milliseconds long."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.active = False
        self.nv = self.h = None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"sampler_error:{type(e).__name__}")

    def run(self):
        nv, h = self.nv, self.h
        if nv is None:
            return
        names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown"}
        try:
            while not self.stop_flag:
                if self.active:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                time.sleep(0.002)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"sampler_error:{type(e).__name__}")

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    # stdout carries exactly ONE JSON line: libraries that write to fd 1 (NCCL prints its version
    # banner there) are pointed at stderr for the whole run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from dask_sql_b200 import Context, executor

    n_total = int(args.rows)
    n_local = n_total // world
    # fit the shard into this GPU (and, for e2e, pinned host memory): shrink, loudly, if needed
    free_b, total_b = torch.cuda.mem_get_info()
    need = n_local * BYTES_PER_FACT_ROW * 1.15 + DIM_ROWS * 24 * 4
    if need > free_b:
        n_local = int((free_b - DIM_ROWS * 96) / (BYTES_PER_FACT_ROW * 1.15))
        n_total = n_local * world
        print(f"[bench] shrinking to {n_total} fact rows to fit HBM", file=sys.stderr)
    args.rows = n_total

    g = torch.Generator(device=dev)
    g.manual_seed(4 + rank)
    fk = torch.randint(0, DIM_ROWS, (n_local,), dtype=torch.int64, device=dev, generator=g)
    x = torch.randint(-2**31, 2**31, (n_local,), dtype=torch.int64, device=dev, generator=g)
    val = torch.rand(n_local, dtype=torch.float64, device=dev, generator=g)
    gd = torch.Generator(device=dev)
    gd.manual_seed(4)
    has_dim = rank == 0 or world == 1
    nd = DIM_ROWS if has_dim else 0
    pk = torch.randperm(DIM_ROWS, device=dev, generator=gd)[:nd]
    flag = torch.randint(0, 10, (DIM_ROWS,), dtype=torch.int64, device=dev, generator=gd)[:nd]
    grp = torch.randint(0, N_GROUPS, (DIM_ROWS,), dtype=torch.int64, device=dev, generator=gd)[:nd]

    c = Context()
    fact_dist = "sharded" if world > 1 else "local"
    dim_dist = "root" if world > 1 else "local"
    c.create_table("fact", {"fk": fk, "x": x, "val": val}, persist=True, npartitions=parts_per_gpu(world),
                   distribution=fact_dist)
    c.create_table("dim", {"pk": pk, "flag": flag, "grp": grp}, persist=True, distribution=dim_dist)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step_resident():
        lazy = c.sql(QUERY)
        return executor.execute(lazy)

    # ---- value: device-resident inputs
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        parts = step_resident()
    n_groups_out = parts[0].n if args.warmup else None
    barrier()
    sampler.active = True
    launches0 = executor.stats["launches"]
    executor.kernel_events = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        parts = step_resident()
    e1.record()
    barrier()
    sampler.active = False
    wall = time.perf_counter() - w0
    dev_s = e0.elapsed_time(e1) * 1e-3
    t_step = max_over_ranks(max(dev_s, 0.0)) / args.steps
    launches = executor.stats["launches"] - launches0
    kev = executor.kernel_events
    executor.kernel_events = None
    sampler.stop_flag = True
    sampler.join(timeout=2)
    n_groups_out = parts[0].n

    # ---- roofline of the dominant kernel (live CUDA events on the launching stream)
    durs = [(rows, a.elapsed_time(b) * 1e-3) for name, rows, a, b in kev if name == "b2_star_agg_kernel"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    roofline = None
    if durs:
        rows_l = sum(r for r, _ in durs) / len(durs)
        avg = sum(d for _, d in durs) / len(durs)
        achieved = rows_l * BYTES_PER_FACT_ROW / avg / 1e9
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tr["b2_star_agg_kernel"]["dram_bytes_per_row"] * rows_l
        except Exception:
            pass
        roofline = {"kernel": "b2_star_agg_kernel", "bound": "hbm", "achieved": achieved, "peak": peak_gbs,
                    "unit": "GB/s", "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": rows_l * BYTES_PER_FACT_ROW, "avg_launch_ms": avg * 1e3,
                    "launches_timed": len(durs),
                    "kernel_share_of_step": sum(d for _, d in durs) / args.steps / (dev_s / args.steps)}

    # ---- full-size parity through size-independent properties (N=1: this rank holds everything)
    verified = None
    if world == 1:
        try:
            verified = verify_full_size(torch, parts, fk, x, val, pk, flag, grp)
        except Exception as e:  # the check must never take the measurement down with it
            verified = {"error": f"{type(e).__name__}: {e}"}

    # ---- e2e: host-resident (pinned) tables through the public API, pandas result
    e2e = None
    if not args.no_e2e:
        try:
            e2e = run_e2e(args, torch, dist, dev, world, rank, fk, x, val, pk, flag, grp, fact_dist, dim_dist,
                          barrier, max_over_ranks)
        except Exception as e:  # e.g. not enough pinnable host memory
            e2e = {"error": f"{type(e).__name__}: {e}"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = run_cpu_baseline(args.cpu_sample_rows)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": "rows/s on Q3-shaped filter->join->groupby", "value": n_total / t_step, "unit": "rows/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args, world), "clocks": sampler.summary(),
            "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "groups_out": n_groups_out, "verified_full_size": verified, "wall_ms_per_step": wall / args.steps * 1e3,
            "fused_star_pipeline": executor.stats["star_fused"] > 0,
        }
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def verify_full_size(torch, parts, fk, x, val, pk, flag, grp):
    """The oracle cannot run 1e9 rows in the bench, so the full-size result is checked through
    properties that do not depend on size (plain torch ops on the resident inputs, fp64):
      * checksum of checksums: the sum over groups of SUM(val) equals the sum of val over the fact
        rows that pass both predicates (1e-9 relative, BASELINE.json north_star tolerance);
      * the number of groups equals the number of distinct grp among dim rows with flag < 5 that at
        least one passing fact row references; group keys are unique."""
    res = parts[0]
    keys, rev = res["grp"].data, res["rev"].data
    nd, dev = pk.numel(), pk.device
    ok_dim = torch.zeros(nd, dtype=torch.bool, device=dev)
    ok_dim[pk] = flag < 5                                   # indexed by key value: pk is a permutation of 0..nd-1
    grp_by_pk = torch.empty_like(grp)
    grp_by_pk[pk] = grp
    hit = torch.zeros(nd, dtype=torch.bool, device=dev)
    total = torch.zeros((), dtype=torch.float64, device=dev)
    rows = 0
    chunk = 1 << 26
    for lo in range(0, fk.numel(), chunk):
        f = fk[lo:lo + chunk]
        m = (x[lo:lo + chunk] > 0) & ok_dim[f]
        total += val[lo:lo + chunk][m].sum()
        hit[f[m]] = True
        rows += int(m.sum().item())
    groups_expected = int(torch.unique(grp_by_pk[hit]).numel())
    got, exp = float(rev.sum().item()), float(total.item())
    rel = abs(got - exp) / max(abs(exp), 1e-300)
    unique = int(torch.unique(keys).numel()) == int(keys.numel())
    return {"sum_of_group_sums_rel_err": rel, "tolerance": 1e-9, "groups": int(keys.numel()),
            "groups_expected": groups_expected, "keys_unique": unique, "rows_contributing": rows,
            "ok": bool(rel <= 1e-9 and int(keys.numel()) == groups_expected and unique)}


def run_e2e(args, torch, dist, dev, world, rank, fk, x, val, pk, flag, grp, fact_dist, dim_dist, barrier,
            max_over_ranks):
    from dask_sql_b200 import Context, executor

    host = {}
    for name, t in (("fk", fk), ("x", x), ("val", val), ("pk", pk), ("flag", flag), ("grp", grp)):
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t)
        host[name] = h
    torch.cuda.synchronize()
    c = Context()
    c.create_table("fact", {k: host[k] for k in ("fk", "x", "val")}, persist=False, npartitions=parts_per_gpu(world),
                   distribution=fact_dist)
    c.create_table("dim", {k: host[k] for k in ("pk", "flag", "grp")}, persist=False, distribution=dim_dist)
    steps = max(2, min(args.steps, 5))
    for _ in range(1):
        c.sql(QUERY, return_futures=False)
    barrier()
    h0, d0 = executor.stats["h2d_bytes"], executor.stats["d2h_bytes"]
    t0 = time.perf_counter()
    for _ in range(steps):
        out = c.sql(QUERY, return_futures=False)
    barrier()
    t = max_over_ranks(time.perf_counter() - t0) / steps
    n_total = int(args.rows)
    return {"value": n_total / t, "unit": "rows/s", "ms_per_step": t * 1e3, "steps": steps,
            "h2d_bytes_per_step": (executor.stats["h2d_bytes"] - h0) // steps,
            "d2h_bytes_per_step": (executor.stats["d2h_bytes"] - d0) // steps, "result_rows": len(out),
            "api": "Context.create_table(host pinned columns, persist=False); Context.sql(Q, return_futures=False)"}


if __name__ == "__main__":
    main()
