"""C5 on N GPUs (SURVEY 8d / BASELINE configs[4]): GROUP BY over 100M keys, SUM + AVG, 500M rows per GPU.
Every rank range-partitions and aggregates its shard into a dense partial table (1.6 GB of accumulators),
the partial tables are reduce-scattered by key range (NCCL; each rank ends up owning the merged groups of
one contiguous range = split_out = world) and compacted per rank.  Prints one JSON line from rank 0 with
the step time (CUDA events, max over ranks), where it goes, and the full-size verification.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29515 scripts/mgpu_c5.py [--rows-per-gpu 5e8] [--keys 1e8] [--steps 3]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from dask_sql_b200 import Context, executor


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows-per-gpu", type=float, default=5e8)
    ap.add_argument("--keys", type=float, default=1e8)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, nkeys = int(args.rows_per_gpu), int(args.keys)
    g = torch.Generator(device=dev)
    g.manual_seed(5 + rank)
    key = torch.randint(0, nkeys, (n,), dtype=torch.int64, device=dev, generator=g)
    val = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
    c = Context()
    c.create_table("t", {"key": key, "val": val}, persist=True, npartitions=8, distribution="sharded")
    q = "SELECT key, SUM(val) AS s, AVG(val) AS a FROM t GROUP BY key"

    def step():
        return executor.execute(c.sql(q), top=True)

    for _ in range(args.warmup):
        parts = step()
    torch.cuda.synchronize()
    dist.barrier()
    executor.prefill_timing_events(args.steps * 96)
    executor.kernel_events, executor.phase_events = [], []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        parts = step()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    kev, pev = executor.kernel_events, executor.phase_events
    executor.kernel_events = executor.phase_events = None
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    where = {}
    for name, _, a, b in kev:
        where[name] = where.get(name, 0.0) + a.elapsed_time(b) / args.steps
    for name, a, b in pev:
        where["phase:" + name] = where.get("phase:" + name, 0.0) + a.elapsed_time(b) / args.steps

    # ---- verification through all-reduced invariants
    res = parts[0]
    k_out, s_out, a_out = res["key"].data, res["s"].data, res["a"].data
    cnt = torch.bincount(key, minlength=nkeys)
    dist.all_reduce(cnt)
    tot = val.sum().reshape(1)
    dist.all_reduce(tot)
    got = torch.stack([s_out.sum(), torch.tensor(float(k_out.numel()), dtype=torch.float64, device=dev)])
    dist.all_reduce(got)
    groups_expected = int((cnt > 0).sum().item())
    rel = abs(float(got[0]) - float(tot[0])) / max(abs(float(tot[0])), 1e-300)
    # this rank's groups: exactly the present keys of its range, AVG = SUM / count
    lo = int(k_out.min().item()) if k_out.numel() else 0
    hi = int(k_out.max().item()) if k_out.numel() else -1
    mine = torch.nonzero(cnt[lo:hi + 1] > 0).reshape(-1) + lo
    keys_ok = bool(torch.equal(k_out, mine))
    avg_ok = False
    if keys_ok and k_out.numel():
        ea = s_out / cnt[k_out].double()
        avg_ok = float(((a_out - ea).abs() / ea.abs().clamp_min(1e-300)).max().item()) <= 1e-9
    spans = [None] * world
    dist.all_gather_object(spans, (lo, hi, int(k_out.numel())))
    disjoint = all(spans[i][1] < spans[i + 1][0] for i in range(world - 1) if spans[i][2] and spans[i + 1][2])
    ok_t = torch.tensor([1 if (keys_ok and avg_ok) else 0], dtype=torch.int64, device=dev)
    dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
    ok = bool(int(ok_t)) and disjoint and rel <= 1e-9 and int(round(float(got[1]))) == groups_expected
    if rank == 0:
        total_rows = n * world
        print(json.dumps({
            "config": "C5: GROUP BY key (100M keys) SUM+AVG, sharded rows, dense partial tables reduce-scattered by key range",
            "n_gpus": world, "rows_total": total_rows, "rows_per_gpu": n, "keys": nkeys, "ms_per_step": ms,
            "rows_per_s": total_rows / (ms * 1e-3), "steps": args.steps, "warmup": args.warmup,
            "where_ms_per_step_rank0": {k: round(v, 3) for k, v in where.items()},
            "partitioned_groupby": executor.stats["partitioned_groupby"] > 0,
            "verified": {"ok": ok, "sum_of_sums_rel_err": rel, "groups": int(round(float(got[1]))),
                         "groups_expected": groups_expected, "key_ranges_disjoint": disjoint,
                         "per_rank_keys_and_avg_exact": bool(int(ok_t))}}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
