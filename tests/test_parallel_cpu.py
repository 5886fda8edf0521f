"""Multi-rank plumbing on CPU (gloo, world_size 2 and 3): sharding, the merge-tree schedule, the
build-side broadcast / all-gather, and the partial-aggregate tree merge protocol.  The merge
arithmetic itself is a CUDA kernel (no CPU fallback), so here the receiver-side combine is the
oracle's (tests may use the oracle); what is under test is who-sends-what-to-whom."""
import os
import socket

import numpy as np
import pandas as pd
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_bounds_cover_rows_without_overlap():
    from dask_sql_b200.parallel import shard_bounds
    for n in (0, 1, 31, 32, 1000, 10**9 + 7):
        for size in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, size) for r in range(size)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            for lo, hi in spans:
                assert lo % 32 == 0 or lo == hi      # non-empty shards start on a bitmap word


@pytest.mark.parametrize("size", [1, 2, 3, 4, 8, 13])
@pytest.mark.parametrize("fan_in", [2, 3, 8])
def test_tree_rounds_reduce_everything_onto_rank0(size, fan_in):
    from dask_sql_b200.parallel import tree_rounds
    holds = {r: {r} for r in range(size)}
    alive = set(range(size))
    for rnd in tree_rounds(size, fan_in):
        receivers = [r for r, _ in rnd]
        senders = [s for _, s in rnd]
        assert len(set(senders)) == len(senders) and not set(senders) & set(receivers)
        assert len(set(receivers)) == len(receivers)      # one message per receiver per round
        for r, s in rnd:
            assert r in alive and s in alive
            holds[r] |= holds.pop(s)
            alive.discard(s)
    assert alive == {0} and holds[0] == set(range(size))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, size, port, fan_in, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        from dask_sql_b200 import merge
        from dask_sql_b200.device import DeviceColumn, I64, F64
        from dask_sql_b200.executor import Part, RawGroups
        dev = torch.device("cpu")

        # --- broadcast of a build side that only rank 0 holds
        if rank == 0:
            part = Part({"pk": DeviceColumn(torch.arange(10, dtype=torch.int64), None, I64),
                         "w": DeviceColumn(torch.arange(10, dtype=torch.float64) * 0.5, None, F64)}, 10)
        else:
            part = Part({}, 0)
        got = merge.broadcast_part(part, dev, src=0)
        assert got.n == 10 and got["pk"].data.tolist() == list(range(10))
        assert got["w"].data.tolist() == [i * 0.5 for i in range(10)]

        # --- all-gather of a pre-sharded build side
        mine = Part({"k": DeviceColumn(torch.arange(rank * 4, rank * 4 + 4, dtype=torch.int64), None, I64)}, 4)
        allp = merge.allgather_part(mine, dev)
        assert allp["k"].data.tolist() == list(range(4 * size))

        # --- tree merge of partial group tables (receiver combine = oracle)
        rng = np.random.default_rng(100 + rank)
        keys = rng.choice(50, size=20, replace=False).astype(np.int64)
        sums = rng.random(20)
        cnts = rng.integers(1, 9, 20).astype(np.int64)

        class KA:
            def __init__(self, op):
                self.op = op

        class Plan:
            kaggs = [KA(0)]

        def oracle_merge(parts, plan, nkeys):
            df = pd.concat([pd.DataFrame({n: p[n].data.numpy() for n in p}) for p in parts])
            m = df.groupby("k0", as_index=False).sum()
            return Part({"k0": DeviceColumn(torch.from_numpy(m["k0"].to_numpy()), None, I64),
                         "a0": DeviceColumn(torch.from_numpy(m["a0"].to_numpy()), None, F64),
                         "c0": DeviceColumn(torch.from_numpy(m["c0"].to_numpy()), None, I64)}, len(m))

        merge.merge_partials = oracle_merge
        raw = RawGroups({"key": DeviceColumn(torch.from_numpy(keys), None, I64)},
                        [DeviceColumn(torch.from_numpy(sums), None, F64)],
                        [DeviceColumn(torch.from_numpy(cnts), None, I64)], None, 20)
        merged = merge.tree_merge_raw(raw, Plan(), {"split_every": fan_in}, dev)
        res = pd.DataFrame({"k": merged.keys["key"].data.numpy(), "s": merged.acc[0].data.numpy(),
                            "c": merged.cnt[0].data.numpy()}).sort_values("k").reset_index(drop=True)
        out_q.put((rank, keys, sums, cnts, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("size,fan_in", [(2, 2), (3, 2), (3, 8)])
def test_gloo_broadcast_allgather_and_tree_merge(size, fan_in):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, size, port, fan_in, q)) for r in range(size)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(size)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    everything = pd.concat([pd.DataFrame({"k": k, "s": s, "c": c}) for _, k, s, c, _ in results])
    exp = everything.groupby("k", as_index=False).sum().sort_values("k").reset_index(drop=True)
    for _, _, _, _, res in results:           # every rank ends with the full merged table
        assert res["k"].tolist() == exp["k"].tolist()
        assert res["c"].tolist() == exp["c"].tolist()
        np.testing.assert_allclose(res["s"].to_numpy(), exp["s"].to_numpy(), rtol=1e-12)


def _dense_worker(rank, size, port, out_q):
    """Dense partial tables -> reduce-scatter by slot range.  The presence derivation is a CUDA
    pass in the product (b2_expr_eval); here it is restated with torch ops so that the PROTOCOL
    (who ends up owning which slots, how existence travels) runs on gloo."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        from dask_sql_b200 import executor as X, parallel as P, merge
        from dask_sql_b200.device import DeviceColumn, I64
        from dask_sql_b200.executor import Part
        dev = torch.device("cpu")

        # --- helpers
        assert P.all_gather_ints([rank, 10 * rank], dev) == [[r, 10 * r] for r in range(size)]
        counts = [r + 1 for r in range(size)]
        mine = torch.arange(counts[rank], dtype=torch.int64) + 100 * rank
        got = P.all_gather_varlen(mine, counts)
        assert got.tolist() == [100 * r + i for r in range(size) for i in range(r + 1)]
        t = torch.arange(4 * size, dtype=torch.float64) * (rank + 1)
        own = P.reduce_scatter_(t.clone(), "sum")
        tot = sum(r + 1 for r in range(size))
        assert own.tolist() == [float(i * tot) for i in range(4 * rank, 4 * rank + 4)]

        # --- uneven all-gather of a partition (rank r contributes r rows; rank 0 none)
        p = Part({"k": DeviceColumn(torch.arange(rank, dtype=torch.int64) + 1000 * rank, None, I64)}, rank)
        allp = merge.allgather_part(p, dev)
        assert allp["k"].data.tolist() == [1000 * r + i for r in range(size) for i in range(r)]

        # --- "can this aggregate input be NULL?" is agreed across the ranks of a sharded plan (it decides
        # which accumulator arrays a table carries; partial tables are merged array by array): NULL on
        # any rank = nullable on every rank, and every rank gets the same answer
        from dask_sql_b200.expr import Lit
        from dask_sql_b200.frame import LazyFrame

        class Src:
            pass

        X._dev = lambda: dev
        frame = LazyFrame.__new__(LazyFrame)
        frame.source = Src()
        agreed = X._nullable_fn(frame, sharded=True)
        local = X._nullable_fn(frame, sharded=False)
        mine_null = Lit(None, I64) if rank == size - 1 else Lit(1, I64)
        assert local(mine_null) == (rank == size - 1)
        assert agreed(mine_null) is True                      # the last rank's NULL makes it nullable everywhere
        assert agreed(Lit(2, I64)) is False

        # --- the dense merge: 1000 logical slots (+ NULL slot), float SUM whose accumulator doubles
        # as the local presence flag (-0.0 = untouched); 30 % of the slots are hit by NO rank
        nslots = 1001
        alloc = X._padded_slots(nslots, True)
        assert alloc % (32 * size) == 0 and alloc >= nslots
        rng = np.random.default_rng(5)
        ever = rng.random(nslots) < 0.7
        ever[-1] = False                                     # the NULL slot stays empty
        rng_r = np.random.default_rng(50 + rank)
        hit = ever & (rng_r.random(nslots) < 0.5)
        vals = np.where(hit, rng_r.random(nslots), 0.0)

        class T:
            pass

        tbl = T()
        tbl.nslots, tbl.alloc, tbl.indicator, tbl.rows, tbl.present = nslots, alloc, 0, None, None
        acc = torch.full((alloc,), -0.0, dtype=torch.float64)
        acc[:nslots][torch.from_numpy(hit)] = torch.from_numpy(vals[hit] + 0.0)
        tbl.acc, tbl.cnt = [acc], [None]

        class KA:
            op = 0

        class Plan:
            kaggs = [KA()]

        X._presence_bytes = lambda t, d, out=None: (t.acc[0].view(torch.int64) != -(1 << 63)).to(torch.uint8)
        view = X._merge_dense(tbl, Plan(), True, dev)
        assert view.dist == "keyrange" and view.count == alloc // size and view.lo == rank * view.count
        assert view.occ_kind == "bytes"
        out_q.put((rank, hit, vals, view.lo, view.occ.numpy().copy(), view.acc[0].numpy().copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("size", [2, 3])
def test_gloo_dense_merge_keeps_group_existence(size):
    """Groups that no rank saw must not appear after the merge (round-1 N=8 failure: existence was
    inferred from a -0.0 that the collective did not preserve), groups any rank saw must."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dense_worker, args=(r, size, port, q)) for r in range(size)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=120) for _ in range(size)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    nslots = 1001
    any_hit = np.zeros(nslots, bool)
    total = np.zeros(nslots)
    for _, hit, vals, _, _, _ in results:
        any_hit |= hit
        total += vals
    pres = np.concatenate([occ for *_, occ, _ in results])[:nslots]
    sums = np.concatenate([acc for *_, acc in results])[:nslots]
    los = [lo for _, _, _, lo, _, _ in results]
    assert los == sorted(los) and los[0] == 0
    assert (pres.astype(bool) == any_hit).all()
    assert 0 < any_hit.sum() < nslots - 200            # the test really has never-hit groups
    np.testing.assert_allclose(sums[any_hit], total[any_hit], rtol=1e-12)
