"""b2_peer_merge (csrc/peer.cuh) on ONE GPU: `world` buffers in the same HBM stand in for the ranks'
symmetric copies and every "rank" launches its kernel on its own stream -- the in-kernel barrier, the
rank-ordered reduction and the three ways existence is merged are then checked bit for bit against the
sequential merge the reference's tree performs (aggregate.py:575-581: sum of sums, min of mins, ...;
a group exists iff some partition saw it, tests/integration/test_groupby.py:526-598).  The real NVLink
path (torch symmetric memory, one process per GPU) is exercised by scripts/mgpu_check.py and bench.py."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu

ALIGN = 256


def _carve(buf, used, n, dtype):
    import torch
    width = torch.empty((), dtype=dtype).element_size()
    off = used
    used = off + (n * width + ALIGN - 1) // ALIGN * ALIGN
    return buf[off: off + n * width].view(dtype), off, used


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("kind", ["rows", "indicator", "bitmap"])
def test_peer_merge_on_one_gpu(world, kind):
    import torch
    from dask_sql_b200 import _lib as L

    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev)
    g.manual_seed(100 * world + len(kind))
    chunk = 32 * 37
    alloc = chunk * world
    nbytes = 5 * alloc * 8 + alloc // 8 + 16 * ALIGN
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=dev) for _ in range(world)]
    # layout (identical in every copy): signals | sumf | sumi | mn | mx | rows | bitmap
    tabs = []
    for r in range(world):
        used = 0
        sig, sig_off, used = _carve(bufs[r], used, L.MAX_PEERS, torch.int64)
        arrs, offs = {}, {}
        for name in ("sumf", "sumi", "mn", "mx", "rows"):
            dt = torch.float64 if name == "sumf" else torch.int64
            arrs[name], offs[name], used = _carve(bufs[r], used, alloc, dt)
        bitmap, bm_off, used = _carve(bufs[r], used, alloc // 32, torch.int32)
        # a rank's partial table: ~40 % of the slots touched
        hit = torch.rand(alloc, device=dev, generator=g) < 0.4
        vals = torch.rand(alloc, dtype=torch.float64, device=dev, generator=g) * 1e6 - 5e5
        ints = torch.randint(-2**62, 2**62, (alloc,), dtype=torch.int64, device=dev, generator=g)
        arrs["sumf"].copy_(torch.where(hit, vals + 0.0, torch.full_like(vals, -0.0)))
        arrs["sumi"].copy_(torch.where(hit, ints, torch.zeros_like(ints)))
        arrs["mn"].copy_(torch.where(hit, ints, torch.full_like(ints, 2**63 - 1)))
        arrs["mx"].copy_(torch.where(hit, ints, torch.full_like(ints, -2**63)))
        arrs["rows"].copy_(torch.where(hit, torch.randint(1, 9, (alloc,), device=dev, generator=g), torch.zeros_like(ints)))
        w = (hit.view(-1, 32).to(torch.int64) << torch.arange(32, device=dev)).sum(1)
        bitmap.copy_(torch.where(w >= 2**31, w - 2**32, w).to(torch.int32))
        tabs.append((arrs, hit))
    names = ["sumf", "sumi", "mn", "mx", "rows"]
    ops = [L.PEER_SUM_F64, L.PEER_SUM_I64, L.PEER_MIN_I64, L.PEER_MAX_I64, L.PEER_SUM_I64]
    ready = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    outs, press, descs = [], [], []
    for r in range(world):
        m = L.PeerMerge()
        m.world, m.rank, m.narrays = world, r, len(names)
        m.lo, m.count = r * chunk, chunk
        m.signal_off, m.bitmap_off = sig_off, bm_off
        m.local_ready = ready[r].data_ptr()
        for p in range(world):
            m.peer_base[p] = bufs[p].data_ptr()
        o = {}
        for a, (name, op) in enumerate(zip(names, ops)):
            m.ops[a], m.array_off[a] = op, offs[name]
            o[name] = torch.empty(chunk, dtype=tabs[r][0][name].dtype, device=dev)
            m.out[a] = o[name].data_ptr()
        pres = torch.full((chunk,), 7, dtype=torch.uint8, device=dev)
        m.out_present = pres.data_ptr()
        m.presence_kind = {"rows": L.PEER_PRESENT_ROWS, "indicator": L.PEER_PRESENT_INDICATOR,
                           "bitmap": L.PEER_PRESENT_BITMAP}[kind]
        m.presence_array = {"rows": 4, "indicator": 0, "bitmap": 0}[kind]
        outs.append(o)
        press.append(pres)
        descs.append(m)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
    for epoch in (1, 2):                      # twice: the signal words only ever grow
        for r in reversed(range(world)):      # last rank first: nobody may pass before rank 0 arrives
            descs[r].epoch = epoch
            L.peer_merge(C.byref(descs[r]), C.c_void_p(streams[r].cuda_stream))
        torch.cuda.synchronize()
    any_hit = torch.stack([h for _, h in tabs]).any(0)
    for r in range(world):
        sl = slice(r * chunk, (r + 1) * chunk)
        exp = {n: tabs[0][0][n][sl].clone() for n in names}
        for p in range(1, world):
            a = tabs[p][0]
            exp["sumf"] = exp["sumf"] + a["sumf"][sl]             # rank order, like the kernel
            exp["sumi"] = exp["sumi"] + a["sumi"][sl]             # int64 wraps in torch as well
            exp["mn"] = torch.minimum(exp["mn"], a["mn"][sl])
            exp["mx"] = torch.maximum(exp["mx"], a["mx"][sl])
            exp["rows"] = exp["rows"] + a["rows"][sl]
        for n in names:
            assert torch.equal(outs[r][n].view(torch.int64), exp[n].view(torch.int64)), (n, r)
        assert torch.equal(press[r], any_hit[sl].to(torch.uint8)), ("presence", kind, r)


def test_peer_merge_rejects_bad_arguments():
    import torch
    from dask_sql_b200 import _lib as L

    m = L.PeerMerge()
    m.world, m.rank = 1, 0
    with pytest.raises(L.B200SqlError):
        L.peer_merge(C.byref(m), C.c_void_p(torch.cuda.current_stream().cuda_stream))
