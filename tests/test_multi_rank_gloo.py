"""N>1 path on CPU: world_size-2 gloo.  Each rank converts its own row range (oracle on the CPU; the
kernels are exercised by the -m gpu tests), the per-rank column chunks are all-gathered in rank order,
and the result must equal the single-process conversion -- including the concatenated mask words, which
is what the 32-row shard alignment guarantees."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nrows, q):
    for p in (ROOT, os.path.join(ROOT, "spark-rapids-jni_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from srj_b200 import sharding
        from util import random_table
        types = [O.INT32, O.INT64, O.FLOAT64, O.BOOL8, O.INT16]
        cols = random_table(types, nrows, seed=99)            # same seeded table on every rank
        (offs, data), = O.convert_to_rows(cols)
        row_size = int(offs[1] - offs[0]) if nrows else 0
        r0, r1 = sharding.row_range(nrows, rank, world)
        b0, b1 = sharding.rows_byte_range(None, r0, r1, row_size)
        mine, nulls = O.convert_from_rows(data[b0:b1], None, r1 - r0, types)
        # equal-sized chunks for the gather: pad the last rank's shard to the common shard height
        per = sharding.row_range(nrows, 0, world)[1]
        chunks, masks = [], []
        for c in mine:
            w = O.size_of(c.type_id)
            buf = np.zeros(per * w, np.uint8)
            buf[: (r1 - r0) * w] = np.ascontiguousarray(c.data).view(np.uint8)
            chunks.append(torch.from_numpy(buf))
            m = np.zeros(per // 32, np.uint32)
            m[: len(c.mask)] = c.mask
            masks.append(torch.from_numpy(m.view(np.int32).copy()))
        full = sharding.gather_fixed_columns(dist, chunks, world)
        fullm = sharding.gather_fixed_columns(dist, masks, world)
        # timing reduction used by bench.py: max over ranks
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = float(t[0]) == world
        want, _ = O.convert_from_rows(data, None, nrows, types)
        for c, g, gm in zip(want, full, fullm):
            w = O.size_of(c.type_id)
            ok &= np.array_equal(g.numpy()[: nrows * w], np.ascontiguousarray(c.data).view(np.uint8))
            ok &= np.array_equal(gm.numpy().view(np.uint32)[: len(c.mask)], c.mask)
        q.put((rank, bool(ok), (r0, r1)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nrows", [1000, 4096 + 17])
def test_row_range_shards_gather_equals_single_process(nrows):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nrows, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    ranges = sorted(r for _, _, r in res)
    assert ranges[0][0] == 0 and ranges[-1][1] == nrows and ranges[0][1] == ranges[1][0] and ranges[0][1] % 32 == 0


def test_row_range_properties():
    sys.path.insert(0, os.path.join(ROOT, "spark-rapids-jni_b200"))
    from srj_b200 import sharding
    for n in (0, 1, 31, 32, 33, 1000, 100_000_000):
        for w in (1, 2, 4, 8):
            rs = sharding.all_ranges(n, w)
            assert rs[0][0] == 0 and rs[-1][1] == n
            for (a0, a1), (b0, b1) in zip(rs[:-1], rs[1:]):
                assert a1 == b0 and a0 <= a1 and (a1 % 32 == 0 or a1 == n)
