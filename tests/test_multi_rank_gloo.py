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


# ---------------------------------------------------------------------------------------------------
# packed slab + ONE all-gather + STRING offsets rebase (the multi-GPU configuration's host logic), world 2 on gloo
# ---------------------------------------------------------------------------------------------------
def _slab_worker(rank, world, port, nrows, q):
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
        types = [O.INT32, O.STRING, O.INT64, O.DECIMAL128, O.STRING, O.BOOL8]
        cols = random_table(types, nrows, seed=5)             # same seeded table on every rank
        (offs, data), = O.convert_to_rows(cols)
        per = nrows // world
        assert per % 32 == 0 and per * world == nrows
        r0, r1 = rank * per, (rank + 1) * per
        b0, b1 = sharding.rows_byte_range(offs, r0, r1)
        mine, _ = O.convert_from_rows(data[b0:b1], (offs[r0:r1 + 1] - offs[r0]).astype(np.int32), per, types)
        want, _ = O.convert_from_rows(data, offs, nrows, types)
        cap = max(sum((len(c.data) + 15) & ~15 for c in want if c.type_id == O.STRING), 16)
        lay = sharding.SlabLayout([0 if t == O.STRING else O.size_of(t) for t in types], per, cap)
        slab = np.zeros(lay.nbytes, np.uint8)
        totals = np.zeros(len(types) + 1, np.int64)
        sizes = [len(c.data) for c in mine if c.type_id == O.STRING]
        at_chars = iter(lay.chars_offsets(sizes))
        for i, c in enumerate(mine):
            if c.type_id == O.STRING:
                slab[lay.at_data[i]: lay.at_data[i] + (per + 1) * 4] = c.offsets.view(np.uint8)
                a = next(at_chars)
                slab[a: a + len(c.data)] = c.data
                totals[i] = len(c.data)
            else:
                raw = np.ascontiguousarray(c.data).view(np.uint8)
                slab[lay.at_data[i]: lay.at_data[i] + len(raw)] = raw
            slab[lay.at_mask[i]: lay.at_mask[i] + lay.words * 4] = c.mask.view(np.uint8)
        slab[lay.at_totals: lay.at_totals + len(totals) * 8] = totals.view(np.uint8)
        gathered, _ = sharding.gather_slab(dist, torch.from_numpy(slab), world)       # ONE collective
        g = gathered.numpy()
        tot_all = sharding.rebase_offsets_host(g, lay, world)
        g2 = g.reshape(world, lay.nbytes)
        ok = True
        for i, c in enumerate(want):
            m = np.concatenate([g2[r, lay.at_mask[i]: lay.at_mask[i] + lay.words * 4].view(np.uint32) for r in range(world)])
            ok &= np.array_equal(m, c.mask)
            if c.type_id == O.STRING:
                o = [g2[r, lay.at_data[i]: lay.at_data[i] + (per + 1) * 4].view(np.int32) for r in range(world)]
                glob = np.concatenate([o[0]] + [x[1:] for x in o[1:]])
                ok &= np.array_equal(glob, c.offsets)
                ok &= all(o[r][0] == o[r - 1][-1] for r in range(1, world))       # chunk r starts where r-1 ended
                chars = []
                for r in range(world):
                    sz = [int(tot_all[r, j]) for j in lay.string_cols]
                    a = lay.chars_offsets(sz)[lay.string_cols.index(i)]
                    chars.append(g2[r, a: a + int(tot_all[r, i])])
                ok &= np.array_equal(np.concatenate(chars), c.data)
            else:
                w = O.size_of(c.type_id)
                d = np.concatenate([g2[r, lay.at_data[i]: lay.at_data[i] + per * w] for r in range(world)])
                ok &= np.array_equal(d, np.ascontiguousarray(c.data).view(np.uint8))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_packed_slab_one_allgather_with_strings():
    world, nrows = 2, 1024
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_slab_worker, args=(r, world, port, nrows, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)
