"""N > 1 logic of the shuffle exchange on CPU (world_size 2 and 3, gloo): every rank hash-partitions its own table with
the oracle, writes the Kudo partitions with the oracle, exchanges them with srj_b200.shuffle.exchange_partitions
(all_to_all_single) and assembles what arrived; the union over the ranks must be the union of the inputs, every row on the
rank its partition id maps to, in (source rank, partition, input order)."""
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


def _worker(rank, world, port, k, q, base_rows=500):
    os.environ["SRJ_TEST_ROWS"] = str(base_rows)
    for p in (ROOT, os.path.join(ROOT, "spark-rapids-jni_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import kudo as K
        from oracle import oracle as O
        from srj_b200.shuffle import exchange_partitions
        from util import cols_equal, random_table
        types = [O.INT32, O.STRING, O.INT64, O.DECIMAL128]
        P = world * k
        nrows = [int(os.environ.get("SRJ_TEST_ROWS", 500)) + 37 * r for r in range(world)]
        tables = [random_table(types, nrows[r], seed=50 + r) for r in range(world)]            # every rank can rebuild all inputs
        mine = tables[rank]
        ids = O.partition_ids([mine[0], mine[2]], P)
        pcols, poffs, _ = O.stable_partition(mine, ids, P)
        buf, boffs = K.split(pcols, poffs)
        recv, roffs = exchange_partitions(torch.from_numpy(buf), torch.from_numpy(boffs), k)
        got = K.assemble(recv.numpy(), roffs.numpy(), types)
        # expected: for every source rank, its rows of partitions [rank * k, (rank + 1) * k), partition-major, input order
        want_parts = []
        for r in range(world):
            rids = O.partition_ids([tables[r][0], tables[r][2]], P)
            rc, ro, _ = O.stable_partition(tables[r], rids, P)
            lo, hi = int(ro[rank * k]), int(ro[(rank + 1) * k])
            want_parts.append([O.take(c, np.arange(lo, hi)) for c in rc])
        ok = True
        for ci in range(len(types)):
            pieces = [p[ci] for p in want_parts]
            n = sum(p.size for p in pieces)
            ok &= got[ci].size == n
            at = 0
            for p in pieces:
                ok &= cols_equal(O.take(got[ci], np.arange(at, at + p.size)), p)
                at += p.size
        q.put((rank, bool(ok), int(got[0].size)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,k,base_rows", [(2, 1, 500), (2, 3, 500), (3, 2, 500), (2, 16, 3)])     # last: more partitions than rows (empty partitions)
def test_shuffle_exchange_over_gloo(world, k, base_rows):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, k, q, base_rows)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=180) for _ in range(world)]
    [p.join(timeout=60) for p in procs]
    assert all(ok for _, ok, _ in res), res
    assert sum(n for _, _, n in res) == sum(base_rows + 37 * r for r in range(world))      # every row arrived exactly once
