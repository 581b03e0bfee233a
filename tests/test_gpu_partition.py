"""GPU parity tests of Spark HashPartitioning (csrc/partition.cu) against the CPU oracle: partition ids
(pmod of the murmur3 row hash), partition offsets, the stable order inside a partition, and the moved columns
(fixed-width values, null masks and counts, STRING offsets and chars), through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from util import cols_equal, random_table

pytestmark = pytest.mark.gpu


def _gpu():
    import gpu_util
    gpu_util.require_cuda()
    return gpu_util


def _check(cols, key_idx, P, seed=42):
    G = _gpu()
    from srj_b200.partitioning import HashPartitioner
    ids = O.partition_ids([cols[i] for i in key_idx], P, seed)
    want_cols, want_offs, want_gmap = O.stable_partition(cols, ids, P)
    pt = HashPartitioner.partition(G.table_to_device(cols), key_idx, P, seed)
    assert np.array_equal(pt.partition_ids.data.view(torch.int32).cpu().numpy(), ids)
    assert pt.getPartitions() == want_offs[:-1].tolist()
    assert pt.getRowCounts() == np.diff(want_offs).tolist()
    for i, (g, w) in enumerate(zip(pt.getTable().columns, want_cols)):
        h = G.to_host(g)
        assert cols_equal(h, w, check_null_payload=w.type_id != O.STRING), f"column {i}"
        if w.type_id == O.STRING:
            assert np.array_equal(h.offsets, w.offsets) and np.array_equal(h.data, w.data), f"string column {i}"
        assert g.getNullCount() == w.null_count(), f"null count, column {i}"


SCHEMA = [O.INT32, O.INT64, O.STRING, O.DECIMAL128, O.INT8, O.FLOAT64, O.STRING, O.INT16, O.BOOL8]


@pytest.mark.parametrize("P", [1, 2, 7, 200, 1000, 5000])
@pytest.mark.parametrize("nrows", [1, 31, 33, 4097, 50_001])
def test_hash_partition_matches_oracle(nrows, P):
    cols = random_table(SCHEMA, nrows, seed=nrows + P)
    _check(cols, [0, 1], P)


def test_string_and_mixed_keys():
    cols = random_table(SCHEMA, 20_000, seed=3)
    _check(cols, [2], 64)                 # STRING key
    _check(cols, [3, 2, 4], 33)           # DECIMAL128 + STRING + INT8
    _check(cols, [0], 200, seed=0)


def test_skewed_keys_and_empty_partitions():
    """Two distinct keys over 100 K rows (most partitions empty, two huge) and an all-null key (every row hashes to the seed)."""
    n = 100_000
    rng = np.random.default_rng(1)
    k = O.HCol(O.INT32, np.where(rng.random(n) < 0.9, 5, 77).astype(np.int32).view(np.uint8))
    v = O.HCol(O.INT64, np.arange(n, dtype=np.int64).view(np.uint8))
    _check([k, v], [0], 200)
    allnull = O.HCol(O.INT32, np.zeros(n, np.int32).view(np.uint8), O.pack_mask(np.zeros(n, bool)))
    _check([allnull, v], [0], 16)


def test_no_masks_and_empty_table():
    cols = random_table([O.INT32, O.STRING, O.INT64], 10_000, seed=9, null_frac=0.0)
    _check(cols, [0], 50)
    _check(random_table([O.INT32, O.STRING], 0, seed=1), [0], 8)


def test_table_partition_by_id_column():
    """ai.rapids.cudf.Table.partition(partitionMap, n): ids given by the caller."""
    G = _gpu()
    import srj_b200 as S
    from srj_b200.partitioning import partition
    n, P = 30_000, 12
    cols = random_table([O.INT64, O.STRING], n, seed=4)
    ids = np.random.default_rng(2).integers(0, P, n).astype(np.int32)
    want_cols, want_offs, _ = O.stable_partition(cols, ids, P)
    pmap = S.ColumnVector(S.DType.INT32, n, torch.from_numpy(ids).cuda().view(torch.uint8))
    pt = partition(G.table_to_device(cols), pmap, P)
    assert pt.getPartitions() == want_offs[:-1].tolist()
    for g, w in zip(pt.getTable().columns, want_cols):
        assert cols_equal(G.to_host(g), w)


def test_partition_is_a_permutation_at_scale():
    """8 M rows, 200 partitions: size-independent properties (ids in range, offsets = histogram, every partition
    holds exactly its rows in increasing input order, the maps are inverse permutations)."""
    G = _gpu()
    import ctypes as C
    import srj_b200 as S
    from srj_b200 import _native as N
    n, P = 8_000_000, 200
    g = torch.Generator(device="cuda").manual_seed(7)
    key = torch.randint(-2**31, 2**31 - 1, (n,), dtype=torch.int32, device="cuda", generator=g)
    kc = S.ColumnVector(S.DType.INT32, n, key.view(torch.uint8))
    lib = N.lib()
    ws = torch.empty(lib.srj_partition_workspace_bytes(n, P), dtype=torch.uint8, device="cuda")
    ids = torch.empty(n, dtype=torch.int32, device="cuda")
    offs = torch.empty(P + 1, dtype=torch.int32, device="cuda")
    smap = torch.empty(n, dtype=torch.int32, device="cuda")
    gmap = torch.empty(n, dtype=torch.int32, device="cuda")
    arr = (N.SrjColumn * 1)(kc._c())
    N.check(lib.srj_hash_partition(arr, 1, n, C.c_uint32(42), P, ids.data_ptr(), offs.data_ptr(), smap.data_ptr(), gmap.data_ptr(),
                                   ws.data_ptr(), int(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    h = S.Hash.murmurHash32(42, [kc]).data.view(torch.int32)
    want_ids = torch.remainder(h.to(torch.int64), P).to(torch.int32)          # floored mod == Spark pmod
    assert torch.equal(ids, want_ids)
    assert torch.equal(offs[1:] - offs[:-1], torch.bincount(ids, minlength=P).to(torch.int32)) and int(offs[0]) == 0
    assert torch.equal(smap[gmap.long()], torch.arange(n, dtype=torch.int32, device="cuda"))      # inverse permutations
    pid_sorted = ids[gmap.long()]
    assert bool((pid_sorted[1:] >= pid_sorted[:-1]).all())                                        # grouped by partition
    same = pid_sorted[1:] == pid_sorted[:-1]
    assert bool((gmap[1:][same] > gmap[:-1][same]).all())                                         # stable inside a partition


def test_concurrent_callers_on_their_own_streams():
    """Spark task threads: 8 threads partition and UnsafeRow-convert their own tables concurrently (no shared state in
    the library: every scratch buffer is the caller's)."""
    import threading
    G = _gpu()
    import srj_b200 as S
    from srj_b200.partitioning import HashPartitioner
    from srj_b200.unsaferow import UnsafeRowConversion as UR
    from oracle import unsafe_row as U
    types = [O.INT32, O.STRING, O.INT64, O.DECIMAL128]
    errors = []

    def work(i):
        try:
            cols = random_table(types, 3000 + 17 * i, seed=100 + i)
            ids = O.partition_ids([cols[0], cols[1]], 11 + i)
            want_cols, want_offs, _ = O.stable_partition(cols, ids, 11 + i)
            uoffs, udata = U.to_unsafe_rows(cols)
            with torch.cuda.stream(torch.cuda.Stream()):
                for _ in range(5):
                    dt = G.table_to_device(cols)
                    pt = HashPartitioner.partition(dt, [0, 1], 11 + i)
                    assert pt.getPartitions() == want_offs[:-1].tolist()
                    for g, w in zip(pt.getTable().columns, want_cols):
                        assert cols_equal(G.to_host(g), w)
                    rows = UR.convertToRows(dt)
                    goffs, gdata = G.rows_to_host(rows)
                    assert np.array_equal(goffs.astype(np.int64), uoffs) and np.array_equal(gdata, udata)
        except Exception as e:  # noqa: BLE001
            errors.append((i, repr(e)))

    ts = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errors, errors
