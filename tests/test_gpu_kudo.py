"""GPU parity tests of the Kudo split / assemble kernels (csrc/kudo.cu) against the CPU restatement of the wire format
(oracle/kudo.py): the split buffer and the partition offsets bit-exact, assemble(split(x)) = x, partitions of several
splits assembled together, a hash-partitioned table shipped partition by partition."""
import numpy as np
import pytest
import torch

from oracle import kudo as K
from oracle import oracle as O
from util import cols_equal, random_table

pytestmark = pytest.mark.gpu

TYPES = [O.INT32, O.STRING, O.INT64, O.DECIMAL128, O.INT8, O.STRING, O.FLOAT64, O.INT16, O.BOOL8]


def _gpu():
    import gpu_util
    gpu_util.require_cuda()
    return gpu_util


def _split_and_check(cols, splits):
    G = _gpu()
    from srj_b200.kudo import KudoGpuSerializer as KS
    want_buf, want_offs = K.split(cols, splits)
    buf, offs = KS.splitAndSerializeToDevice(G.table_to_device(cols), splits)
    assert np.array_equal(offs.cpu().numpy(), want_offs)
    got = buf.cpu().numpy()
    assert np.array_equal(got, want_buf), f"first diff at byte {np.flatnonzero(got != want_buf)[:5]} of {len(want_buf)}"
    return buf, offs


@pytest.mark.parametrize("nrows", [1, 7, 8, 9, 33, 1000, 20_000])
def test_split_bytes_match_oracle(nrows):
    cols = random_table(TYPES, nrows, seed=nrows)
    rng = np.random.default_rng(nrows)
    cuts = sorted(rng.integers(0, nrows + 1, 6).tolist())
    _split_and_check(cols, [0] + cuts + [nrows])          # empty partitions, unaligned starts
    _split_and_check(cols, [0, nrows])


def test_no_masks_and_single_column():
    cols = random_table([O.INT64, O.STRING], 5000, seed=2, null_frac=0.0)
    _split_and_check(cols, [0, 1, 2500, 4999, 5000])
    _split_and_check(random_table([O.STRING], 300, seed=3), [0, 100, 300])
    _split_and_check(random_table([O.INT8] * 20, 300, seed=4), [0, 13, 300])     # 3 bitset bytes


def test_assemble_round_trip():
    G = _gpu()
    import srj_b200 as S
    from srj_b200.kudo import KudoGpuSerializer as KS
    cols = random_table(TYPES, 12_345, seed=6)
    splits = [0, 0, 5, 5, 1000, 1001, 7777, 12_345, 12_345]
    buf, offs = _split_and_check(cols, splits)
    tbl = KS.assembleFromDeviceRaw([S.DType(t) for t in TYPES], buf, offs)
    want = K.assemble(*K.split(cols, splits), TYPES)
    for i, (g, w, c) in enumerate(zip(tbl.columns, want, cols)):
        h = G.to_host(g)
        assert cols_equal(h, w), f"column {i} vs oracle"
        assert cols_equal(h, c), f"column {i} vs the input table"
        if TYPES[i] == O.STRING:
            assert np.array_equal(h.offsets, c.offsets)


def test_assemble_partitions_of_two_tables():
    """The reader side of a shuffle: partitions written by different map tasks concatenated into one table."""
    G = _gpu()
    import srj_b200 as S
    from srj_b200.kudo import KudoGpuSerializer as KS
    a = random_table(TYPES, 3000, seed=7)
    b = random_table(TYPES, 2000, seed=8, null_frac=0.0)                      # no masks: partitions without validity
    ba, oa = K.split(a, [0, 1234, 3000])
    bb, ob = K.split(b, [0, 77, 2000])
    # take partition 1 of a, partition 0 of b, partition 0 of a
    pieces = [ba[oa[1]:oa[2]], bb[ob[0]:ob[1]], ba[oa[0]:oa[1]]]
    buf = np.concatenate(pieces)
    offs = np.zeros(4, np.int64)
    np.cumsum([len(p) for p in pieces], out=offs[1:])
    tbl = KS.assembleFromDeviceRaw([S.DType(t) for t in TYPES], torch.from_numpy(buf).cuda(), torch.from_numpy(offs).cuda())
    want = K.assemble(buf, offs, TYPES)
    for i, (g, w) in enumerate(zip(tbl.columns, want)):
        assert cols_equal(G.to_host(g), w), f"column {i}"
    assert tbl.getRowCount() == (3000 - 1234) + 77 + 1234


def test_hash_partition_then_split():
    """GpuHashPartitioning -> shuffle_split: the partition offsets of srj_hash_partition are the splits."""
    G = _gpu()
    import srj_b200 as S
    from srj_b200.kudo import KudoGpuSerializer as KS
    from srj_b200.partitioning import HashPartitioner
    cols = random_table(TYPES, 30_000, seed=9)
    P = 37
    pt = HashPartitioner.partition(G.table_to_device(cols), [0, 2], P)
    splits = pt.getPartitions() + [30_000]
    buf, offs = KS.splitAndSerializeToDevice(pt.getTable(), splits)
    ids = O.partition_ids([cols[0], cols[2]], P)
    want_cols, want_offs, _ = O.stable_partition(cols, ids, P)
    want_buf, want_boffs = K.split(want_cols, want_offs)
    assert np.array_equal(offs.cpu().numpy(), want_boffs) and np.array_equal(buf.cpu().numpy(), want_buf)
    back = KS.assembleFromDeviceRaw([S.DType(t) for t in TYPES], buf, offs)
    for g, w in zip(back.columns, want_cols):
        assert cols_equal(G.to_host(g), w)


def test_malformed_header_is_rejected():
    G = _gpu()
    import srj_b200 as S
    from srj_b200.kudo import KudoGpuSerializer as KS
    cols = random_table([O.INT32], 100, seed=1)
    buf, offs = K.split(cols, [0, 100])
    buf = buf.copy()
    buf[0] = ord("X")
    with pytest.raises(S.CudfException):
        KS.assembleFromDeviceRaw([S.DType(O.INT32)], torch.from_numpy(buf).cuda(), torch.from_numpy(offs).cuda())
    with pytest.raises(S.CudfException):
        KS.splitAndSerializeToDevice(S.Table([S.ColumnView.makeStructView(G.to_device(cols[0]))]), [0, 100])
    with pytest.raises(S.CudfColumnSizeOverflowException):                      # splits must be increasing
        KS.splitAndSerializeToDevice(G.table_to_device(cols), [0, 60, 40, 100])
