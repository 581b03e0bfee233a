"""Pins the CPU oracle's JCUDF row layout against the reference's known-answer material
(SURVEY.md 8c): tests/row_conversion.cpp:457-498 (PivotLikeLayout), the Javadoc example
RowConversion.java:77-105, the config layouts of SURVEY 8(d), and round-trip identity
(the reference's own strategy, tests/row_conversion.cpp:37-455,500-1091)."""
import numpy as np
import pytest

from oracle import oracle as O
from util import cols_equal, random_table, col_from_values


def test_javadoc_example_layout():
    # | A_0 | P | B_0 | B_1 | C_0 | C_1 | C_2 | C_3 | V0 | P*7 |   (RowConversion.java:77-85)
    types = [O.BOOL8, O.INT16, O.DURATION_DAYS]
    starts, sizes, voff, spr = O.compute_layout(types)
    assert list(starts) == [0, 2, 4] and list(sizes) == [1, 2, 4] and voff == 8 and spr == 9
    cols = [col_from_values("BOOL8", [1, 0]), col_from_values("INT16", [0x1234, None]),
            O.HCol(O.DURATION_DAYS, np.array([0x0A0B0C0D, 7], np.int32).view(np.uint8), None, None, 0, 2)]
    (offs, data), = O.convert_to_rows(cols)
    assert list(offs) == [0, 16, 32]
    assert data[:16].tolist() == [1, 0, 0x34, 0x12, 0x0D, 0x0C, 0x0B, 0x0A, 0b111, 0, 0, 0, 0, 0, 0, 0]
    assert data[16 + 8] == 0b101                       # B is null in row 1
    # reordered C, B, A -> | C_0..C_3 | B_0 | B_1 | A_0 | V0 | = 8 bytes  (RowConversion.java:99-103)
    starts, sizes, voff, spr = O.compute_layout([O.DURATION_DAYS, O.INT16, O.BOOL8])
    assert list(starts) == [0, 4, 6] and voff == 7 and spr == 8


def test_pivot_like_layout():
    # tests/row_conversion.cpp:457-498: 191 x INT64 + INT32, 100 rows; int at byte 1528, stride 1560
    nl, nrows = 191, 100
    ints = (0x11223344 + np.arange(nrows)).astype(np.int32)
    cols = [O.HCol(O.INT64, np.zeros(nrows, np.int64).view(np.uint8), None, None, 0, nrows) for _ in range(nl)]
    cols.append(O.HCol(O.INT32, ints.view(np.uint8), None, None, 0, nrows))
    batches = O.convert_to_rows(cols)
    assert len(batches) == 1
    offs, data = batches[0]
    int_offset, validity_bytes = nl * 8, (nl + 1 + 7) // 8
    stride = (int_offset + 4 + validity_bytes + 7) & ~7
    assert stride == 1560 and len(data) >= stride * nrows
    got = np.stack([data[r * stride + int_offset: r * stride + int_offset + 4] for r in range(nrows)]).view(np.int32)
    assert np.array_equal(got.ravel(), ints)
    assert np.array_equal(offs, np.arange(nrows + 1) * stride)


def test_config_layouts():
    # SURVEY 8(a1)/(d)
    assert O.compute_layout([O.INT32, O.INT64, O.FLOAT64, O.BOOL8])[2:] == (25, 26)              # C1: 32 B rows
    c2 = [O.INT8, O.INT16, O.INT32, O.INT64, O.FLOAT32, O.FLOAT64, O.BOOL8, O.TIMESTAMP_MICROSECONDS] * 4
    assert O.compute_layout(c2)[2:] == (192, 196)                                                  # 200 B rows
    c3 = [O.INT32, O.INT64, O.DECIMAL128, O.STRING] * 64
    st, sz, voff, spr = O.compute_layout(c3)
    assert list(st[:8]) == [0, 8, 16, 32, 40, 48, 64, 80] and (voff, spr) == (3064, 3096)
    c4 = [O.INT32] * 9 + [O.INT64, O.INT32] + [O.DECIMAL32] * 12
    st, sz, voff, spr = O.compute_layout(c4)
    assert list(st[:12]) == [0, 4, 8, 12, 16, 20, 24, 28, 32, 40, 48, 52] and (voff, spr) == (100, 103)


def test_unsupported_schema():
    with pytest.raises(NotImplementedError):
        O.compute_layout([O.INT32, O.LIST])


@pytest.mark.parametrize("nrows", [0, 1, 31, 32, 33, 1000, 6701])
def test_roundtrip_fixed(nrows):
    types = [O.INT8, O.INT16, O.INT32, O.INT64, O.FLOAT32, O.FLOAT64, O.BOOL8, O.TIMESTAMP_MICROSECONDS,
             O.DECIMAL32, O.DECIMAL64, O.DECIMAL128, O.UINT8, O.UINT64]
    cols = random_table(types, nrows, seed=nrows + 1)
    batches = O.convert_to_rows(cols)
    assert len(batches) == 1
    offs, data = batches[0]
    back, nulls = O.convert_from_rows(data, None, nrows, types)
    for a, b, nc in zip(cols, back, nulls):
        assert cols_equal(a, b, check_null_payload=True)      # null payload bytes are copied blindly (App. A.6)
        assert nc == a.null_count()
    # fixed-width LIST offsets are r * row_size (RC:2026-2032)
    _, _, _, spr = O.compute_layout(types)
    assert np.array_equal(offs, np.arange(nrows + 1) * ((spr + 7) // 8 * 8))


@pytest.mark.parametrize("nrows", [0, 1, 33, 500])
def test_roundtrip_strings(nrows):
    types = [O.INT32, O.STRING, O.INT64, O.DECIMAL128, O.STRING, O.BOOL8, O.STRING]
    cols = random_table(types, nrows, seed=5)
    (offs, data), = O.convert_to_rows(cols)
    back, nulls = O.convert_from_rows(data, offs, nrows, types)
    for a, b in zip(cols, back):
        assert cols_equal(a, b)
    # row r: pairs hold (running offset from size_per_row, len); chars unpadded in column order
    st, sz, voff, spr = O.compute_layout(types)
    for r in range(min(nrows, 20)):
        row = data[offs[r]:offs[r + 1]]
        run = spr
        for c, t in enumerate(types):
            if t != O.STRING:
                continue
            so, ln = row[st[c]:st[c] + 8].view(np.uint32)
            assert so == run and ln == cols[c].offsets[r + 1] - cols[c].offsets[r]
            run += ln
        assert len(row) == (run + 7) // 8 * 8 and not row[run:].any()   # zero tail padding


def test_c1_plumbing_pyarrow_roundtrip():
    """BASELINE config C1: 64K rows x [INT32, INT64, FLOAT64, BOOL8], 20% nulls, row<->column round
    trip on the host through pyarrow, bit-exact (doubles are random 64-bit patterns incl. NaN payloads)."""
    pa = pytest.importorskip("pyarrow")
    n = 65536
    types = [O.INT32, O.INT64, O.FLOAT64, O.BOOL8]
    cols = random_table(types, n, seed=42, null_frac=0.2)
    (offs, data), = O.convert_to_rows(cols)
    assert len(data) == 32 * n
    back, _ = O.convert_from_rows(data, None, n, types)
    arrs = []
    for c, pat in zip(back, [pa.int32(), pa.int64(), pa.float64(), pa.uint8()]):
        v = c.valid()
        arrs.append(pa.Array.from_buffers(pat, n, [pa.py_buffer(np.packbits(v, bitorder="little").tobytes()),
                                                   pa.py_buffer(np.ascontiguousarray(c.data).tobytes())]))
    rb = pa.RecordBatch.from_arrays(arrs, names=["i", "l", "d", "b"])
    for a, orig in zip(rb.columns, cols):
        bufs = a.buffers()
        assert np.array_equal(np.unpackbits(np.frombuffer(bufs[0], np.uint8), bitorder="little")[:n].astype(bool),
                              orig.valid())
        got = np.frombuffer(bufs[1], np.uint8)[: n * O.size_of(orig.type_id)]
        assert np.array_equal(got, orig.data.view(np.uint8))      # bit patterns, not float compare


def test_build_batches_rule():
    # RC:1500-1517: cut at <= INT_MAX bytes, rounded down to 32 rows
    sizes = np.full(100_000, 65536, dtype=np.uint64)       # 64 KiB rows -> 32767.99 rows per 2 GiB
    b = O.build_batches(sizes)
    assert b[0] == 0 and b[-1] == 100_000
    for lo, hi in zip(b[:-1], b[1:]):
        assert (hi - lo) * 65536 <= 2**31 - 1
        assert hi == 100_000 or (hi - lo) % 32 == 0
    assert b[1] == 32736        # lower_bound gives 32768 (cum[i]-cum[0] >= INT_MAX), round down 32 -> overflow guard
    assert O.build_batches(np.zeros(0, np.uint64)) == [0]
