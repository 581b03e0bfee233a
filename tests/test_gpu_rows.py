"""GPU parity tests (call through the C ABI via the srj_b200 host mirror) against the CPU oracle.
Bit-exact: every data byte, mask word, offset, null count and (to_rows) every row byte incl. the
zero padding.  Mirrors the reference's tests/row_conversion.cpp cases (Single, Tall, Wide,
SingleByteWide, Non2Power, Big, AllTypes, PivotLikeLayout, SimpleString, DoubleString, ManyStrings)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from util import cols_equal, random_table

pytestmark = pytest.mark.gpu

ALL_FIXED = [O.INT8, O.INT16, O.INT32, O.INT64, O.UINT8, O.UINT16, O.UINT32, O.UINT64, O.FLOAT32, O.FLOAT64,
             O.BOOL8, O.TIMESTAMP_DAYS, O.TIMESTAMP_SECONDS, O.TIMESTAMP_MILLISECONDS, O.TIMESTAMP_MICROSECONDS,
             O.TIMESTAMP_NANOSECONDS, O.DURATION_DAYS, O.DURATION_SECONDS, O.DURATION_MILLISECONDS,
             O.DURATION_MICROSECONDS, O.DURATION_NANOSECONDS, O.DECIMAL32, O.DECIMAL64, O.DECIMAL128]

FIXED_SCHEMAS = {
    "single": [O.INT32],                                                          # ColumnToRowTests.Single
    "c1": [O.INT32, O.INT64, O.FLOAT64, O.BOOL8],
    "c2": [O.INT8, O.INT16, O.INT32, O.INT64, O.FLOAT32, O.FLOAT64, O.BOOL8, O.TIMESTAMP_MICROSECONDS] * 4,
    "c4_store_sales": [O.INT32] * 9 + [O.INT64, O.INT32] + [O.DECIMAL32] * 12,
    "all_types": ALL_FIXED,                                                        # AllTypes
    "wide_int32": [O.INT32] * 256,                                                 # Wide
    "single_byte_wide": [O.INT8] * 256,                                            # SingleByteWide
    "non2power": [O.INT64 if i % 2 else O.INT16 for i in range(131)],              # Non2Power (131 cols)
    "pivot": [O.INT64] * 191 + [O.INT32],                                          # PivotLikeLayout
    "row_gt_2k": [O.DECIMAL128] * 200,                                             # 3225 B rows: 2-stage ring
    "row_8k": [O.DECIMAL128] * 520,                                                # 8 rows/tile
}


def _gpu():
    import gpu_util
    gpu_util.require_cuda()
    return gpu_util


@pytest.mark.parametrize("nrows", [1, 31, 32, 33, 257, 6701, 100_003])
@pytest.mark.parametrize("name", sorted(FIXED_SCHEMAS))
def test_fixed_width_both_directions(name, nrows):
    G = _gpu()
    import srj_b200 as S
    types = FIXED_SCHEMAS[name]
    if len(types) * nrows > 30_000_000:
        nrows = 30_000_000 // len(types)
    cols = random_table(types, nrows, seed=nrows * 7 + len(types))
    (offs, data), = O.convert_to_rows(cols)                      # oracle rows

    # ---- to_rows: every output byte (incl. zero padding) and the LIST offsets
    out = S.RowConversion.convertToRows(G.table_to_device(cols))
    assert len(out) == 1
    goffs, gdata = G.rows_to_host(out[0])
    assert np.array_equal(goffs, offs)
    assert np.array_equal(gdata, data), f"first diff at byte {np.flatnonzero(gdata != data)[:5]}"

    # ---- from_rows: data bytes (incl. null payload), masks, null counts
    tbl = S.RowConversion.convertFromRows(G.rows_to_device(offs, data), [S.DType(t) for t in types])
    ocols, onulls = O.convert_from_rows(data, None, nrows, types)
    for i, (g, o) in enumerate(zip(tbl.columns, ocols)):
        h = G.to_host(g)
        assert cols_equal(h, o, check_null_payload=True), f"column {i}"
        assert np.array_equal(h.mask, o.mask), f"mask words differ, column {i}"   # incl. zero tail bits
        assert g.getNullCount() == int(onulls[i]) == cols[i].null_count()
        assert cols_equal(h, cols[i], check_null_payload=True)


def test_legacy_entry_points_match_general():
    """tests/row_conversion.cpp:37-55 etc.: *_fixed_width_optimized == general path, byte for byte."""
    G = _gpu()
    import srj_b200 as S
    types = [O.INT8, O.INT32, O.INT16, O.INT64, O.INT32, O.BOOL8, O.UINT16, O.UINT8, O.UINT64]
    cols = random_table(types, 4097, seed=3)
    t = G.table_to_device(cols)
    a = S.RowConversion.convertToRows(t)
    b = S.RowConversion.convertToRowsFixedWidthOptimized(t)
    assert torch.equal(a[0].child.data, b[0].child.data) and torch.equal(a[0].offsets, b[0].offsets)
    dts = [S.DType(x) for x in types]
    ta = S.RowConversion.convertFromRows(a[0], dts)
    tb = S.RowConversion.convertFromRowsFixedWidthOptimized(a[0], *dts)
    for x, y, c in zip(ta.columns, tb.columns, cols):
        assert torch.equal(x.data, y.data) and torch.equal(x.mask, y.mask)
        assert cols_equal(G.to_host(x), c, check_null_payload=True)
    with pytest.raises(S.CudfException):
        S.RowConversion.convertToRowsFixedWidthOptimized(G.table_to_device(random_table([O.STRING], 4)))
    with pytest.raises(S.CudfException):   # RC:1184-1191: row_size * 32 must fit 48 KB
        S.RowConversion.convertToRowsFixedWidthOptimized(G.table_to_device(random_table([O.INT64] * 300, 4)))


STRING_SCHEMAS = {
    "simple_string": [O.STRING],                                                   # SimpleString
    "double_string": [O.INT32, O.STRING, O.STRING],                                # DoubleString
    "mixed": [O.INT32, O.STRING, O.INT64, O.DECIMAL128, O.STRING, O.BOOL8, O.STRING, O.INT16],
    "c3_small": [O.INT32, O.INT64, O.DECIMAL128, O.STRING] * 8,
    "c3": [O.INT32, O.INT64, O.DECIMAL128, O.STRING] * 64,
    "many_strings": [O.STRING] * 50,                                               # ManyStrings (50 cols)
}


@pytest.mark.parametrize("nrows", [1, 8, 33, 1000, 20_011])
@pytest.mark.parametrize("name", sorted(STRING_SCHEMAS))
def test_strings_both_directions(name, nrows):
    G = _gpu()
    import srj_b200 as S
    types = STRING_SCHEMAS[name]
    if len(types) * nrows > 1_500_000:
        nrows = 1_500_000 // len(types)
    cols = random_table(types, nrows, seed=nrows + 13)
    (offs, data), = O.convert_to_rows(cols)

    out = S.RowConversion.convertToRows(G.table_to_device(cols))
    assert len(out) == 1
    goffs, gdata = G.rows_to_host(out[0])
    assert np.array_equal(goffs, offs)
    assert np.array_equal(gdata, data), f"first diff at byte {np.flatnonzero(gdata != data)[:5]}"

    tbl = S.RowConversion.convertFromRows(G.rows_to_device(offs, data), [S.DType(t) for t in types])
    ocols, onulls = O.convert_from_rows(data, offs, nrows, types)
    for i, (g, o) in enumerate(zip(tbl.columns, ocols)):
        h = G.to_host(g)
        assert np.array_equal(h.mask, o.mask), f"mask, column {i}"
        if types[i] == O.STRING:
            assert np.array_equal(h.offsets, o.offsets), f"offsets, column {i}"
            assert np.array_equal(h.data, o.data), f"chars, column {i}"
        else:
            assert cols_equal(h, o, check_null_payload=True), f"column {i}"
        assert g.getNullCount() == int(onulls[i])
        assert cols_equal(h, cols[i])


def test_big_strings_and_empty_strings():
    """BigStrings (tests/row_conversion.cpp:943) + rows larger than a pipeline stage (SAFE path)."""
    G = _gpu()
    import srj_b200 as S
    rng = np.random.default_rng(1)
    big = [bytes(rng.integers(32, 127, n, dtype=np.uint8)) for n in (0, 1, 300_000, 5, 0, 70_000, 1_100_000, 3)]
    vals = big + [b"", None, b"x"] * 10
    c0 = O.strings_col(vals)
    c1 = O.HCol(O.INT64, rng.integers(0, 2**62, len(vals)).astype(np.int64).view(np.uint8), None, None, 0, len(vals))
    c2 = O.strings_col([b"tail%d" % i for i in range(len(vals))])
    cols = [c0, c1, c2]
    types = [c.type_id for c in cols]
    (offs, data), = O.convert_to_rows(cols)
    out = S.RowConversion.convertToRows(G.table_to_device(cols))
    goffs, gdata = G.rows_to_host(out[0])
    assert np.array_equal(goffs, offs) and np.array_equal(gdata, data)
    tbl = S.RowConversion.convertFromRows(out[0], [S.DType(t) for t in types])
    for g, c in zip(tbl.columns, cols):
        assert cols_equal(G.to_host(g), c)


def test_empty_table_and_errors():
    G = _gpu()
    import srj_b200 as S
    types = [O.INT32, O.STRING]
    cols = random_table(types, 0)
    out = S.RowConversion.convertToRows(G.table_to_device(cols))
    assert len(out) == 1 and out[0].size == 0 and out[0].child.size == 0
    tbl = S.RowConversion.convertFromRows(out[0], [S.DType(t) for t in types])
    assert tbl.getRowCount() == 0 and tbl.getNumberOfColumns() == 2
    # unsupported schema (RowConversion.java:131)
    with pytest.raises(S.CudfException):
        S.RowConversion.convertFromRows(out[0], [S.DType(S.DType.LIST)])
    # "The layout of the data appears to be off" (RC:2197)
    small = G.rows_to_device(np.arange(11, dtype=np.int32) * 8, np.zeros(80, np.uint8))
    with pytest.raises(S.CudfException):
        S.RowConversion.convertFromRows(small, [S.DType(S.DType.INT64), S.DType(S.DType.INT64)])
    # only LIST<INT8/UINT8> input (RC:2157)
    with pytest.raises(S.CudfException):
        S.RowConversion.convertFromRows(S.ColumnVector(S.DType.INT32, 0), [S.DType(S.DType.INT32)])


def test_null_patterns_all_none_half_sparse():
    """AllTypesLarge null patterns (tests/row_conversion.cpp:654-777): all / none / 1-in-2 / 1-in-13."""
    G = _gpu()
    import srj_b200 as S
    n = 20_000
    types = [O.INT8, O.INT64, O.DECIMAL128, O.FLOAT32]
    cols = random_table(types, n, seed=9, null_frac=0.0)
    pats = [np.zeros(n, bool), np.ones(n, bool), np.arange(n) % 2 == 0, np.arange(n) % 13 != 0]
    for c, v in zip(cols, pats):
        c.mask = O.pack_mask(v)
    out = S.RowConversion.convertToRows(G.table_to_device(cols))
    (offs, data), = O.convert_to_rows(cols)
    assert np.array_equal(G.rows_to_host(out[0])[1], data)
    tbl = S.RowConversion.convertFromRows(out[0], [S.DType(t) for t in types])
    for g, c, v in zip(tbl.columns, cols, pats):
        assert cols_equal(G.to_host(g), c, check_null_payload=True)
        assert g.getNullCount() == int((~v).sum())
    # an input without a null mask == all valid (RC:757-759)
    cols[1].mask = None
    out2 = S.RowConversion.convertToRows(G.table_to_device(cols))
    assert torch.equal(out2[0].child.data, out[0].child.data)


@pytest.mark.parametrize("kind", ["xxhash64", "murmur3", "hive"])
def test_fused_from_rows_hash(kind):
    """BASELINE config 4: store_sales schema, from_rows fused with the partition hash of
    (ss_item_sk, ss_ticket_number) -- must equal the standalone hash of the converted columns."""
    G = _gpu()
    import srj_b200 as S
    types = [O.INT32] * 9 + [O.INT64, O.INT32] + [O.DECIMAL32] * 12
    n = 50_001
    cols = random_table(types, n, seed=4, null_frac=0.04)
    (offs, data), = O.convert_to_rows(cols)
    keys = [1, 9] if kind != "hive" else [1, 9, 10]
    tbl, h = S.RowConversion.convertFromRowsWithHash(G.rows_to_device(offs, data), [S.DType(t) for t in types], keys,
                                                     kind=kind, seed=42)
    kc = [cols[k] for k in keys]
    want = {"xxhash64": lambda: O.xxhash64(kc, 42), "murmur3": lambda: O.murmur_hash3_32(kc, 42),
            "hive": lambda: O.hive_hash(kc)}[kind]()
    got = h.data.cpu().numpy().view(want.dtype)
    assert np.array_equal(got, want)
    for g, c in zip(tbl.columns, cols):
        assert cols_equal(G.to_host(g), c, check_null_payload=True)


def test_batch_split_over_2gib():
    """build_batches (RC:1466-1557): > 2 GiB of rows -> several LIST columns cut on 32-row boundaries;
    checked by round trip + per-batch oracle bytes on a sample."""
    G = _gpu()
    import srj_b200 as S
    types = [O.INT64] * 100                          # 800 B data + 13 B validity -> 816 B rows
    n = 3_000_000                                    # 2.45 GB -> 2 batches
    g = torch.Generator(device="cuda").manual_seed(5)
    dcols = [S.ColumnVector(S.DType(t), n, torch.randint(0, 256, (n * 8,), dtype=torch.uint8, device="cuda", generator=g),
                            torch.randint(-2**31, 2**31 - 1, ((n + 31) // 32,), dtype=torch.int32, device="cuda", generator=g))
             for t in types]
    out = S.RowConversion.convertToRows(S.Table(dcols))
    assert len(out) == 2
    cut = out[0].size
    assert cut % 32 == 0 and out[0].child.size <= 2**31 - 1 and cut + out[1].size == n
    assert cut == (2**31 - 1 + 815) // 816 // 32 * 32            # lower_bound then round down to 32 rows
    dts = [S.DType(t) for t in types]
    r0 = 0
    for b in out:
        tbl = S.RowConversion.convertFromRows(b, dts)
        for c, src in zip(tbl.columns, dcols):
            assert torch.equal(c.data, src.data[r0 * 8:(r0 + b.size) * 8])
        # validity round trip (compare bits of rows r0.. in the source masks)
        m0 = tbl.columns[7].mask.cpu().numpy().view(np.uint32)
        src_bits = np.unpackbits(dcols[7].mask.cpu().numpy().view(np.uint8), bitorder="little")[r0:r0 + b.size]
        assert np.array_equal(np.unpackbits(m0.view(np.uint8), bitorder="little")[:b.size], src_bits)
        r0 += b.size
    # oracle bytes for the first 1000 rows of batch 1
    sample = [O.HCol(t, c.data[cut * 8:(cut + 1000) * 8].cpu().numpy(),
                     O.pack_mask(np.unpackbits(c.mask.cpu().numpy().view(np.uint8), bitorder="little")[cut:cut + 1000].astype(bool)),
                     None, 0, 1000) for t, c in zip(types, dcols)]
    (_, odata), = O.convert_to_rows(sample)
    assert np.array_equal(out[1].child.data[: len(odata)].cpu().numpy(), odata)


def test_non_canonical_string_layout_follows_pair_offsets():
    """copy_strings_from_rows reads chars at row + pair.offset (RC:1143), whatever the order inside the
    row.  Rows whose variable section is permuted (pairs updated) must still convert correctly: phase 1
    flags them and phase 2 falls back from its canonical fast path to the generic gather."""
    G = _gpu()
    import srj_b200 as S
    types = [O.INT32, O.STRING, O.INT64, O.STRING]
    n = 3000
    cols = random_table(types, n, seed=21)
    (offs, data), = O.convert_to_rows(cols)
    st, sz, voff, spr = O.compute_layout(types)
    data = data.copy()
    for r in range(n):
        row = data[offs[r]:offs[r + 1]]
        (o1, l1), (o3, l3) = row[st[1]:st[1] + 8].view(np.uint32), row[st[3]:st[3] + 8].view(np.uint32)
        a, b = row[o1:o1 + l1].copy(), row[o3:o3 + l3].copy()
        row[spr:spr + l3] = b                      # column 3's chars first ...
        row[spr + l3:spr + l3 + l1] = a            # ... then column 1's
        row[st[3]:st[3] + 8].view(np.uint32)[:] = (spr, l3)
        row[st[1]:st[1] + 8].view(np.uint32)[:] = (spr + l3, l1)
    want, _ = O.convert_from_rows(data, offs, n, types)
    tbl = S.RowConversion.convertFromRows(G.rows_to_device(offs, data), [S.DType(t) for t in types])
    for g, w, c in zip(tbl.columns, want, cols):
        assert cols_equal(G.to_host(g), w) and cols_equal(G.to_host(g), c)
