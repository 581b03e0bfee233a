"""GPU parity tests closing the round-1 holes: unaligned buffers for every kernel family, a > 2 GiB STRING table
(multi-batch split), the fused hash over every supported key type and over schemas with STRING columns, very wide
schemas (shared-memory table guard), many threads sharing one plan."""
import threading

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


def _shifted(t: torch.Tensor, shift: int) -> torch.Tensor:
    big = torch.zeros(t.numel() * t.element_size() + 64, dtype=torch.uint8, device=t.device)
    v = big[shift: shift + t.numel() * t.element_size()]
    v.copy_(t.contiguous().view(torch.uint8))
    return v


# schemas that exercise each kernel family: fixed-width fast kernels, the whole-row var kernels
# (narrow rows with strings: to_rows_w / multi-group from_rows), the wide path, the generic mid-size string rows
FAMILIES = {
    "fixed_c2": [O.INT8, O.INT16, O.INT32, O.INT64, O.FLOAT32, O.FLOAT64, O.BOOL8, O.TIMESTAMP_MICROSECONDS] * 4,
    "fixed_dec": [O.DECIMAL128, O.INT8, O.DECIMAL128, O.INT16],
    "narrow_str": [O.INT32, O.STRING, O.INT64, O.DECIMAL128, O.STRING, O.BOOL8, O.STRING, O.INT16],
    "mid_str": [O.INT64, O.STRING, O.DECIMAL128] * 24,          # ~1 KB rows: generic to_rows kernel
    "wide_c3": [O.INT32, O.INT64, O.DECIMAL128, O.STRING] * 64,
}


@pytest.mark.parametrize("shift", [1, 4, 8])
@pytest.mark.parametrize("name", sorted(FAMILIES))
def test_from_rows_unaligned_row_buffer(name, shift):
    """rows child at +1 / +4 (SAFE paths: rows not 8-byte aligned) and +8 (aligned rows, 16-byte windows clipped at
    the buffer ends)."""
    G = _gpu()
    import srj_b200 as S
    types = FAMILIES[name]
    nrows = 3000 if len(types) < 100 else 700
    cols = random_table(types, nrows, seed=shift * 11 + len(types))
    (offs, data), = O.convert_to_rows(cols)
    d = _shifted(torch.from_numpy(data).cuda(), shift)
    assert d.data_ptr() % 16 == shift % 16
    vec = S.ColumnVector(S.DType.LIST, nrows, None, None, torch.from_numpy(offs).cuda(), S.ColumnVector(S.DType.INT8, len(data), d))
    tbl = S.RowConversion.convertFromRows(vec, [S.DType(t) for t in types])
    has_str = O.STRING in types
    want, nulls = O.convert_from_rows(data, offs if has_str else None, nrows, types)
    for i, (g, w) in enumerate(zip(tbl.columns, want)):
        h = G.to_host(g)
        assert np.array_equal(h.mask, w.mask), f"mask, column {i}"
        if types[i] == O.STRING:
            assert np.array_equal(h.offsets, w.offsets) and np.array_equal(h.data, w.data), f"string column {i}"
        else:
            assert cols_equal(h, w, check_null_payload=True), f"column {i}"
        assert g.getNullCount() == int(nulls[i])


@pytest.mark.parametrize("shift", [8, 16, 24])
@pytest.mark.parametrize("name", sorted(FAMILIES))
def test_to_rows_sliced_inputs_and_unaligned_output(name, shift):
    """to_rows with column buffers that are element-offset slices of larger allocations (fixed-width data at
    non-16-byte addresses, chars at odd addresses) and, through the C ABI, an output buffer at +8."""
    G = _gpu()
    import ctypes as C
    import srj_b200 as S
    from srj_b200 import _native as N
    types = FAMILIES[name]
    nrows = 2000 if len(types) < 100 else 500
    cols = random_table(types, nrows, seed=shift + len(types))
    (offs, data), = O.convert_to_rows(cols)
    dcols = []
    for c in cols:
        dc = G.to_device(c)
        sz = max(1, S.DType(c.type_id).size_in_bytes())
        sh = sz if c.type_id != O.STRING else 3             # one element (odd byte count for chars)
        if dc.data is not None and dc.data.numel():
            dc.data = _shifted(dc.data, sh if (sh % sz == 0 or c.type_id == O.STRING) else sz)
        if dc.mask is not None:
            dc.mask = _shifted(dc.mask, 4).view(torch.int32)
        if dc.offsets is not None:
            dc.offsets = _shifted(dc.offsets, 4).view(torch.int32)
        dcols.append(dc)
    out = S.RowConversion.convertToRows(S.Table(dcols))
    goffs, gdata = G.rows_to_host(out[0])
    assert np.array_equal(goffs, offs)
    assert np.array_equal(gdata, data), f"first diff at byte {np.flatnonzero(gdata != data)[:5]}"
    # the same conversion into an output buffer that is only 8-byte aligned (C ABI: caller-owned batch buffers)
    plan = S.Plan.get([c.dtype for c in dcols])
    lib = N.lib()
    carr = (N.SrjColumn * len(dcols))()
    for i, c in enumerate(dcols):
        carr[i] = c._c()
    st = int(torch.cuda.current_stream().cuda_stream)
    ws = torch.empty(max(8, lib.srj_to_rows_workspace_bytes(plan.handle, nrows)), dtype=torch.uint8, device="cuda")
    rb = (N.SrjRowBatch * 4)()
    nb = C.c_int32(0)
    N.check(lib.srj_to_rows_plan_batches(plan.handle, carr, nrows, ws.data_ptr(), rb, 4, C.byref(nb), st))
    assert nb.value == 1 and rb[0].num_bytes == len(data)
    big = torch.zeros(len(data) + 64, dtype=torch.uint8, device="cuda")
    o_off = torch.empty(nrows + 1, dtype=torch.int32, device="cuda")
    op, dp = (C.c_void_p * 1)(o_off.data_ptr()), (C.c_void_p * 1)(big.data_ptr() + shift)
    N.check(lib.srj_convert_to_rows(plan.handle, carr, nrows, ws.data_ptr(), rb, 1, op, dp, st))
    torch.cuda.synchronize()
    got = big[shift: shift + len(data)].cpu().numpy()
    assert np.array_equal(got, data), f"unaligned output: first diff at byte {np.flatnonzero(got != data)[:5]}"
    assert not big[:shift].any() and not big[shift + len(data):].any(), "wrote outside the batch buffer"


def test_string_table_over_2gib_is_split_into_batches():
    """build_batches (RC:1466-1557) on a variable-width table: > 2 GiB of rows -> two LIST columns cut on a 32-row
    boundary; oracle bytes on a sample of each batch, and every batch converts back to its source rows."""
    G = _gpu()
    import srj_b200 as S
    types = [O.INT64, O.STRING] * 20
    n = 4_200_000                                   # ~ 590 B rows -> ~2.5 GB
    g = torch.Generator(device="cuda").manual_seed(9)
    words = (n + 31) // 32
    dcols = []
    for t in types:
        mask = torch.randint(-2**31, 2**31 - 1, (words,), dtype=torch.int32, device="cuda", generator=g)
        if t == O.STRING:
            lens = torch.randint(0, 27, (n,), dtype=torch.int64, device="cuda", generator=g)
            offs = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
            offs[1:] = torch.cumsum(lens, 0)
            chars = torch.randint(32, 127, (int(offs[-1]),), dtype=torch.uint8, device="cuda", generator=g)
            dcols.append(S.ColumnVector(S.DType(t), n, chars, mask, offs.to(torch.int32)))
        else:
            dcols.append(S.ColumnVector(S.DType(t), n, torch.randint(0, 256, (n * 8,), dtype=torch.uint8, device="cuda", generator=g), mask))
    out = S.RowConversion.convertToRows(S.Table(dcols))
    assert len(out) == 2, f"{len(out)} batches"
    cut = out[0].size
    assert cut % 32 == 0 and cut + out[1].size == n
    assert all(b.child.size <= 2**31 - 1 for b in out)
    # the cut is the LAST 32-row boundary that keeps batch 0 within INT32_MAX bytes
    rs = [int(b.offsets[-1]) for b in out]
    assert rs[0] == out[0].child.size and rs[1] == out[1].child.size
    extra = int(out[1].offsets[32])                 # bytes of the next 32 rows
    assert out[0].child.size + extra > 2**31 - 1

    def host_sample(r0, cnt):
        cols = []
        for t, c in zip(types, dcols):
            valid = np.unpackbits(c.mask[r0 // 32:(r0 + cnt + 31) // 32].cpu().numpy().view(np.uint8), bitorder="little")[: cnt].astype(bool)
            if t == O.STRING:
                o = c.offsets[r0: r0 + cnt + 1].cpu().numpy().astype(np.int64)
                cols.append(O.HCol(t, c.data[int(o[0]): int(o[-1])].cpu().numpy(), O.pack_mask(valid), (o - o[0]).astype(np.int32), 0, cnt))
            else:
                cols.append(O.HCol(t, c.data[r0 * 8:(r0 + cnt) * 8].cpu().numpy(), O.pack_mask(valid), None, 0, cnt))
        return cols
    for bi, r0 in ((0, 0), (0, cut - 512), (1, cut), (1, (n - 500) // 32 * 32)):
        cnt = 512 if r0 + 512 <= n else n - r0
        cnt = min(cnt, (cut - r0) if bi == 0 else (n - r0))
        (ooffs, odata), = O.convert_to_rows(host_sample(r0, cnt))
        b = out[bi]
        lr0 = r0 - (0 if bi == 0 else cut)
        lo, hi = int(b.offsets[lr0]), int(b.offsets[lr0 + cnt])
        assert hi - lo == len(odata)
        assert np.array_equal(b.child.data[lo:hi].cpu().numpy(), odata), f"batch {bi} rows {r0}.."
        assert np.array_equal((b.offsets[lr0: lr0 + cnt + 1] - b.offsets[lr0]).cpu().numpy(), ooffs)
    # round trip of each batch
    dts = [S.DType(t) for t in types]
    r0 = 0
    for b in out:
        tbl = S.RowConversion.convertFromRows(b, dts)
        for t, c, src in zip(types, tbl.columns, dcols):
            if t == O.STRING:
                o0, o1 = int(src.offsets[r0]), int(src.offsets[r0 + b.size])
                assert torch.equal(c.offsets, src.offsets[r0: r0 + b.size + 1] - src.offsets[r0])
                assert torch.equal(c.data, src.data[o0:o1])
            else:
                assert torch.equal(c.data, src.data[r0 * 8:(r0 + b.size) * 8])
        r0 += b.size


KEY_TYPES = [O.BOOL8, O.INT8, O.INT16, O.INT32, O.INT64, O.FLOAT32, O.FLOAT64, O.TIMESTAMP_DAYS, O.TIMESTAMP_MICROSECONDS,
             O.DECIMAL32, O.DECIMAL64, O.DECIMAL128, O.UINT8, O.UINT16, O.UINT32, O.UINT64]
HIVE_KEY_TYPES = [O.BOOL8, O.INT8, O.INT16, O.INT32, O.INT64, O.FLOAT32, O.FLOAT64, O.TIMESTAMP_DAYS, O.TIMESTAMP_MICROSECONDS]


@pytest.mark.parametrize("with_strings", [False, True])
@pytest.mark.parametrize("kind", ["xxhash64", "murmur3", "hive"])
def test_fused_hash_every_key_type(kind, with_strings):
    """from_rows fused with the row hash, one run per key type plus all keys at once, on a fixed-width schema and on a
    schema with STRING non-key columns (wide enough for the wide plan: the fused call falls back to the whole-row
    kernel and phase 2 must notice that the offsets are finished)."""
    G = _gpu()
    import srj_b200 as S
    keyt = HIVE_KEY_TYPES if kind == "hive" else KEY_TYPES
    types = list(keyt)
    if with_strings:
        types = types + [O.STRING, O.INT64] * 30          # 30 STRING columns, ~600-byte fixed section: a wide plan
    n = 5003
    cols = random_table(types, n, seed=17, null_frac=0.15)
    # float keys: sprinkle NaNs and signed zeros (normalisation rules differ per hash, Appendix B)
    for ci, t in enumerate(types):
        if t in (O.FLOAT32, O.FLOAT64):
            v = cols[ci].data.view(np.float32 if t == O.FLOAT32 else np.float64)
            v[::7] = np.nan
            v[1::7] = -0.0
            v[2::7] = 0.0
    (offs, data), = O.convert_to_rows(cols)
    vec = G.rows_to_device(offs, data)
    dts = [S.DType(t, c.scale) for t, c in zip(types, cols)]
    runs = [[k] for k in range(len(keyt))] + [list(range(len(keyt)))]
    for keys in runs[:: (1 if not with_strings else 3)] + [runs[-1]]:
        tbl, h = S.RowConversion.convertFromRowsWithHash(vec, dts, keys, kind=kind, seed=42)
        kc = [cols[k] for k in keys]
        want = {"xxhash64": lambda: O.xxhash64(kc, 42), "murmur3": lambda: O.murmur_hash3_32(kc, 42), "hive": lambda: O.hive_hash(kc)}[kind]()
        got = h.data.cpu().numpy().view(want.dtype)
        assert np.array_equal(got, want), f"{kind} keys {keys}: {np.flatnonzero(got != want)[:5]}"
        for i, (g, c) in enumerate(zip(tbl.columns, cols)):
            assert cols_equal(G.to_host(g), c), f"column {i}"
    if with_strings:
        # unfused conversion of the same rows (wide path) agrees
        t2 = S.RowConversion.convertFromRows(vec, dts)
        for g, c in zip(t2.columns, cols):
            assert cols_equal(G.to_host(g), c)


@pytest.mark.parametrize("types", [[O.INT32] * 1500, [O.INT8, O.INT64, O.DECIMAL128, O.INT16] * 500,
                                   [O.INT32, O.STRING] * 600], ids=["1500xINT32", "2000xmixed", "1200 with 600 strings"])
def test_very_wide_schemas(types):
    """Schemas whose per-column shared-memory tables are tens of KB: the tilings must shrink instead of failing."""
    G = _gpu()
    import srj_b200 as S
    n = 257
    cols = random_table(types, n, seed=3)
    (offs, data), = O.convert_to_rows(cols)
    out = S.RowConversion.convertToRows(G.table_to_device(cols))
    goffs, gdata = G.rows_to_host(out[0])
    assert np.array_equal(goffs, offs) and np.array_equal(gdata, data)
    tbl = S.RowConversion.convertFromRows(out[0], [S.DType(t, c.scale) for t, c in zip(types, cols)])
    for i, (g, c) in enumerate(zip(tbl.columns, cols)):
        assert cols_equal(G.to_host(g), c, check_null_payload=(c.type_id != O.STRING)), f"column {i}"


def test_sixteen_threads_share_one_plan():
    """Many Spark task threads, one schema: 16 threads convert different tables through the same plan on their own
    streams at once (the plan's pointer-table ring has 8 slots)."""
    G = _gpu()
    import srj_b200 as S
    types = [O.INT32, O.STRING, O.INT64, O.DECIMAL128, O.STRING, O.BOOL8]
    fixed = [O.INT8, O.INT16, O.INT32, O.INT64, O.FLOAT64]
    wide = [O.INT32, O.INT64, O.DECIMAL128, O.STRING] * 64
    errs = []

    def work(tid):
        try:
            torch.cuda.set_device(0)
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                for it in range(6):
                    sch = (types, fixed, wide)[(tid + it) % 3]
                    n = 500 + 97 * tid + 13 * it
                    cols = random_table(sch, n, seed=1000 * tid + it)
                    (offs, data), = O.convert_to_rows(cols)
                    out = S.RowConversion.convertToRows(G.table_to_device(cols))
                    goffs, gdata = G.rows_to_host(out[0])
                    assert np.array_equal(goffs, offs) and np.array_equal(gdata, data), f"thread {tid} it {it} to_rows"
                    tbl = S.RowConversion.convertFromRows(out[0], [S.DType(t) for t in sch])
                    for g, c in zip(tbl.columns, cols):
                        assert cols_equal(G.to_host(g), c), f"thread {tid} it {it} from_rows"
        except Exception as ex:            # noqa: BLE001
            errs.append((tid, repr(ex)))
    th = [threading.Thread(target=work, args=(t,)) for t in range(16)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs[:3]
