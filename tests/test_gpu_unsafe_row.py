"""GPU parity tests of the UnsafeRow codec (csrc/unsafe_row.cu) against the CPU restatement of Spark's format
(oracle/unsafe_row.py): row bytes and offsets bit-exact, columns back, null masks and counts, edge cases."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import oracle as O
from oracle import unsafe_row as U
from util import cols_equal, random_table

pytestmark = pytest.mark.gpu


def _gpu():
    import gpu_util
    gpu_util.require_cuda()
    return gpu_util


MIXED = [O.INT32, O.STRING, O.DECIMAL128, O.INT64, O.BOOL8, O.STRING, O.DECIMAL64, O.FLOAT64, O.INT16, O.DECIMAL32, O.INT8, O.FLOAT32]
SCHEMAS = {
    "mixed": MIXED,
    "fixed_only": [O.INT8, O.INT16, O.INT32, O.INT64, O.FLOAT32, O.FLOAT64, O.BOOL8, O.DECIMAL32, O.DECIMAL64],
    "strings_only": [O.STRING] * 5,
    "decimals": [O.DECIMAL128] * 3 + [O.STRING, O.DECIMAL32],
    "wide_70": [O.INT32, O.INT8, O.STRING, O.INT64, O.DECIMAL128, O.INT16, O.BOOL8] * 10,      # two bitset words
    "wide_200": [O.INT32, O.INT64, O.DECIMAL128, O.STRING] * 50,
}


def _check(cols, types):
    G = _gpu()
    import srj_b200 as S
    from srj_b200.unsaferow import UnsafeRowConversion as UR
    offs, data = U.to_unsafe_rows(cols)
    rows = UR.convertToRows(G.table_to_device(cols))
    goffs, gdata = G.rows_to_host(rows)
    assert np.array_equal(goffs.astype(np.int64), offs)
    assert np.array_equal(gdata, data), f"first diff at byte {np.flatnonzero(gdata != data)[:5]} of {len(data)}"
    tbl = UR.convertFromRows(rows, [S.DType(t) for t in types])
    want = U.from_unsafe_rows(data, offs, types)
    for i, (g, w, c) in enumerate(zip(tbl.columns, want, cols)):
        h = G.to_host(g)
        assert cols_equal(h, w), f"column {i} vs oracle"
        assert cols_equal(h, c), f"column {i} vs input (round trip)"
        assert g.getNullCount() == c.null_count(), f"null count, column {i}"


@pytest.mark.parametrize("nrows", [1, 31, 32, 33, 1000, 5001])
@pytest.mark.parametrize("name", sorted(SCHEMAS))
def test_unsafe_rows_match_oracle(name, nrows):
    types = SCHEMAS[name]
    if len(types) * nrows > 300_000:
        nrows = max(33, 300_000 // len(types))
    _check(random_table(types, nrows, seed=nrows + len(types)), types)


@pytest.mark.parametrize("null_frac", [0.0, 1.0])
def test_no_nulls_and_all_nulls(null_frac):
    types = SCHEMAS["mixed"]
    _check(random_table(types, 700, seed=3, null_frac=null_frac), types)


def test_long_and_empty_strings():
    rng = np.random.default_rng(2)
    vals = [b"", None, bytes(rng.integers(32, 127, 5000, dtype=np.uint8)), b"x" * 7, b"y" * 8, b"z" * 9, None, b""] * 20
    cols = [O.strings_col(vals), O.HCol(O.INT64, np.arange(len(vals), dtype=np.int64).view(np.uint8)), O.strings_col(vals[::-1])]
    _check(cols, [O.STRING, O.INT64, O.STRING])


def test_decimal128_extremes():
    vals = [0, 1, -1, 127, 128, -128, -129, 2**63, -2**63, 2**127 - 1, -2**127, 10**37, -10**37, 255, 256, -256, -257]
    raw = b"".join(v.to_bytes(16, "little", signed=True) for v in vals)
    cols = [O.HCol(O.DECIMAL128, np.frombuffer(raw, dtype=np.uint8).copy(), O.pack_mask(np.array([i % 5 != 4 for i in range(len(vals))])))]
    _check(cols, [O.DECIMAL128])


def test_fixed_width_table_without_row_offsets():
    """C ABI: tables without STRING columns convert with d_row_offsets = NULL (rows fixed_bytes apart)."""
    G = _gpu()
    import srj_b200 as S
    from srj_b200 import _native as N
    from srj_b200.unsaferow import UnsafeRowConversion as UR
    types = SCHEMAS["fixed_only"] + [O.DECIMAL128]
    n = 3000
    cols = random_table(types, n, seed=8)
    offs, data = U.to_unsafe_rows(cols)
    bitset, fixed = UR.layout([S.DType(t) for t in types])
    assert bitset == 8 and fixed == 8 + 8 * len(types) + 16 and len(data) == n * fixed
    dcols = [G.to_device(c) for c in cols]
    lib = N.lib()
    st = int(torch.cuda.current_stream().cuda_stream)
    ws = torch.empty(lib.srj_unsafe_row_workspace_bytes(len(types), n), dtype=torch.uint8, device="cuda")
    rows = torch.empty(n * fixed, dtype=torch.uint8, device="cuda")
    carr = (N.SrjColumn * len(dcols))(*[c._c() for c in dcols])
    N.check(lib.srj_convert_to_unsafe_rows(carr, len(dcols), n, None, rows.data_ptr(), ws.data_ptr(), st))
    assert np.array_equal(rows.cpu().numpy(), data)
    outs = [S.ColumnVector(S.DType(t), n, torch.empty(n * O.size_of(t), dtype=torch.uint8, device="cuda"),
                           torch.empty((n + 31) // 32, dtype=torch.int32, device="cuda")) for t in types]
    nulls = torch.zeros(len(types), dtype=torch.int64, device="cuda")
    oarr = (N.SrjColumn * len(outs))(*[c._c() for c in outs])
    N.check(lib.srj_convert_from_unsafe_rows(rows.data_ptr(), None, n, oarr, len(outs), nulls.data_ptr(), ws.data_ptr(), st))
    torch.cuda.synchronize()
    for g, c, k in zip(outs, cols, nulls.cpu().numpy()):
        assert cols_equal(G.to_host(g), c) and int(k) == c.null_count()


def test_errors():
    G = _gpu()
    import srj_b200 as S
    from srj_b200.unsaferow import UnsafeRowConversion as UR
    with pytest.raises(S.CudfException):
        UR.layout([S.DType(S.DType.LIST)])
    with pytest.raises(S.CudfException):
        UR.layout([S.DType(S.DType.INT32)] * 300)


def test_round_trip_at_scale():
    """2 M rows x (INT32, INT64, STRING, DECIMAL128): sizes multiple of 8, offsets monotone, from(to(x)) = x."""
    G = _gpu()
    import srj_b200 as S
    from srj_b200.unsaferow import UnsafeRowConversion as UR
    n = 2_000_000
    g = torch.Generator(device="cuda").manual_seed(5)
    i32 = torch.randint(-2**31, 2**31 - 1, (n,), dtype=torch.int32, device="cuda", generator=g)
    i64 = torch.randint(-2**62, 2**62, (n,), dtype=torch.int64, device="cuda", generator=g)
    lens = torch.randint(0, 24, (n,), dtype=torch.int32, device="cuda", generator=g)
    soff = torch.zeros(n + 1, dtype=torch.int32, device="cuda")
    soff[1:] = torch.cumsum(lens, 0)
    chars = torch.randint(32, 127, (int(soff[-1]),), dtype=torch.uint8, device="cuda", generator=g)
    dec = torch.randint(-2**62, 2**62, (n, 2), dtype=torch.int64, device="cuda", generator=g)
    dec[:, 1] >>= 40                                             # keep some magnitude variety in the high word
    mask = torch.randint(-2**31, 2**31 - 1, ((n + 31) // 32,), dtype=torch.int32, device="cuda", generator=g)
    cols = [S.ColumnVector(S.DType.INT32, n, i32.view(torch.uint8), mask), S.ColumnVector(S.DType.INT64, n, i64.view(torch.uint8), None),
            S.ColumnVector(S.DType.STRING, n, chars, mask.clone(), soff), S.ColumnVector(S.DType.DECIMAL128, n, dec.view(torch.uint8).reshape(-1), None)]
    rows = UR.convertToRows(S.Table(cols))
    ro = rows.offsets.to(torch.int64)
    sizes = ro[1:] - ro[:-1]
    assert bool((sizes % 8 == 0).all()) and bool((sizes >= 8 + 32 + 16).all()) and int(ro[0]) == 0
    back = UR.convertFromRows(rows, [c.dtype for c in cols])
    valid = ((mask.view(torch.int32)[torch.arange(n, device="cuda") // 32] >> (torch.arange(n, device="cuda") % 32)) & 1).bool()
    assert torch.equal(back.columns[0].data.view(torch.int32)[valid], i32[valid])
    assert torch.equal(back.columns[1].data.view(torch.int64), i64)
    assert torch.equal(back.columns[3].data.view(torch.int64).reshape(n, 2), dec)
    blen = back.columns[2].offsets[1:] - back.columns[2].offsets[:-1]
    assert torch.equal(blen[valid], lens[valid]) and bool((blen[~valid] == 0).all())
    assert back.columns[0].getNullCount() == int((~valid).sum())
