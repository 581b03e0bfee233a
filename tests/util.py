"""Shared test helpers: golden-case loader and seeded synthetic tables (numpy, host side)."""
from __future__ import annotations

import numpy as np

from oracle import oracle as O

TYPE_BY_NAME = {n: getattr(O, n) for n in (
    "INT8 INT16 INT32 INT64 UINT8 UINT16 UINT32 UINT64 FLOAT32 FLOAT64 BOOL8 TIMESTAMP_DAYS "
    "TIMESTAMP_SECONDS TIMESTAMP_MILLISECONDS TIMESTAMP_MICROSECONDS TIMESTAMP_NANOSECONDS "
    "DURATION_DAYS DURATION_SECONDS DURATION_MILLISECONDS DURATION_MICROSECONDS DURATION_NANOSECONDS "
    "STRING DECIMAL32 DECIMAL64 DECIMAL128").split()}


def col_from_values(tname: str, values, scale: int = 0) -> O.HCol:
    """Build a host column from python values (None = null; ("bits32"/"bits64", int) = raw float bits)."""
    t = TYPE_BY_NAME[tname]
    n = len(values)
    valid = np.array([v is not None for v in values], dtype=bool)
    mask = None if valid.all() else O.pack_mask(valid)
    if t == O.STRING:
        c = O.strings_col(values)
        return c
    sz = O.size_of(t)
    buf = np.zeros(n * sz, dtype=np.uint8)
    for i, v in enumerate(values):
        if v is None:
            continue
        if isinstance(v, tuple):
            b = int(v[1]).to_bytes(sz, "little")
        elif t == O.FLOAT32:
            b = np.float32(v).tobytes()
        elif t == O.FLOAT64:
            b = np.float64(v).tobytes()
        elif t in (O.BOOL8, O.UINT8, O.UINT16, O.UINT32, O.UINT64):
            b = int(v).to_bytes(sz, "little", signed=False)
        else:
            b = int(v).to_bytes(sz, "little", signed=True)
        buf[i * sz:(i + 1) * sz] = np.frombuffer(b, dtype=np.uint8)
    return O.HCol(t, buf, mask, None, scale, n)


def cols_from_case(case) -> list:
    out = []
    for spec in case["cols"]:
        tname, vals = spec[0], spec[1]
        scale = spec[2] if len(spec) > 2 else 0
        out.append(col_from_values(tname, vals, scale))
    return out


# ------------------------------------------------------------------------------------------------
# seeded synthetic tables (SURVEY.md 8d canonical inputs, numpy Philox so CPU/GPU tests agree)
# ------------------------------------------------------------------------------------------------
def random_table(types, nrows: int, seed: int = 42, null_frac: float = 0.2, max_str: int = 32,
                 all_valid_cols=()):
    rng = np.random.Generator(np.random.Philox(seed))
    cols = []
    for ci, t in enumerate(types):
        if null_frac > 0 and ci not in all_valid_cols:
            valid = rng.random(nrows) >= null_frac
            mask = O.pack_mask(valid)
        else:
            valid = np.ones(nrows, bool)
            mask = None
        if t == O.STRING:
            lens = np.clip(np.rint(rng.normal(16, 8, nrows)), 0, max_str).astype(np.int64)
            lens[~valid] = 0          # null strings have length 0
            offs = np.zeros(nrows + 1, dtype=np.int32)
            np.cumsum(lens, out=offs[1:])
            chars = rng.integers(32, 127, int(offs[-1]), dtype=np.uint8)
            # ~5% multibyte: overwrite some byte pairs with a 2-byte UTF-8 sequence (0xC3 0xA9 = e-acute)
            if len(chars) > 2:
                idx = rng.integers(0, len(chars) - 1, max(1, len(chars) // 40))
                # keep sequences inside one string is not required for byte-exact tests
                chars[idx] = 0xC3
                chars[idx + 1] = 0xA9
            cols.append(O.HCol(t, chars, mask, offs, 0, nrows))
        else:
            sz = O.size_of(t)
            raw = rng.integers(0, 256, nrows * sz, dtype=np.uint8)
            if t == O.BOOL8:
                raw = (raw & 1).astype(np.uint8)
            scale = -11 if t == O.DECIMAL128 else (-2 if t in (O.DECIMAL32, O.DECIMAL64) else 0)
            cols.append(O.HCol(t, raw, mask, None, scale, nrows))
    return cols


def cols_equal(a: O.HCol, b: O.HCol, check_null_payload: bool = False) -> bool:
    """Column equality the way cudf's CUDF_TEST_EXPECT_COLUMNS_EQUAL defines it for this path:
    same type/size, same validity, equal values where valid (null payload bytes are undefined)."""
    if a.type_id != b.type_id or a.size != b.size:
        return False
    va, vb = a.valid(), b.valid()
    if not np.array_equal(va, vb):
        return False
    if a.type_id == O.STRING:
        la = np.diff(a.offsets.astype(np.int64)); lb = np.diff(b.offsets.astype(np.int64))
        if not np.array_equal(la[va], lb[va]):
            return False
        for r in np.nonzero(va)[0]:
            if a.data[a.offsets[r]:a.offsets[r + 1]].tobytes() != b.data[b.offsets[r]:b.offsets[r + 1]].tobytes():
                return False
        return True
    sz = O.size_of(a.type_id)
    da = np.ascontiguousarray(a.data).view(np.uint8).reshape(a.size, sz)
    db = np.ascontiguousarray(b.data).view(np.uint8).reshape(b.size, sz)
    if check_null_payload:
        return np.array_equal(da, db)
    return np.array_equal(da[va], db[va])
