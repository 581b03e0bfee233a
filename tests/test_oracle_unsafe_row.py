"""Known answers of the UnsafeRow restatement (oracle/unsafe_row.py), derived by hand from the rules of Apache Spark's
UnsafeRow.java / UnsafeRowWriter.java (no vector of this format exists in the reference repo), and its round trip."""
import numpy as np

from oracle import oracle as O
from oracle import unsafe_row as U
from util import cols_equal, random_table


def _i(t, vals, valid=None):
    arr = np.array(vals, dtype=O.np_dtype(t))
    return O.HCol(t, arr.view(np.uint8), None if valid is None else O.pack_mask(np.array(valid, bool)))


def test_known_answer_int_string_null_long():
    """Row (INT32 1, STRING "ab", INT64 NULL): bitset 0b100 | 01 00.. | (24 << 32) | 2 | 0 | "ab" + 6 zero bytes."""
    cols = [_i(O.INT32, [1]), O.strings_col([b"ab"]), _i(O.INT64, [77], [False])]
    offs, data = U.to_unsafe_rows(cols)
    want = (bytes([0b100]) + bytes(7) +                      # null bitset: field 2 is null
            (1).to_bytes(8, "little") +                      # int 1 in the low 4 bytes of its slot
            ((32 << 32) | 2).to_bytes(8, "little") +         # string: offset 32 (8 + 3 * 8), 2 bytes
            bytes(8) +                                       # null long: slot 0
            b"ab" + bytes(6))
    assert offs.tolist() == [0, 40] and data.tobytes() == want


def test_known_answer_small_types_and_negative_int():
    """(BOOL8 true, INT8 -2, INT16 -3, INT32 -4, FLOAT32 1.0): values in the low bytes, the rest of every slot zero
    (ints are NOT sign-extended; UnsafeRowWriter zeroes the slot first)."""
    cols = [_i(O.BOOL8, [1]), _i(O.INT8, [-2]), _i(O.INT16, [-3]), _i(O.INT32, [-4]), _i(O.FLOAT32, [1.0])]
    _, data = U.to_unsafe_rows(cols)
    want = bytes(8) + b"\x01" + bytes(7) + b"\xfe" + bytes(7) + b"\xfd\xff" + bytes(6) + b"\xfc\xff\xff\xff" + bytes(4) + \
        b"\x00\x00\x80\x3f" + bytes(4)
    assert data.tobytes() == want


def test_known_answer_decimals():
    """DECIMAL32 -5 -> the long -5; DECIMAL128 values -> 16 reserved bytes with BigInteger.toByteArray():
    0 -> 00; 127 -> 7f; 128 -> 00 80; -128 -> 80; -129 -> ff 7f; NULL -> size 0, offset kept, null bit set."""
    d32 = _i(O.DECIMAL32, [-5] * 6)
    raw = b"".join(v.to_bytes(16, "little", signed=True) for v in (0, 127, 128, -128, -129, 99))
    d128 = O.HCol(O.DECIMAL128, np.frombuffer(raw, dtype=np.uint8).copy(), O.pack_mask(np.array([1, 1, 1, 1, 1, 0], bool)))
    offs, data = U.to_unsafe_rows([d32, d128])
    assert np.diff(offs).tolist() == [8 + 16 + 16] * 6
    rows = [data[offs[i]:offs[i + 1]].tobytes() for i in range(6)]
    payloads = [b"\x00", b"\x7f", b"\x00\x80", b"\x80", b"\xff\x7f", b""]
    for i, (row, pl) in enumerate(zip(rows, payloads)):
        assert row[0] == (0b10 if i == 5 else 0)
        assert row[8:16] == (-5).to_bytes(8, "little", signed=True)
        assert row[16:24] == ((24 << 32) | len(pl)).to_bytes(8, "little")
        assert row[24:40] == pl + bytes(16 - len(pl))


def test_bitset_beyond_64_fields():
    cols = [_i(O.INT8, [i], [i != 64 and i != 3]) for i in range(70)]
    offs, data = U.to_unsafe_rows(cols)
    assert offs.tolist() == [0, 16 + 70 * 8]
    assert int.from_bytes(data[:8].tobytes(), "little") == 1 << 3 and int.from_bytes(data[8:16].tobytes(), "little") == 1 << 0


def test_round_trip():
    types = [O.INT32, O.STRING, O.DECIMAL128, O.INT64, O.BOOL8, O.STRING, O.DECIMAL64, O.FLOAT64, O.INT16, O.DECIMAL32]
    cols = random_table(types, 300, seed=11)
    offs, data = U.to_unsafe_rows(cols)
    assert np.all(np.diff(offs) % 8 == 0)
    back = U.from_unsafe_rows(data, offs, types)
    for a, b in zip(cols, back):
        assert cols_equal(a, b)


def test_vectorised_fixed_width_form_equals_the_loop():
    types = [O.INT8, O.INT16, O.INT32, O.INT64, O.FLOAT32, O.FLOAT64, O.BOOL8, O.DECIMAL32, O.DECIMAL64] * 8     # 72 fields: two bitset words
    cols = random_table(types, 257, seed=13)
    _, data = U.to_unsafe_rows(cols)
    assert np.array_equal(U.to_unsafe_rows_fixed(cols).reshape(-1), data)
