"""Known answers of the Kudo wire-format restatement (oracle/kudo.py): the header of KudoSerializerTest.java:77-87,
hand-derived partitions following KudoSerializer.java:49-171 / SlicedValidityBufferInfo.java:63-77, and
split -> assemble = identity."""
import struct

import numpy as np

from oracle import kudo as K
from oracle import oracle as O
from util import cols_equal, random_table


def test_header_of_a_table_without_columns():
    """KudoSerializerTest.java:70-87: no columns, 5 rows -> 28 bytes, every length 0."""
    b = K.write_partition([], 0, 5)
    assert len(b) == 28 and struct.unpack(">7i", b) == (0x4B554430, 0, 5, 0, 0, 0, 0) and b[:4] == b"KUD0"


def test_hand_derived_partition():
    """rows [3, 9) of (INT32 with nulls, STRING without nulls):
    header 28 + 1 bitset byte (bit 0) = 29; validity: bytes 0..1 of the mask (rows 3..8) = 2 bytes, padded so that
    29 + 2 -> 32: validityBufferLen 3; offsets: 7 ints = 28; data: 6 * 4 int bytes + the chars of rows 3..8, padded to 4."""
    ints = np.arange(10, dtype=np.int32) * 3
    valid = np.array([1, 1, 0, 1, 1, 1, 0, 1, 1, 1], bool)
    c0 = O.HCol(O.INT32, ints.view(np.uint8), O.pack_mask(valid))
    c1 = O.strings_col([b"a", b"bb", b"", b"dddd", b"e", b"ff", b"g", b"", b"iii", b"j"])
    b = K.write_partition([c0, c1], 3, 6)
    chars = b"dddd" + b"e" + b"ff" + b"g" + b"" + b"iii"            # 11 bytes
    magic, off, n, vlen, olen, total, nc = struct.unpack(">7i", b[:28])
    assert (magic, off, n, vlen, olen, nc) == (0x4B554430, 3, 6, 3, 28, 2)
    assert total == 3 + 28 + (24 + 11 + 1) and len(b) == 29 + total
    assert b[28] == 0b01
    assert b[29:31] == O.pack_mask(valid).view(np.uint8)[0:2].tobytes() and b[31] == 0
    assert np.frombuffer(b[32:60], dtype="<i4").tolist() == c1.offsets[3:10].tolist()          # raw, not rebased
    assert b[60:84] == ints[3:9].tobytes() and b[84:95] == chars and b[95] == 0


def test_validity_slice_lengths():
    """SlicedValidityBufferInfo.calc: (rowOffset + numRows - 1) / 8 - rowOffset / 8 + 1 bytes from byte rowOffset / 8."""
    valid = np.ones(64, bool)
    valid[::3] = False
    c = O.HCol(O.INT8, np.arange(64, dtype=np.int8).view(np.uint8), O.pack_mask(valid))
    for off, n, want in ((0, 8, 1), (0, 9, 2), (7, 1, 1), (7, 2, 2), (8, 8, 1), (5, 20, 4), (63, 1, 1)):
        b = K.write_partition([c], off, n)
        vlen = struct.unpack(">7i", b[:28])[3]
        assert vlen == ((want + 29 + 3) & ~3) - 29
        assert b[29:29 + want] == O.pack_mask(valid).view(np.uint8)[off // 8: off // 8 + want].tobytes()
    assert struct.unpack(">7i", K.write_partition([c], 10, 0)[:28])[3:6] == (3, 0, 3)          # no rows: no validity, padding only


def test_split_assemble_round_trip():
    types = [O.INT32, O.STRING, O.INT64, O.DECIMAL128, O.INT8, O.STRING, O.FLOAT64]
    cols = random_table(types, 1000, seed=4)
    splits = [0, 0, 17, 17, 300, 301, 640, 1000, 1000]
    buf, offs = K.split(cols, splits)
    assert np.all(np.diff(offs) % 4 == 0) and offs[-1] == len(buf)
    back = K.assemble(buf, offs, types)
    for a, b in zip(cols, back):
        assert cols_equal(a, b)
