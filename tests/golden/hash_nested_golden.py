"""Nested-type hash goldens transcribed from the reference's own tests (src/test/java/.../HashTest.java): LIST / STRUCT
keys for xxhash64 (:444-545) and hive hash (:754-858), plus the murmur identities of :225-270 (a list hashes like the
columns of its elements).  Builders return oracle HCol trees."""
import numpy as np

from oracle import oracle as O

LONG_MD5 = ("A very long (greater than 128 bytes/char string) to test a multi hash-step data point "
            "in the MD5 hash function. This string needed to be longer.")
LONG_HIVE = ("This is a long string (greater than 128 bytes/char string) case to test this "
             "hash function. Just want an abnormal case here to see if any error may happen when"
             "doing the hive hashing")
NAN64_LO, NAN64_HI = 0x7ff0000000000001, 0x7fffffffffffffff     # POSITIVE_DOUBLE_NAN_{LOWER,UPPER}_RANGE
NAN32_LO, NAN32_HI = 0xff800001, 0xffffffff                     # NEGATIVE_FLOAT_NAN_{LOWER,UPPER}_RANGE
INT_MIN, INT_MAX = -2**31, 2**31 - 1


def ints(vals):
    valid = np.array([v is not None for v in vals], bool)
    data = np.array([0 if v is None else v for v in vals], np.int32)
    return O.HCol(O.INT32, data.view(np.uint8), None if valid.all() else O.pack_mask(valid), None, 0, len(vals))


def lists_of(rows, leaf_builder):
    """rows: list of python lists (or None) -> LIST column over leaf_builder(flat elements)."""
    offs, flat, valid = [0], [], []
    for r in rows:
        valid.append(r is not None)
        flat.extend(r or [])
        offs.append(len(flat))
    return O.list_col(offs, leaf_builder(flat), valid=valid)


def int_lists():
    return lists_of([[], [0, -2, 3], [INT_MAX], [5, -6, None], [INT_MIN], None], ints)


def nested_int_lists():          # HashTest.java:486-493
    return lists_of([[], [[0], [-2], [3]], [[INT_MAX]], [[5], [-6, None]], [[INT_MIN]], None], lambda f: lists_of(f, ints))


def string_lists(long=LONG_MD5):
    return lists_of([[None, "a"], ["B\n", ""], ["dE\"Ā\tā", " 휠휡"], [long], [""], None], O.strings_col)


def nested_string_lists(long=LONG_MD5):
    return lists_of([[None, ["a"]], [["B\n", ""]], [["dE\"Ā\tā"], [" 휠휡"]], [[long]], [[""]], None],
                    lambda f: lists_of(f, O.strings_col))


def bits_col(t, bits, width):
    valid = np.array([b is not None for b in bits], bool)
    dt = np.uint32 if width == 4 else np.uint64
    data = np.array([0 if b is None else b for b in bits], dt)
    return O.HCol(t, data.view(np.uint8), None if valid.all() else O.pack_mask(valid), None, 0, len(bits))


def f64(vals):
    return bits_col(O.FLOAT64, [None if v is None else (v if isinstance(v, int) else int(np.float64(v).view(np.uint64))) for v in vals], 8)


def f32(vals):
    return bits_col(O.FLOAT32, [None if v is None else (v if isinstance(v, int) else int(np.float32(v).view(np.uint32))) for v in vals], 4)


def bools(vals):
    valid = np.array([v is not None for v in vals], bool)
    return O.HCol(O.BOOL8, np.array([1 if v else 0 for v in vals], np.uint8), None if valid.all() else O.pack_mask(valid), None, 0, len(vals))


DOUBLES = [0.0, 100.0, -100.0, NAN64_LO, NAN64_HI, None]
FLOATS = [0.0, 100.0, -100.0, NAN32_LO, NAN32_HI, None]


def struct_of_list(long=LONG_MD5, first_null_list=False):
    il = lists_of([[None] if first_null_list else [], [0, -2, 3], [INT_MAX], [5, -6, None], [INT_MIN], None], ints)
    return O.struct_col(il, string_lists(long), f64(DOUBLES), f32(FLOATS))


def list_of_struct(long=LONG_MD5):
    # rows: [], [(a,0,0.0,0f,true)], [(B\n,100,100.0,100f,false), (dE..,-100,-100.0,-100f,null)], [(long,MIN,NaNlo,NaNlo,false)],
    #       [(null,MAX,NaNhi,NaNhi,true), (null,null,null,null,null)], null
    s = O.strings_col(["a", "B\n", "dE\"Ā\tā 휠휡", long, None, None])
    i = ints([0, 100, -100, INT_MIN, INT_MAX, None])
    d = f64([0.0, 100.0, -100.0, NAN64_LO, NAN64_HI, None])
    f = f32([0.0, 100.0, -100.0, NAN32_LO, NAN32_HI, None])
    b = bools([True, False, None, False, True, None])
    st = O.struct_col(s, i, d, f, b)
    return O.list_col([0, 0, 1, 3, 4, 6, 6], st, valid=[1, 1, 1, 1, 1, 0])


XX_CASES = [
    ("int_lists", int_lists, [42, -4022702357093761688, 1508894993788531228, 7329154841501342665, 2073849959933241805, 42]),
    ("nested_int_lists", nested_int_lists, [42, -4022702357093761688, 1508894993788531228, 7329154841501342665, 2073849959933241805, 42]),
    ("string_lists", string_lists, [-8582455328737087284, 7160715839242204087, -862482741676457612, -3700309651391443614, -7444071767201028348, 42]),
    ("nested_string_lists", nested_string_lists, [-8582455328737087284, 7160715839242204087, -862482741676457612, -3700309651391443614, -7444071767201028348, 42]),
    ("struct_of_list", struct_of_list, [-8492741646850220468, -6547737320918905493, -8718220625378038731, 5441580647216064522, 3645801243834961127, 42]),
    ("list_of_struct", list_of_struct, [42, 7451748878409563026, 948372773124634350, 8444697026100086329, -5888679192448042852, 42]),
]


def hive_nested_int_lists():     # HashTest.java:800-807
    return lists_of([[[None, None], None], [[0], [-2], [3]], [None, [INT_MAX]], [[5], [-6, None]], [[INT_MIN], None], None],
                    lambda f: lists_of(f, ints))


def hive_nested_string_lists():  # HashTest.java:785-795
    return lists_of([[None, ["a", None]], [["B\n", ""]], [["dE\"Ā\tā"], [" 휠휡"]], [[LONG_HIVE]], [[""], None], None],
                    lambda f: lists_of(f, O.strings_col))


HIVE_CASES = [
    ("int_lists", int_lists, [0, -59, 2147483647, 4619, -2147483648, 0]),
    ("string_lists", lambda: string_lists(LONG_HIVE), [97, 63736, -96263528, 2112075710, 0, 0]),
    ("nested_int_lists", hive_nested_int_lists, [0, -59, 2147483647, -31, -2147483648, 0]),
    ("nested_string_lists", hive_nested_string_lists, [3007, 63736, -96263528, 2112075710, 0, 0]),
    ("struct_of_list", lambda: struct_of_list(LONG_HIVE, first_null_list=True), [93217, 286968083, 59992121, -1697616301, 2127036416, 0]),
    ("list_of_struct", lambda: list_of_struct(LONG_HIVE), [0, 89581538, -1201635432, 1272817854, -323360610, 0]),
]
