"""The oracle's nested-type (LIST / STRUCT) hash restatement against the reference's own goldens (HashTest.java)."""
import numpy as np
import pytest

from golden import hash_nested_golden as NG
from oracle import oracle as O


@pytest.mark.parametrize("name,build,want", NG.XX_CASES, ids=[c[0] for c in NG.XX_CASES])
def test_xxhash64_nested_goldens(name, build, want):
    assert O.nested_hash("xxhash64", [build()], 42).tolist() == want


@pytest.mark.parametrize("name,build,want", NG.HIVE_CASES, ids=[c[0] for c in NG.HIVE_CASES])
def test_hive_nested_goldens(name, build, want):
    assert O.nested_hash("hive", [build()]).tolist() == want


def test_murmur_lists_hash_like_their_elements():
    """HashTest.java:225-270: a list of ints hashes like the columns of its elements; a struct like its fields."""
    il = NG.lists_of([None, [0, -2, 3], [NG.INT_MAX], [5, -6, None], [NG.INT_MIN], None], NG.ints)
    c1, c2, c3 = NG.ints([None, 0, None, 5, NG.INT_MIN, None]), NG.ints([None, -2, NG.INT_MAX, None, None, None]), NG.ints([None, 3, None, -6, None, None])
    assert np.array_equal(O.nested_hash("murmur3", [il], 1868), O.murmur_hash3_32([c1, c2, c3], 1868))
    st = O.struct_col(c1, c2, c3)
    assert np.array_equal(O.nested_hash("murmur3", [st], 1868), O.murmur_hash3_32([c1, c2, c3], 1868))
