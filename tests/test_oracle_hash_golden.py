"""Pins the CPU oracle's hashes against the reference's Spark-derived golden vectors
(tests/golden/hash_golden.py, transcribed from hash.cpp / HashTest.java) and against the
independent `xxhash` python package (standard XXH64 == Spark's)."""
import numpy as np
import pytest

from golden import hash_golden as G
from oracle import oracle as O
from util import cols_from_case


@pytest.mark.parametrize("case", G.CASES, ids=[c["name"] for c in G.CASES])
def test_oracle_matches_reference_golden(case):
    cols = cols_from_case(case)
    if case["kind"] == "murmur":
        got = O.murmur_hash3_32(cols, case["seed"])
    elif case["kind"] == "xxhash64":
        got = O.xxhash64(cols, case["seed"])
    else:
        got = O.hive_hash(cols)
    assert [int(x) for x in got] == list(case["expected"]), case["src"]


def test_oracle_xxh64_is_standard_xxh64():
    xxhash = pytest.importorskip("xxhash")
    rng = np.random.default_rng(7)
    for n in list(range(0, 100)) + [255, 256, 1000, 4099]:
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        for seed in (0, 42, 2**63 + 5):
            assert O.xxh64_bytes(b, seed) == xxhash.xxh64(b, seed=seed).intdigest()


def test_unsupported_types():
    lst = O.HCol(O.LIST, np.zeros(4, np.uint8), None, None, 0, 1)
    with pytest.raises(NotImplementedError):
        O.xxhash64([lst])
    dec = O.HCol(O.DECIMAL32, np.zeros(4, np.uint8), None, None, 0, 1)
    with pytest.raises(NotImplementedError):
        O.hive_hash([dec])     # hive_hash.cu:63-66
