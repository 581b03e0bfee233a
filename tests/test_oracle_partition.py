"""CPU tests of the HashPartitioning oracle: Spark's pmod on the murmur3 goldens of the reference's HashTest.java,
and the stable-partition restatement (offsets, order inside a partition)."""
import numpy as np

from oracle import oracle as O


def test_pmod_matches_the_jvm_definition():
    h = np.array([0, 1, -1, 7, -7, 2**31 - 1, -2**31, 200, -200, 199], dtype=np.int32)
    for n in (1, 2, 7, 200, 32768):
        want = [((int(x) % n) + n) % n if True else 0 for x in h]          # python % is already floored: equals pmod
        assert O.spark_pmod(h, n).tolist() == want


def test_partition_ids_on_reference_goldens():
    """HashTest.java:69-73 and :128-134 (murmur, seed 42, one 4-byte value per row: 0 -> 933211791, 100 -> 751823303,
    -100 -> -1080202046, MIN -> 723455942, MAX -> 133916647): the partition id is Spark's pmod of those."""
    v = np.array([0, 100, -100, -2**31, 2**31 - 1], dtype=np.int32)
    col = O.HCol(O.INT32, v.view(np.uint8))
    h = O.murmur_hash3_32([col], 42).astype(np.int32)
    assert h.tolist() == [933211791, 751823303, -1080202046, 723455942, 133916647]
    assert O.partition_ids([col], 200).tolist() == [x % 200 for x in h.tolist()]
    assert O.partition_ids([col], 7).tolist() == [x % 7 for x in h.tolist()]


def test_stable_partition_keeps_input_order():
    rng = np.random.default_rng(5)
    n, P = 1000, 13
    ids = rng.integers(0, P, n).astype(np.int32)
    vals = np.arange(n, dtype=np.int64)
    strs = O.strings_col([None if i % 11 == 0 else (b"s%d" % i) for i in range(n)])
    cols, offs, gmap = O.stable_partition([O.HCol(O.INT64, vals.view(np.uint8)), strs], ids, P)
    assert offs[0] == 0 and offs[-1] == n and np.array_equal(np.diff(offs), np.bincount(ids, minlength=P))
    got = cols[0].data.view(np.int64)
    for p in range(P):
        seg = got[offs[p]:offs[p + 1]]
        assert np.all(ids[seg] == p) and np.all(np.diff(seg) > 0)         # right partition, input order
    for d in (0, 17, n - 1):
        s = gmap[d]
        assert cols[1].data[cols[1].offsets[d]:cols[1].offsets[d + 1]].tobytes() == strs.data[strs.offsets[s]:strs.offsets[s + 1]].tobytes()
        assert cols[1].valid()[d] == strs.valid()[s]
