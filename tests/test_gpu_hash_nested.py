"""LIST / STRUCT keys through the CUDA path: the reference's nested goldens, random nested tables against the
(golden-pinned) oracle, and the reference's error behaviour (depth limit, murmur LIST<STRUCT>)."""
import numpy as np
import pytest

from golden import hash_nested_golden as NG
from oracle import oracle as O
from util import random_table

pytestmark = pytest.mark.gpu


def _gpu():
    import gpu_util
    gpu_util.require_cuda()
    return gpu_util


@pytest.mark.parametrize("name,build,want", NG.XX_CASES, ids=[c[0] for c in NG.XX_CASES])
def test_xxhash64_nested_goldens(name, build, want):
    G = _gpu()
    import srj_b200 as S
    got = S.Hash.xxhash64(42, [G.to_device(build())]).data.cpu().numpy().view(np.int64)
    assert got.tolist() == want


@pytest.mark.parametrize("name,build,want", NG.HIVE_CASES, ids=[c[0] for c in NG.HIVE_CASES])
def test_hive_nested_goldens(name, build, want):
    G = _gpu()
    import srj_b200 as S
    got = S.Hash.hiveHash([G.to_device(build())]).data.cpu().numpy().view(np.int32)
    assert got.tolist() == want


def _random_nested(n, seed):
    rng = np.random.default_rng(seed)
    flat = random_table([O.INT32, O.STRING, O.INT64, O.FLOAT64, O.BOOL8, O.INT16], n, seed=seed)

    def rand_list(child_n_builder, rows):
        lens = rng.integers(0, 5, rows)
        offs = np.zeros(rows + 1, np.int32)
        np.cumsum(lens, out=offs[1:])
        valid = rng.random(rows) > 0.15
        return offs, int(offs[-1]), valid
    # LIST<INT32>
    o1, m1, v1 = rand_list(None, n)
    l_int = O.list_col(o1, random_table([O.INT32], m1, seed=seed + 1)[0], valid=v1)
    # LIST<LIST<STRING>>
    o2, m2, v2 = rand_list(None, n)
    o3, m3, v3 = rand_list(None, m2)
    l_ls = O.list_col(o2, O.list_col(o3, random_table([O.STRING], m3, seed=seed + 2)[0], valid=v3), valid=v2)
    # STRUCT<INT64, STRUCT<STRING, FLOAT64>, LIST<INT32>>
    inner = O.struct_col(flat[1], flat[3])
    st = O.struct_col(flat[2], inner, l_int)
    # LIST<STRUCT<INT32, STRING>>
    o4, m4, v4 = rand_list(None, n)
    e = random_table([O.INT32, O.STRING], m4, seed=seed + 3)
    l_st = O.list_col(o4, O.struct_col(e[0], e[1]), valid=v4)
    return flat, l_int, l_ls, st, l_st


@pytest.mark.parametrize("n", [1, 257, 3001])
def test_random_nested_tables_match_oracle(n):
    G = _gpu()
    import srj_b200 as S
    flat, l_int, l_ls, st, l_st = _random_nested(n, seed=n + 3)
    keysets = {"mixed": [flat[0], l_int, st, flat[4]], "lists": [l_ls, l_int], "list_of_struct": [l_st, flat[5]]}
    for name, keys in keysets.items():
        dk = [G.to_device(k) for k in keys]
        assert np.array_equal(S.Hash.xxhash64(42, dk).data.cpu().numpy().view(np.int64), O.nested_hash("xxhash64", keys, 42)), name
        assert np.array_equal(S.Hash.hiveHash(dk).data.cpu().numpy().view(np.int32), O.nested_hash("hive", keys)), name
        if name != "list_of_struct":
            assert np.array_equal(S.Hash.murmurHash32(7, dk).data.cpu().numpy().view(np.int32), O.nested_hash("murmur3", keys, 7)), name


def test_nested_error_behaviour():
    G = _gpu()
    import srj_b200 as S
    flat, l_int, l_ls, st, l_st = _random_nested(64, seed=5)
    with pytest.raises(S.CudfException):                     # murmur_hash.cu:173-175
        S.Hash.murmurHash32(0, [G.to_device(l_st)])
    deep = flat[0]
    for _ in range(9):                                       # HashTest.java:547-574: nesting beyond MAX_STACK_DEPTH
        deep = O.struct_col(deep)
    with pytest.raises(S.CudfException):
        S.Hash.xxhash64(42, [G.to_device(deep)])
    with pytest.raises(S.CudfException):
        S.Hash.hiveHash([G.to_device(deep)])
    ok = flat[0]
    for _ in range(7):
        ok = O.struct_col(ok)
    assert np.array_equal(S.Hash.xxhash64(42, [G.to_device(ok)]).data.cpu().numpy().view(np.int64), O.nested_hash("xxhash64", [ok], 42))
