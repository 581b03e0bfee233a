"""GPU hash parity: the reference's Spark-derived golden vectors through the CUDA path, plus
random tables against the (golden-pinned) oracle."""
import numpy as np
import pytest

from golden import hash_golden as GOLD
from oracle import oracle as O
from util import cols_from_case, random_table

pytestmark = pytest.mark.gpu


def _gpu():
    import gpu_util
    gpu_util.require_cuda()
    return gpu_util


def _run(kind, cols, seed):
    G = _gpu()
    import srj_b200 as S
    d = [G.to_device(c) for c in cols]
    if kind == "murmur":
        r = S.Hash.murmurHash32(seed, d)
        return r.data.cpu().numpy().view(np.int32)
    if kind == "xxhash64":
        r = S.Hash.xxhash64(seed, d)
        return r.data.cpu().numpy().view(np.int64)
    r = S.Hash.hiveHash(d)
    return r.data.cpu().numpy().view(np.int32)


@pytest.mark.parametrize("case", GOLD.CASES, ids=[c["name"] for c in GOLD.CASES])
def test_gpu_matches_reference_golden(case):
    got = _run(case["kind"], cols_from_case(case), case["seed"])
    assert [int(x) for x in got] == list(case["expected"]), case["src"]


HASH_TYPES = [O.INT8, O.INT16, O.INT32, O.INT64, O.UINT8, O.UINT16, O.UINT32, O.UINT64, O.FLOAT32, O.FLOAT64, O.BOOL8,
              O.TIMESTAMP_DAYS, O.TIMESTAMP_SECONDS, O.TIMESTAMP_MILLISECONDS, O.TIMESTAMP_MICROSECONDS,
              O.TIMESTAMP_NANOSECONDS, O.DECIMAL32, O.DECIMAL64, O.DECIMAL128, O.STRING]


@pytest.mark.parametrize("nrows", [1, 1000, 200_003])
def test_random_tables_xx_mm(nrows):
    cols = random_table(HASH_TYPES, nrows, seed=nrows, max_str=70)     # strings > 32 B hit the 32-byte stripe loop
    for seed in (42, 0, -7):
        assert np.array_equal(_run("xxhash64", cols, seed), O.xxhash64(cols, seed))
        assert np.array_equal(_run("murmur", cols, seed & 0xFFFFFFFF), O.murmur_hash3_32(cols, seed & 0xFFFFFFFF))


def test_random_tables_hive():
    types = [O.BOOL8, O.INT8, O.INT16, O.INT32, O.INT64, O.FLOAT32, O.FLOAT64, O.TIMESTAMP_DAYS,
             O.TIMESTAMP_MICROSECONDS, O.STRING]
    cols = random_table(types, 50_001, seed=2)
    assert np.array_equal(_run("hive", cols, 0), O.hive_hash(cols))


def test_wide_table_chains_column_chunks():
    """More columns than one launch carries (48): chunks chain through the output accumulator."""
    cols = random_table([O.INT32, O.INT64, O.STRING, O.FLOAT64] * 40, 3001, seed=8)
    assert np.array_equal(_run("xxhash64", cols, 42), O.xxhash64(cols, 42))
    assert np.array_equal(_run("murmur", cols, 42), O.murmur_hash3_32(cols, 42))
    hive_cols = [c for c in cols]
    assert np.array_equal(_run("hive", hive_cols, 0), O.hive_hash(hive_cols))


def test_unsupported_and_default_seeds():
    G = _gpu()
    import srj_b200 as S
    c = G.to_device(random_table([O.DECIMAL64], 10)[0])
    with pytest.raises(S.CudfException):
        S.Hash.hiveHash([c])                         # hive_hash.cu:63-66
    i = G.to_device(random_table([O.INT32], 10, null_frac=0)[0])
    assert np.array_equal(S.Hash.xxhash64([i]).data.cpu().numpy(), S.Hash.xxhash64(42, [i]).data.cpu().numpy())
    assert np.array_equal(S.Hash.murmurHash32([i]).data.cpu().numpy(), S.Hash.murmurHash32(0, [i]).data.cpu().numpy())
    assert S.Hash.getMaxStackDepth() == 8


@pytest.mark.parametrize("nrows", [8192 * 3 + 777, 8192 * 2])
@pytest.mark.parametrize("name", ["int32_int64", "all_fixed", "floats_decimals", "no_masks"])
def test_streaming_hash_kernel_matches_oracle(name, nrows):
    """Tables large enough for row_hash_stream_kernel (whole 2048-row chunks staged by TMA) + the tail rows on the
    per-thread kernels; every fixed-width key type, with and without null masks."""
    import gpu_util as G
    import numpy as np
    import srj_b200 as S
    from oracle import oracle as O
    from util import random_table
    G.require_cuda()
    schemas = {
        "int32_int64": [O.INT32, O.INT64],
        "all_fixed": [O.BOOL8, O.INT8, O.INT16, O.INT32, O.INT64, O.FLOAT32, O.FLOAT64, O.TIMESTAMP_DAYS, O.TIMESTAMP_MICROSECONDS],
        "floats_decimals": [O.FLOAT64, O.DECIMAL32, O.DECIMAL64, O.DECIMAL128, O.FLOAT32, O.UINT32, O.UINT64],
        "no_masks": [O.INT64, O.INT32, O.INT16],
    }
    types = schemas[name]
    cols = random_table(types, nrows, seed=nrows % 97 + len(types), null_frac=0.0 if name == "no_masks" else 0.2)
    for c in cols:
        if c.type_id in (O.FLOAT32, O.FLOAT64):
            v = c.data.view(np.float32 if c.type_id == O.FLOAT32 else np.float64)
            v[::5] = np.nan
            v[1::5] = -0.0
    dk = [G.to_device(c) for c in cols]
    assert np.array_equal(S.Hash.xxhash64(42, dk).data.cpu().numpy().view(np.int64), O.xxhash64(cols, 42))
    assert np.array_equal(S.Hash.murmurHash32(42, dk).data.cpu().numpy().view(np.int32), O.murmur_hash3_32(cols, 42))
    hive_ok = [c for c in cols if c.type_id not in (O.DECIMAL32, O.DECIMAL64, O.DECIMAL128, O.UINT32, O.UINT64)]
    if hive_ok:
        dh = [G.to_device(c) for c in hive_ok]
        assert np.array_equal(S.Hash.hiveHash(dh).data.cpu().numpy().view(np.int32), O.hive_hash(hive_ok))
