"""End-to-end host-buffer entry point (srj_convert_from_rows_host): chunked H2D -> kernel -> D2H."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O
from util import random_table

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nrows,chunk", [(1, 0), (100_001, 32768), (300_000, 0)])
def test_host_from_rows(nrows, chunk):
    import gpu_util
    gpu_util.require_cuda()
    import srj_b200 as S
    from srj_b200 import _native as N
    types = [O.INT8, O.INT16, O.INT32, O.INT64, O.FLOAT32, O.FLOAT64, O.BOOL8, O.TIMESTAMP_MICROSECONDS] * 4
    cols = random_table(types, nrows, seed=11)
    (offs, data), = O.convert_to_rows(cols)
    plan = S.Plan.get([S.DType(t) for t in types])
    outs = []
    arr = (N.SrjColumn * len(types))()
    for i, t in enumerate(types):
        d = np.zeros(nrows * O.size_of(t), np.uint8)
        m = np.zeros((nrows + 31) // 32, np.uint32)
        outs.append((d, m))
        arr[i].type_id, arr[i].scale, arr[i].size = t, 0, nrows
        arr[i].data, arr[i].null_mask, arr[i].offsets = d.ctypes.data, m.ctypes.data, None
    nulls = np.zeros(len(types), np.int64)
    data = np.ascontiguousarray(data)
    N.check(N.lib().srj_convert_from_rows_host(plan.handle, data.ctypes.data, nrows, arr, nulls.ctypes.data, chunk))
    ocols, onulls = O.convert_from_rows(data, None, nrows, types)
    for (d, m), o in zip(outs, ocols):
        assert np.array_equal(d, np.ascontiguousarray(o.data).view(np.uint8))
        assert np.array_equal(m, o.mask)
    assert np.array_equal(nulls, onulls)
