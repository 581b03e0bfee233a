"""The host-buffer entry points (srj_convert_from_rows_host / srj_convert_to_rows_host): host rows <-> host columns in
one C call each, against the oracle; fixed-width (chunk-pipelined), STRING schemas (whole-row and wide paths), repeated
calls reusing the plan's device staging, concurrent callers."""
import threading

import numpy as np
import pytest
import torch

from oracle import oracle as O
from util import cols_equal, random_table

pytestmark = pytest.mark.gpu

SCHEMAS = {
    "c1": [O.INT32, O.INT64, O.FLOAT64, O.BOOL8],
    "c2": [O.INT8, O.INT16, O.INT32, O.INT64, O.FLOAT32, O.FLOAT64, O.BOOL8, O.TIMESTAMP_MICROSECONDS] * 4,
    "mixed": [O.INT32, O.STRING, O.INT64, O.DECIMAL128, O.STRING, O.BOOL8, O.STRING, O.INT16],
    "c3": [O.INT32, O.INT64, O.DECIMAL128, O.STRING] * 64,
}


def _hcols(out, types, n):
    cols = []
    for i, t in enumerate(types):
        mask = out.mask[i].numpy().view(np.uint32)
        if t == O.STRING:
            cols.append(O.HCol(t, out.data[i].numpy(), mask, out.offsets[i].numpy(), 0, n))
        else:
            cols.append(O.HCol(t, out.data[i].numpy(), mask, None, 0, n))
    return cols


@pytest.mark.parametrize("nrows", [1, 1000, 70_001])
@pytest.mark.parametrize("name", sorted(SCHEMAS))
def test_host_from_rows_and_to_rows(name, nrows):
    import gpu_util as G
    import srj_b200 as S
    from srj_b200 import hostpath
    G.require_cuda()
    types = SCHEMAS[name]
    if len(types) * nrows > 3_000_000:
        nrows = 3_000_000 // len(types)
    cols = random_table(types, nrows, seed=nrows + 5)
    (offs, data), = O.convert_to_rows(cols)
    dts = [S.DType(t) for t in types]
    has_str = O.STRING in types
    out = hostpath.convert_from_rows_host(torch.from_numpy(data), torch.from_numpy(offs) if has_str else None, nrows, dts,
                                          chunk_rows=4096 if name == "c2" else 0)
    want, nulls = O.convert_from_rows(data, offs if has_str else None, nrows, types)
    got = _hcols(out, types, nrows)
    for i, (g, w) in enumerate(zip(got, want)):
        assert np.array_equal(g.mask, w.mask), f"mask, column {i}"
        if types[i] == O.STRING:
            assert np.array_equal(g.offsets, w.offsets) and np.array_equal(g.data, w.data), f"string column {i}"
        else:
            assert cols_equal(g, w, check_null_payload=True), f"column {i}"
    assert np.array_equal(out.null_counts.numpy(), nulls)
    # ... and back: host columns -> host rows
    hc = [(torch.from_numpy(np.ascontiguousarray(c.data).view(np.uint8).copy()) if c.data is not None else None,
           torch.from_numpy(c.mask.view(np.int32).copy()) if c.mask is not None else None,
           torch.from_numpy(c.offsets.copy()) if c.offsets is not None else None) for c in cols]
    rows = hostpath.convert_to_rows_host(hc, dts, nrows)
    assert len(rows) == 1
    assert np.array_equal(rows[0][0].numpy(), offs)
    assert np.array_equal(rows[0][1].numpy(), data)


def test_host_path_reuses_staging_and_is_reentrant():
    """Repeated calls with growing and shrinking tables (the pooled device staging only grows) and 6 threads at once
    (more callers than pool entries)."""
    import gpu_util as G
    import srj_b200 as S
    from srj_b200 import hostpath
    G.require_cuda()
    types = SCHEMAS["mixed"]
    dts = [S.DType(t) for t in types]
    errs = []

    def work(tid):
        try:
            for n in (500 + tid, 40_000 + tid, 3, 12_345):
                cols = random_table(types, n, seed=tid * 31 + n)
                (offs, data), = O.convert_to_rows(cols)
                out = hostpath.convert_from_rows_host(torch.from_numpy(data), torch.from_numpy(offs), n, dts, pin=False)
                for g, c in zip(_hcols(out, types, n), cols):
                    assert cols_equal(g, c), f"thread {tid} n {n}"
        except Exception as ex:      # noqa: BLE001
            errs.append((tid, repr(ex)))
    th = [threading.Thread(target=work, args=(t,)) for t in range(6)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs[:2]
