"""GPU-test plumbing: host (numpy/oracle HCol) <-> device (srj_b200.ColumnVector) conversion."""
from __future__ import annotations

import numpy as np
import torch

import srj_b200 as S
from oracle import oracle as O


def require_cuda():
    # '-m gpu' tests must FAIL (not skip) if the CUDA path cannot run: a silent fallback is a bug
    assert torch.cuda.is_available(), "CUDA device required for -m gpu tests"


def to_device(c: O.HCol) -> S.ColumnVector:
    if c.type_id in (O.LIST, O.STRUCT):
        mask = torch.from_numpy(c.mask.view(np.int32).copy()).cuda() if c.mask is not None else None
        kids = [to_device(k) for k in (c.children or [])]
        if c.type_id == O.LIST:
            offs = torch.from_numpy(np.ascontiguousarray(c.offsets, dtype=np.int32)).cuda()
            return S.ColumnVector(S.DType.LIST, c.size, None, mask, offs, kids[0])
        return S.ColumnVector(S.DType.STRUCT, c.size, None, mask, None, None, None, children=kids)
    return S.ColumnVector.from_numpy(c.type_id, c.data if c.data is not None else np.zeros(0, np.uint8), c.mask,
                                     c.offsets, c.scale, c.size)


def to_host(c: S.ColumnView) -> O.HCol:
    d, m, o = c.to_numpy()
    if c.dtype.type_id == S.DType.STRING and d is None:
        d = np.zeros(0, np.uint8)
    return O.HCol(c.dtype.type_id, d, m, o, c.dtype.scale, c.size)


def table_to_device(cols) -> S.Table:
    return S.Table([to_device(c) for c in cols])


def rows_to_device(offs: np.ndarray, data: np.ndarray) -> S.ColumnVector:
    o = torch.from_numpy(np.ascontiguousarray(offs, dtype=np.int32)).cuda()
    d = torch.from_numpy(np.ascontiguousarray(data, dtype=np.uint8)).cuda()
    return S.ColumnVector(S.DType.LIST, len(offs) - 1, None, None, o, S.ColumnVector(S.DType.INT8, len(data), d))


def rows_to_host(v: S.ColumnView):
    return v.offsets.cpu().numpy(), v.child.data.cpu().numpy()
