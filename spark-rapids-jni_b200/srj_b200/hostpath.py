"""Host-buffer entry points of the C ABI (srj_convert_from_rows_host / srj_convert_to_rows_host) behind the host mirror:
what a caller holding HOST memory uses -- H2D, the conversion and D2H all happen inside one C call.  Buffers are torch
CPU tensors (pinned for full PCIe speed); the library asks for the output buffers whose size it only learns during the
call (STRING chars, row batches) through the allocator callback, exactly as the JNI shim would hand out
HostMemoryBuffers."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import DType, Plan, _as_dtype, _native as N


def _pinned(n: int, dtype=torch.uint8, pin: bool = True) -> torch.Tensor:
    return torch.empty(int(n), dtype=dtype, pin_memory=pin)


class HostColumns:
    """Output of convert_from_rows_host: per column (data, mask, offsets) CPU tensors + null counts."""

    def __init__(self):
        self.data: List[Optional[torch.Tensor]] = []
        self.mask: List[torch.Tensor] = []
        self.offsets: List[Optional[torch.Tensor]] = []
        self.null_counts = None


def convert_from_rows_host(h_rows: torch.Tensor, h_offsets: Optional[torch.Tensor], num_rows: int, schema: Sequence,
                           out: Optional[HostColumns] = None, pin: bool = True, chunk_rows: int = 0) -> HostColumns:
    """JCUDF rows in host memory -> columns in host memory (one C call).  Pass `out` to reuse buffers across calls (its
    STRING data tensors are reused when large enough)."""
    dts = [_as_dtype(d) for d in schema]
    plan = Plan.get(dts)
    n, nc = int(num_rows), len(dts)
    words = (n + 31) // 32
    fresh = out is None
    if fresh:
        out = HostColumns()
        for d in dts:
            out.mask.append(_pinned(words, torch.int32, pin))
            if d.type_id == DType.STRING:
                out.data.append(None)
                out.offsets.append(_pinned(n + 1, torch.int32, pin))
            else:
                out.data.append(_pinned(n * d.size_in_bytes(), torch.uint8, pin))
                out.offsets.append(None)
    arr = (N.SrjColumn * max(1, nc))()
    for i, d in enumerate(dts):
        arr[i].type_id, arr[i].scale, arr[i].size = d.type_id, d.scale, n
        arr[i].null_mask = out.mask[i].data_ptr()
        if d.type_id == DType.STRING:
            arr[i].data, arr[i].offsets = None, out.offsets[i].data_ptr()
        else:
            arr[i].data, arr[i].offsets = out.data[i].data_ptr(), None
    sizes = {}

    def alloc(_ctx, index, nbytes):
        cur = out.data[index]
        if cur is None or cur.numel() < nbytes:
            cur = _pinned(nbytes, torch.uint8, pin)
            out.data[index] = cur
        sizes[index] = int(nbytes)
        return cur.data_ptr()
    cb = N.HOST_ALLOC_FN(alloc)
    nulls = torch.zeros(max(nc, 1), dtype=torch.int64)
    N.check(N.lib().srj_convert_from_rows_host(plan.handle, h_rows.data_ptr(), h_offsets.data_ptr() if h_offsets is not None else None,
                                               h_rows.numel(), n, arr, nulls.data_ptr(), int(chunk_rows),
                                               C.cast(cb, C.c_void_p), None), "convertFromRows(host)")
    for i, d in enumerate(dts):
        if d.type_id == DType.STRING:
            out.data[i] = out.data[i][: sizes.get(i, 0)] if out.data[i] is not None else torch.empty(0, dtype=torch.uint8)
    out.null_counts = nulls[:nc]
    return out


def convert_to_rows_host(cols: Sequence[Tuple[Optional[torch.Tensor], Optional[torch.Tensor], Optional[torch.Tensor]]],
                         schema: Sequence, num_rows: int, pin: bool = True):
    """Host columns [(data, mask, offsets)] -> list of (offsets int32[rows + 1], row bytes) per <= 2 GiB batch."""
    dts = [_as_dtype(d) for d in schema]
    plan = Plan.get(dts)
    n, nc = int(num_rows), len(dts)
    arr = (N.SrjColumn * max(1, nc))()
    for i, (d, (data, mask, offs)) in enumerate(zip(dts, cols)):
        arr[i].type_id, arr[i].scale, arr[i].size = d.type_id, d.scale, n
        arr[i].data = data.data_ptr() if data is not None and data.numel() else None
        arr[i].null_mask = mask.data_ptr() if mask is not None else None
        arr[i].offsets = offs.data_ptr() if offs is not None else None
    bufs = {}

    def alloc(_ctx, index, nbytes):
        t = _pinned(max(int(nbytes), 1), torch.uint8, pin)
        bufs[index] = (t, int(nbytes))
        return t.data_ptr()
    cb = N.HOST_ALLOC_FN(alloc)
    maxb = 4096
    batches = (N.SrjRowBatch * maxb)()
    nb = C.c_int32(0)
    po, pd = (C.c_void_p * maxb)(), (C.c_void_p * maxb)()
    N.check(N.lib().srj_convert_to_rows_host(plan.handle, arr, n, batches, maxb, C.byref(nb), po, pd, C.cast(cb, C.c_void_p), None),
            "convertToRows(host)")
    out = []
    for b in range(nb.value):
        o, ob = bufs[2 * b]
        d, db = bufs[2 * b + 1]
        out.append((o[:ob].view(torch.int32), d[: batches[b].num_bytes]))
    if nb.value == 0:
        out.append((torch.zeros(1, dtype=torch.int32), torch.empty(0, dtype=torch.uint8)))
    return out
