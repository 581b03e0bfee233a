"""The Kudo shuffle wire format on the device: host-side mirror of the reference's KudoGpuSerializer
(kudo/KudoGpuSerializer.java: splitAndSerializeToDevice / assembleFromDeviceRaw over shuffle_split / shuffle_assemble)
on the C ABI (include/srj_b200.h: srj_kudo_split_sizes / srj_kudo_split / srj_kudo_assemble_sizes / srj_kudo_assemble).

    buf, offsets = KudoGpuSerializer.splitAndSerializeToDevice(table, splits)   # splits: row indices 0 .. n (P + 1 of them)
    table = KudoGpuSerializer.assembleFromDeviceRaw(schema, buf, offsets)
"""
import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _native as N
from . import ColumnVector, DType, Table, _as_dtype, _carray, _empty, _stream_ptr


class KudoGpuSerializer:
    @staticmethod
    def splitAndSerializeToDevice(table: Table, splits) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (uint8 device buffer with the P partitions back to back, int64 device offsets[P + 1])."""
        cols = table.columns
        n = table.getRowCount()
        dev = next((t.device for c in cols for t in (c.data, c.offsets, c.mask) if t is not None), torch.device("cuda", torch.cuda.current_device()))
        lib = N.lib()
        with torch.cuda.device(dev):
            st = _stream_ptr()
            d_splits = splits if isinstance(splits, torch.Tensor) else torch.tensor(list(splits), dtype=torch.int32, device=dev)
            d_splits = d_splits.to(torch.int32)
            P = d_splits.numel() - 1
            ws = _empty(lib.srj_kudo_workspace_bytes(len(cols), P), torch.uint8, dev)
            offs = _empty(P + 1, torch.int64, dev)
            total = C.c_int64(0)
            carr = _carray(cols)
            N.check(lib.srj_kudo_split_sizes(carr, len(cols), n, d_splits.data_ptr(), P, offs.data_ptr(), C.byref(total), ws.data_ptr(), st),
                    "kudo split")
            buf = _empty(total.value, torch.uint8, dev)
            N.check(lib.srj_kudo_split(carr, len(cols), n, d_splits.data_ptr(), P, offs.data_ptr(), buf.data_ptr(), ws.data_ptr(), st), "kudo split")
        return buf, offs

    @staticmethod
    def assembleFromDeviceRaw(schema: Sequence, buf: torch.Tensor, offsets: torch.Tensor) -> Table:
        dts = [_as_dtype(d) for d in schema]
        dev = buf.device
        lib = N.lib()
        P = offsets.numel() - 1
        with torch.cuda.device(dev):
            st = _stream_ptr()
            ws = _empty(lib.srj_kudo_workspace_bytes(len(dts), P), torch.uint8, dev)
            ids = (C.c_int32 * len(dts))(*[d.type_id for d in dts])
            rows = C.c_int64(0)
            chars = (C.c_int64 * len(dts))()
            N.check(lib.srj_kudo_assemble_sizes(buf.data_ptr(), offsets.data_ptr(), P, ids, len(dts), C.byref(rows), chars, ws.data_ptr(), st),
                    "kudo assemble")
            n = rows.value
            words = (n + 31) // 32
            outs: List[ColumnVector] = []
            for i, d in enumerate(dts):
                mask = _empty(max(1, words), torch.int32, dev)
                if d.type_id == DType.STRING:
                    outs.append(ColumnVector(d, n, _empty(int(chars[i]), torch.uint8, dev), mask, _empty(n + 1, torch.int32, dev)))
                else:
                    outs.append(ColumnVector(d, n, _empty(n * d.size_in_bytes(), torch.uint8, dev), mask))
            N.check(lib.srj_kudo_assemble(buf.data_ptr(), offsets.data_ptr(), P, _carray(outs), len(dts), n, ws.data_ptr(), st), "kudo assemble")
        return Table(outs)
