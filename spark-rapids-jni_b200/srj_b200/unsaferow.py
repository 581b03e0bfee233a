"""Columns <-> Apache Spark UnsafeRow on the device: the host-side mirror over the C ABI (include/srj_b200.h:
srj_unsafe_row_sizes / srj_convert_to_unsafe_rows / srj_convert_from_unsafe_rows[_strings]).  Shaped like
RowConversion (RowConversion.java:120-174): the rows travel as one LIST<INT8> column (offsets + bytes).

    rows = UnsafeRowConversion.convertToRows(table)            # ColumnVector LIST<INT8>
    table = UnsafeRowConversion.convertFromRows(rows, schema)  # schema: DTypes
"""
import ctypes as C
from typing import List, Sequence

import numpy as np
import torch

from . import _native as N
from . import ColumnVector, ColumnView, CudfColumnSizeOverflowException, DType, Table, _as_dtype, _carray, _empty, _stream_ptr


class UnsafeRowConversion:
    @staticmethod
    def layout(schema: Sequence) -> tuple:
        """(bitset bytes, bytes of a row without its strings)."""
        ids = (C.c_int32 * len(schema))(*[_as_dtype(d).type_id for d in schema])
        a, b = C.c_int32(0), C.c_int32(0)
        N.check(N.lib().srj_unsafe_row_layout(ids, len(schema), C.byref(a), C.byref(b)), "unsafeRowLayout")
        return a.value, b.value

    @staticmethod
    def convertToRows(table: Table) -> ColumnVector:
        cols = table.columns
        n = table.getRowCount()
        if not cols:
            raise ValueError("convertToRows needs at least one column")
        dev = next((t.device for c in cols for t in (c.data, c.offsets, c.mask) if t is not None), torch.device("cuda", torch.cuda.current_device()))
        lib = N.lib()
        with torch.cuda.device(dev):
            st = _stream_ptr()
            ws = _empty(lib.srj_unsafe_row_workspace_bytes(len(cols), n), torch.uint8, dev)
            offs = _empty(n + 1, torch.int32, dev)
            total = C.c_int64(0)
            carr = _carray(cols)
            N.check(lib.srj_unsafe_row_sizes(carr, len(cols), n, offs.data_ptr(), C.byref(total), ws.data_ptr(), st), "unsafe convertToRows")
            data = _empty(total.value, torch.uint8, dev)
            N.check(lib.srj_convert_to_unsafe_rows(carr, len(cols), n, offs.data_ptr(), data.data_ptr(), ws.data_ptr(), st),
                    "unsafe convertToRows")
            child = ColumnVector(DType.INT8, total.value, data, None, null_count=0)
            return ColumnVector(DType.LIST, n, None, None, offs, child, null_count=0)

    @staticmethod
    def convertFromRows(vec: ColumnView, *schema) -> Table:
        if len(schema) == 1 and isinstance(schema[0], (list, tuple)):
            schema = tuple(schema[0])
        dts = [_as_dtype(d) for d in schema]
        n = vec.size
        rows, offs = vec.child.data, vec.offsets
        dev = offs.device
        lib = N.lib()
        with torch.cuda.device(dev):
            st = _stream_ptr()
            ws = _empty(lib.srj_unsafe_row_workspace_bytes(len(dts), n), torch.uint8, dev)
            words = (n + 31) // 32
            outs: List[ColumnVector] = []
            for d in dts:
                mask = _empty(max(1, words), torch.int32, dev)
                if d.type_id == DType.STRING:
                    outs.append(ColumnVector(d, n, None, mask, _empty(n + 1, torch.int32, dev)))
                else:
                    outs.append(ColumnVector(d, n, _empty(n * d.size_in_bytes(), torch.uint8, dev), mask))
            nulls = torch.zeros(len(dts), dtype=torch.int64, device=dev)
            carr = _carray(outs)
            N.check(lib.srj_convert_from_unsafe_rows(rows.data_ptr(), offs.data_ptr(), n, carr, len(dts), nulls.data_ptr(), ws.data_ptr(), st),
                    "unsafe convertFromRows")
            sidx = [i for i, d in enumerate(dts) if d.type_id == DType.STRING]
            if sidx:
                totals = torch.stack([outs[i].offsets[n] for i in sidx]).cpu().numpy()      # the one read-back: chars per column
                for i, t in zip(sidx, totals):
                    outs[i].data = _empty(int(t), torch.uint8, dev)
                carr = _carray(outs)
                N.check(lib.srj_convert_from_unsafe_rows_strings(rows.data_ptr(), offs.data_ptr(), n, carr, len(dts), st),
                        "unsafe convertFromRows")
            h_nulls = nulls.cpu().numpy()
            for i, o in enumerate(outs):
                o._null_count = int(h_nulls[i])
                if o._null_count == 0:
                    o.mask = None                         # like the row path: no nulls, no mask
        return Table(outs)
