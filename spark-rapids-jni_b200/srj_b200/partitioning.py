"""Spark HashPartitioning on the device: the host-side mirror of the plugin's GpuHashPartitioning step
(partition id = pmod(murmur3_32(42, keys), numPartitions), then ai.rapids.cudf.Table.partition) over the C ABI
(include/srj_b200.h: srj_hash_partition / srj_partition_plan / srj_partition_columns / srj_partition_strings).

    pt = HashPartitioner.partition(table, key_columns=[0, 3], num_partitions=200)     # seed 42
    pt.getTable()          -> Table whose rows are grouped by partition, input order kept inside a partition
    pt.getPartitions()     -> int[numPartitions]: first row of every partition (cudf PartitionedTable.getPartitions)
"""
import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _native as N
from . import ColumnVector, ColumnView, DType, Table, _carray, _empty, _stream_ptr


class PartitionedTable:
    """ai.rapids.cudf.PartitionedTable: the partitioned table + where each partition starts."""

    def __init__(self, table: Table, offsets: np.ndarray, partition_ids: Optional[ColumnVector] = None):
        self._table = table
        self._offsets = offsets                  # int32[P + 1]
        self.partition_ids = partition_ids       # INT32 column of the input rows' partition ids (hash partitioning)

    def getTable(self) -> Table:
        return self._table

    def getPartitions(self) -> List[int]:
        return [int(x) for x in self._offsets[:-1]]

    def getRowCounts(self) -> List[int]:
        return [int(x) for x in np.diff(self._offsets)]

    def close(self):
        self._table.close()


def _partition(table: Table, ids: torch.Tensor, num_partitions: int, hash_keys: Optional[Sequence[ColumnView]], seed: int):
    lib = N.lib()
    cols = table.columns
    n = table.getRowCount()
    dev = ids.device
    with torch.cuda.device(dev):
        st = _stream_ptr()
        ws = _empty(max(8, lib.srj_partition_workspace_bytes(n, num_partitions)), torch.uint8, dev)
        offsets = _empty(num_partitions + 1, torch.int32, dev)
        smap = _empty(max(1, n), torch.int32, dev)
        gmap = _empty(max(1, n), torch.int32, dev)
        if hash_keys is not None:
            N.check(lib.srj_hash_partition(_carray(hash_keys), len(hash_keys), n, C.c_uint32(seed & 0xFFFFFFFF), num_partitions,
                                           ids.data_ptr(), offsets.data_ptr(), smap.data_ptr(), gmap.data_ptr(), ws.data_ptr(), st),
                    "hashPartition")
        else:
            N.check(lib.srj_partition_plan(ids.data_ptr(), n, num_partitions, offsets.data_ptr(), smap.data_ptr(), gmap.data_ptr(),
                                           ws.data_ptr(), st), "partition")
        outs: List[ColumnVector] = []
        words = (n + 31) // 32
        for c in cols:
            mask = _empty(max(1, words), torch.int32, dev) if c.mask is not None else None
            if c.dtype.type_id == DType.STRING:
                outs.append(ColumnVector(c.dtype, n, None, mask, _empty(n + 1, torch.int32, dev)))
            elif c.dtype.is_fixed_width():
                outs.append(ColumnVector(c.dtype, n, _empty(n * c.dtype.size_in_bytes(), torch.uint8, dev), mask))
            else:
                raise ValueError(f"partition: unsupported column type {c.dtype}")
        nulls = torch.zeros(max(1, len(cols)), dtype=torch.int64, device=dev)
        cin, cout = _carray(cols), _carray(outs)
        N.check(lib.srj_partition_columns(cin, cout, len(cols), n, num_partitions, smap.data_ptr(), gmap.data_ptr(), nulls.data_ptr(), ws.data_ptr(), st),
                "partition")
        sidx = [i for i, c in enumerate(cols) if c.dtype.type_id == DType.STRING]
        if sidx:
            # the one host read of the string path: how many chars every output column needs (its offsets[n])
            totals = torch.stack([outs[i].offsets[n] for i in sidx]).cpu().numpy() if n else np.zeros(len(sidx), np.int64)
            for i, t in zip(sidx, totals):
                outs[i].data = _empty(int(t), torch.uint8, dev)
            cout = _carray(outs)
            N.check(lib.srj_partition_strings(cin, cout, len(cols), n, gmap.data_ptr(), st), "partition")
        h_nulls = nulls.cpu().numpy()
        for i, o in enumerate(outs):
            o._null_count = int(h_nulls[i]) if o.mask is not None else 0
        h_off = offsets.cpu().numpy()
    return Table(outs), h_off


class HashPartitioner:
    """GpuHashPartitioning: Spark's murmur3 (seed 42) over the key columns, pmod numPartitions, stable partition."""
    DEFAULT_SEED = 42

    @staticmethod
    def partitionIds(keys: Sequence[ColumnView], num_partitions: int, seed: int = 42) -> ColumnVector:
        """pmod(murmur3_32(seed, keys), numPartitions) as an INT32 column (no data movement)."""
        from . import Hash
        h = Hash.murmurHash32(seed, list(keys))
        ids = h.data.view(torch.int32)
        n = ids.numel()
        lib = N.lib()
        with torch.cuda.device(ids.device):
            ws = _empty(max(8, lib.srj_partition_workspace_bytes(n, num_partitions)), torch.uint8, ids.device)
            offsets = _empty(num_partitions + 1, torch.int32, ids.device)
            N.check(lib.srj_partition_plan(ids.data_ptr(), n, num_partitions, offsets.data_ptr(), None, None, ws.data_ptr(), _stream_ptr()),
                    "partitionIds")
        return h

    @staticmethod
    def partition(table: Table, key_columns: Sequence[int], num_partitions: int, seed: int = 42) -> PartitionedTable:
        if num_partitions <= 0:
            raise ValueError("numPartitions must be positive")
        keys = [table.columns[i] for i in key_columns]
        if not keys:
            raise ValueError("hash partitioning needs at least one key column")
        n = table.getRowCount()
        dev = next(t.device for c in table.columns for t in (c.data, c.offsets, c.mask) if t is not None)
        with torch.cuda.device(dev):
            ids = _empty(max(1, n), torch.int32, dev)
        out, h_off = _partition(table, ids, num_partitions, keys, seed)
        return PartitionedTable(out, h_off, ColumnVector(DType.INT32, n, ids[:n].view(torch.uint8), None, null_count=0))


def partition(table: Table, partition_map: ColumnView, num_partitions: int) -> PartitionedTable:
    """ai.rapids.cudf.Table.partition(partitionMap, numberOfPartitions): partition_map = INT32 ids in [0, P)."""
    if partition_map.dtype.type_id != DType.INT32 or partition_map.size != table.getRowCount():
        raise ValueError("partitionMap must be an INT32 column with one id per row")
    ids = partition_map.data.view(torch.int32).clone()       # the plan reduces ids in place
    out, h_off = _partition(table, ids, num_partitions, None, 0)
    return PartitionedTable(out, h_off)
