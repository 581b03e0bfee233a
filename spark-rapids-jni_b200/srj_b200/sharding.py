"""Row-range sharding of a conversion across the GPUs of one node.

Rows are independent units, so the path shards by contiguous row range, one process per GPU, with no
data-path collective; shard boundaries are multiples of 32 rows so that validity-mask words are never
shared between ranks (the reference cuts its own batches on 32-row boundaries for the same reason,
RC:1498, 1515-1517).  The optional column all-gather (north_star's multi-GPU config) concatenates the
per-rank column chunks in rank order; because of the 32-row alignment the mask words concatenate too.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def row_range(num_rows: int, rank: int, world: int, align: int = 32) -> Tuple[int, int]:
    """Contiguous [r0, r1) of `rank`: equal shares rounded up to `align` rows, the last rank takes the rest."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    per = (num_rows + world - 1) // world
    per = (per + align - 1) // align * align
    r0 = min(num_rows, rank * per)
    r1 = min(num_rows, r0 + per)
    return r0, r1


def all_ranges(num_rows: int, world: int, align: int = 32) -> List[Tuple[int, int]]:
    return [row_range(num_rows, r, world, align) for r in range(world)]


def rows_byte_range(row_offsets: Sequence[int], r0: int, r1: int, fixed_row_size: int = 0) -> Tuple[int, int]:
    """Byte range of rows [r0, r1) in a JCUDF row buffer: LIST offsets for variable-width tables
    (offsets[r0] .. offsets[r1]), r * fixed_row_size for fixed-width ones."""
    if row_offsets is None:
        return r0 * fixed_row_size, r1 * fixed_row_size
    return int(row_offsets[r0]), int(row_offsets[r1])


def gather_fixed_columns(dist, chunks, world: int):
    """All-gather equal-sized per-rank column chunks (torch tensors) into full columns, rank order.
    `dist` is torch.distributed (NCCL on the GPUs, gloo in the CPU tests)."""
    import torch
    out = []
    for c in chunks:
        full = torch.empty((world,) + tuple(c.shape), dtype=c.dtype, device=c.device)
        dist.all_gather_into_tensor(full.view(-1), c.contiguous().view(-1))
        out.append(full.view(-1))
    return out
