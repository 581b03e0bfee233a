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


# ---------------------------------------------------------------------------------------------------
# Packed slab: everything one rank contributes to the column all-gather, in ONE contiguous buffer, so that the
# collective is a single ncclAllGather per round (north_star's multi-GPU configuration).
# ---------------------------------------------------------------------------------------------------
class SlabLayout:
    """Byte layout of a rank's packed slab for `rows` rows of a schema:

        [per column: fixed-width data (rows * size)  |  STRING: int32 offsets[rows + 1]]   256-byte aligned pieces
        [per column: validity mask words]
        [phase-1 totals: int64[ncols + 1]]
        [chars of the STRING columns, back to back at 16-byte alignment, padded to the common capacity]

    The shard height is a multiple of 32 rows, so the mask words of consecutive ranks concatenate.  After the
    all-gather the chunks of rank r sit at gathered[r * nbytes + at_*]; STRING offsets of rank r > 0 are local to its
    shard until srj_shard_rebase_offsets (CUDA) / rebase_offsets_host (numpy, tests) adds the chars of the ranks before.
    """

    def __init__(self, elem_sizes: Sequence[int], rows: int, chars_capacity: int, align: int = 256):
        """elem_sizes[c] = bytes per element, 0 for a STRING column."""
        self.rows = int(rows)
        self.ncols = len(elem_sizes)
        self.words = (self.rows + 31) // 32
        self.string_cols = [c for c, s in enumerate(elem_sizes) if s == 0]
        size = 0

        def add(n):
            nonlocal size
            off = size
            size = (off + int(n) + align - 1) // align * align
            return off
        self.at_data = [add((self.rows + 1) * 4 if s == 0 else self.rows * s) for s in elem_sizes]
        self.data_bytes = [(self.rows + 1) * 4 if s == 0 else self.rows * s for s in elem_sizes]
        self.at_mask = [add(self.words * 4) for _ in elem_sizes]
        self.at_totals = add((self.ncols + 1) * 8)
        self.chars_capacity = (int(chars_capacity) + align - 1) // align * align
        self.at_chars = add(self.chars_capacity)
        self.nbytes = size

    def chars_offsets(self, chars_sizes: Sequence[int]) -> List[int]:
        """Slab byte offset of each STRING column's chars given their sizes (16-byte aligned, back to back)."""
        out, at = [], self.at_chars
        for n in chars_sizes:
            out.append(at)
            at += (int(n) + 15) & ~15
        if at - self.at_chars > self.chars_capacity:
            raise ValueError("chars exceed the slab's capacity")
        return out


def rebase_offsets_host(gathered, layout: SlabLayout, world: int):
    """numpy twin of srj_shard_rebase_offsets: gathered = uint8 array [world * layout.nbytes], modified in place."""
    import numpy as np
    g = gathered.reshape(world, layout.nbytes)
    totals = np.stack([g[r, layout.at_totals: layout.at_totals + (layout.ncols + 1) * 8].view(np.int64) for r in range(world)])
    for c in layout.string_cols:
        delta = 0
        for r in range(world):
            if r:
                o = g[r, layout.at_data[c]: layout.at_data[c] + (layout.rows + 1) * 4].view(np.int32)
                o += np.int32(delta)
            delta += int(totals[r, c])
    return totals


def gather_slab(dist, slab, world: int, out=None, async_op: bool = False):
    """ONE all-gather of the rank's packed slab (torch uint8 tensor) -> [world * nbytes]."""
    import torch
    if out is None:
        out = torch.empty(world * slab.numel(), dtype=slab.dtype, device=slab.device)
    work = dist.all_gather_into_tensor(out, slab, async_op=async_op)
    return out, work


def convert_from_rows_into_slab(vec, dtypes, layout: SlabLayout, slab):
    """This rank's shard: JCUDF rows (LIST<INT8> ColumnView `vec`, layout.rows rows) -> columns written straight into the
    packed slab (torch uint8 CUDA tensor of layout.nbytes) through the C ABI -- phase 1, the size read-back, phase 2.
    Returns the output ColumnVectors (views into the slab)."""
    import ctypes as C

    import torch

    from . import ColumnVector, DType, Plan, _native as N
    n, nc = layout.rows, layout.ncols
    assert vec.size == n and slab.numel() >= layout.nbytes
    plan = Plan.get(dtypes)
    lib = N.lib()
    st = int(torch.cuda.current_stream().cuda_stream)
    outs = []
    for i, d in enumerate(dtypes):
        m = slab[layout.at_mask[i]: layout.at_mask[i] + layout.words * 4].view(torch.int32)
        if d.type_id == DType.STRING:
            outs.append(ColumnVector(d, n, None, m, slab[layout.at_data[i]: layout.at_data[i] + (n + 1) * 4].view(torch.int32)))
        else:
            outs.append(ColumnVector(d, n, slab[layout.at_data[i]: layout.at_data[i] + n * d.size_in_bytes()], m))
    carr = (N.SrjColumn * nc)()
    for i, c in enumerate(outs):
        carr[i] = c._c()
    totals = slab[layout.at_totals: layout.at_totals + (nc + 1) * 8].view(torch.int64)
    nulls = torch.zeros(nc, dtype=torch.int64, device=slab.device)
    wsb = lib.srj_from_rows_workspace_bytes(plan.handle, n)
    ws = torch.empty(max(wsb, 8), dtype=torch.uint8, device=slab.device)
    child = vec.child
    N.check(lib.srj_convert_from_rows_fixed(plan.handle, child.data.data_ptr(), vec.offsets.data_ptr(), child.size, n, carr,
                                            nulls.data_ptr(), totals.data_ptr(), None, ws.data_ptr(), st), "convertFromRows")
    if layout.string_cols:
        h_tot = totals.cpu().numpy()
        at = layout.chars_offsets([int(h_tot[c]) for c in layout.string_cols])
        for a, c in zip(at, layout.string_cols):
            outs[c].data = slab[a: a + int(h_tot[c])]
            carr[c] = outs[c]._c()
        N.check(lib.srj_convert_from_rows_strings(plan.handle, child.data.data_ptr(), vec.offsets.data_ptr(), child.size, n, carr,
                                                  totals.data_ptr(), ws.data_ptr(), st), "convertFromRows")
    h_nulls = nulls.cpu().numpy()
    for i, o in enumerate(outs):
        o._null_count = int(h_nulls[i])
    return outs


def rebase_gathered_offsets(gathered, layout: SlabLayout, world: int):
    """srj_shard_rebase_offsets on the all-gathered slabs (torch uint8 CUDA tensor [world * layout.nbytes]), in place."""
    import torch

    from . import _native as N
    dev = gathered.device
    d_at = torch.tensor([layout.at_data[c] for c in layout.string_cols], dtype=torch.int64, device=dev)
    d_sc = torch.tensor(layout.string_cols, dtype=torch.int32, device=dev)
    gt = gathered.view(world, layout.nbytes)[:, layout.at_totals: layout.at_totals + (layout.ncols + 1) * 8].contiguous().view(torch.int64)
    N.check(N.lib().srj_shard_rebase_offsets(gathered.data_ptr(), layout.nbytes, d_at.data_ptr(), d_sc.data_ptr(), gt.data_ptr(),
                                             layout.rows, layout.ncols, len(layout.string_cols), world,
                                             int(torch.cuda.current_stream().cuda_stream)), "shard_rebase_offsets")
    torch.cuda.current_stream().synchronize()     # d_at / d_sc / gt must outlive the kernel
    return gt


class PeerGather:
    """All-gather of the packed slabs over NVLink / NVSwitch PEER MEMORY with the copy engines: every rank copies its
    slab straight into slot `rank` of every peer's gathered buffer (symmetric memory), so the collective takes no SM and
    no shared memory away from the conversion kernels it overlaps with (those are persistent, one CTA per SM with
    ~all of its shared memory: an SM-based collective kernel can only run in the gaps between them).

    gather(slab, buf_index) enqueues, on the side streams, for round-robin buffer `buf_index`:
      barrier (every peer is done with this buffer's previous contents) -> N copies -> barrier (all copies landed).
    `done` (a CUDA event) then marks the gathered buffer complete.
    """

    def __init__(self, dist, slab_bytes: int, world: int, rank: int, device, nbuf: int = 2):
        import torch
        import torch.distributed._symmetric_memory as symm_mem
        self.torch, self.world, self.rank, self.slab_bytes = torch, world, rank, int(slab_bytes)
        self.bufs, self.hdls, self.peers = [], [], []
        group = dist.group.WORLD
        for _ in range(nbuf):
            t = symm_mem.empty(world * self.slab_bytes, dtype=torch.uint8, device=device)
            h = symm_mem.rendezvous(t, group.group_name)
            self.bufs.append(t)
            self.hdls.append(h)
            self.peers.append([h.get_buffer(p, (world * self.slab_bytes,), torch.uint8) for p in range(world)])
        self.main = torch.cuda.Stream()
        self.side = [torch.cuda.Stream() for _ in range(world)]

    def gather(self, slab, b: int, after_event):
        """Enqueue the all-gather of `slab` into buffer b once `after_event` (the conversion) has completed.
        Returns an event recorded when every rank's slab has landed in this rank's buffer b."""
        torch = self.torch
        self.main.wait_event(after_event)
        with torch.cuda.stream(self.main):
            self.hdls[b].barrier(channel=0)              # peers have consumed the buffer's previous contents
            start = torch.cuda.Event()
            start.record(self.main)
        lo, hi = self.rank * self.slab_bytes, (self.rank + 1) * self.slab_bytes
        for p in range(self.world):                      # one stream per destination: the copy engines run in parallel
            q = (self.rank + p) % self.world             # staggered so that no destination is hit by everyone at once
            s = self.side[p]
            s.wait_event(start)
            with torch.cuda.stream(s):
                self.peers[b][q][lo:hi].copy_(slab, non_blocking=True)
            self.main.wait_stream(s)
        with torch.cuda.stream(self.main):
            self.hdls[b].barrier(channel=1)              # every rank's copies into every buffer b have completed
            done = torch.cuda.Event()
            done.record(self.main)
        return done
