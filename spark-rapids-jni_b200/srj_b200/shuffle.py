"""The shuffle exchange across GPUs: hash partition -> Kudo split -> one all-to-all of the partitions -> assemble.

This is the step of the widened path (SURVEY 8f ranks 1 + 2) that really exchanges data between ranks -- the reference
leaves it to Spark's shuffle (UCX / host files in the plugin); here the partitions travel GPU to GPU over NVLink with
`torch.distributed.all_to_all_single` (NCCL), one process per GPU.  Rank r ends up with the rows whose partition id is in
[r * parts_per_rank, (r + 1) * parts_per_rank), assembled into one table, rows ordered by (source rank, partition, input
order).

    out = ShuffleExchange().shuffle(table, key_columns=[0, 1], parts_per_rank=4)      # torchrun, backend nccl

`exchange_partitions` is the transport alone (sizes, then bytes, then the partition offsets of what arrived): it works on
CPU tensors over gloo as well, which is how the N > 1 logic is tested without GPUs.
"""
from typing import Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def exchange_partitions(buf: torch.Tensor, part_offsets: torch.Tensor, parts_per_rank: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """buf: uint8 partitions back to back (world * parts_per_rank of them), part_offsets: int64[P + 1].
    -> (received buffer: for every source rank in order, the parts_per_rank partitions addressed to this rank;
        int64 offsets[world * parts_per_rank + 1] of those partitions in the received buffer)."""
    world = dist.get_world_size(group)
    k = parts_per_rank
    assert part_offsets.numel() == world * k + 1, "one partition range per rank"
    sizes = (part_offsets[1:] - part_offsets[:-1]).reshape(world, k).contiguous()           # [dest rank][partition]
    got_sizes = torch.empty_like(sizes)                                                      # [source rank][partition]
    dist.all_to_all_single(got_sizes, sizes, group=group)
    send_split = sizes.sum(dim=1).tolist()                                                   # one host read-back
    recv_split = got_sizes.sum(dim=1).tolist()
    recv = torch.empty(int(sum(recv_split)), dtype=torch.uint8, device=buf.device)
    dist.all_to_all_single(recv, buf, output_split_sizes=recv_split, input_split_sizes=send_split, group=group)
    recv_offsets = torch.zeros(world * k + 1, dtype=torch.int64, device=buf.device)
    torch.cumsum(got_sizes.reshape(-1), 0, out=recv_offsets[1:])
    return recv, recv_offsets


class ShuffleExchange:
    def __init__(self, group=None):
        self.group = group

    def shuffle(self, table, key_columns: Sequence[int], parts_per_rank: int = 1, seed: int = 42):
        """Spark HashPartitioning with world * parts_per_rank partitions, then the exchange; returns this rank's table."""
        from .kudo import KudoGpuSerializer
        from .partitioning import HashPartitioner
        world = dist.get_world_size(self.group)
        P = world * parts_per_rank
        pt = HashPartitioner.partition(table, key_columns, P, seed)
        splits = pt.getPartitions() + [table.getRowCount()]
        buf, offs = KudoGpuSerializer.splitAndSerializeToDevice(pt.getTable(), splits)
        recv, recv_offs = exchange_partitions(buf, offs, parts_per_rank, self.group)
        return KudoGpuSerializer.assembleFromDeviceRaw([c.dtype for c in table.columns], recv, recv_offs)
