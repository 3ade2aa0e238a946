"""Batch sharding across the GPUs of one box.

Every signal / image of a batch is an independent transform, so multi-GPU execution is a pure batch
split: rank r owns a contiguous block of slices along the last (batch) dimension and runs the single-GPU
path on it.  There is no data-path collective; NCCL is used only for the setup broadcast of the tiny
wavelet descriptor (when rank 0 owns it), an optional gather of results and the scalar all-reduce that
closes a timed region (bench.py).
"""
from __future__ import annotations

from typing import Tuple

__all__ = ["shard_range", "shard_sizes", "broadcast_descriptor", "allreduce_max", "allreduce_sum", "scatter_columns", "gather_columns"]


def shard_sizes(batch: int, world: int):
    """Contiguous blocks of batch//world slices, the remainder going to the low ranks."""
    q, r = divmod(int(batch), int(world))
    return [q + (1 if i < r else 0) for i in range(world)]


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    sizes = shard_sizes(batch, world)
    lo = sum(sizes[:rank])
    return lo, lo + sizes[rank]


def broadcast_descriptor(obj, src: int = 0):
    """Broadcast a small picklable wavelet descriptor (qmf / step table) from `src` to every rank."""
    import torch.distributed as dist
    box = [obj]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def _allreduce(value: float, op, device):
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=op)
    return float(t.item())


def allreduce_max(value: float, device="cpu") -> float:
    import torch.distributed as dist
    return _allreduce(value, dist.ReduceOp.MAX, device)


def allreduce_sum(value: float, device="cpu") -> float:
    import torch.distributed as dist
    return _allreduce(value, dist.ReduceOp.SUM, device)


def scatter_columns(full, batch: int, tail_shape, dtype, device, src: int = 0):
    """The split itself when the batch starts on ONE rank (optional row): `src` holds `full` with the batch as its
    LAST (slowest-varying, Julia layout) dimension and sends every other rank its contiguous block of slices as one
    point-to-point message (NCCL send/recv = a peer copy over NVLink between the GPUs of one box); returns this rank's
    shard, shape tail_shape + (b_r,), row-major over the reversed dims so that `.permute` gives the column-major view.
    `full` is laid out (batch, *reversed tail dims) contiguous -- i.e. the storage of a column-major (tail..., batch) array."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    sizes = shard_sizes(batch, world)
    rev = tuple(reversed(tuple(int(v) for v in tail_shape)))
    if rank == src:
        assert full.shape == (batch,) + rev and full.is_contiguous()
        lo = 0
        mine = None
        for r in range(world):
            part = full[lo:lo + sizes[r]]
            if r == src:
                mine = part
            elif sizes[r]:
                dist.send(part, dst=r)
            lo += sizes[r]
        return mine
    buf = torch.empty((sizes[rank],) + rev, dtype=dtype, device=device)
    if sizes[rank]:
        dist.recv(buf, src=src)
    return buf


def gather_columns(local, batch: int, dst: int = 0):
    """Gather each rank's (n..., b_r) shard on `dst` and concatenate along the batch dimension (the inverse
    of the split).  Shards may have different sizes; they travel as separate point-to-point messages."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    sizes = shard_sizes(batch, world)
    if rank == dst:
        parts = []
        for r in range(world):
            if r == dst:
                parts.append(local.contiguous())
            else:
                shape = list(local.shape[:-1]) + [sizes[r]]
                buf = torch.empty(shape, dtype=local.dtype, device=local.device)
                if sizes[r]:
                    dist.recv(buf, src=r)
                parts.append(buf)
        return torch.cat(parts, dim=-1)
    if local.shape[-1]:
        dist.send(local.contiguous(), dst=dst)
    return None
