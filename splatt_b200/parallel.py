"""Multi-GPU plumbing: one process per GPU (torch.distributed), the tensor's fiber
streams sharded by equal-nnz contiguous ranges, one all-reduce(sum) of the output
factor per mode -- the exchange step the north star names.  NCCL on GPUs; the same
code runs over gloo on CPU tensors for the host-logic tests."""
from __future__ import annotations

import ctypes as C

from . import _abi as A


def shard_range(nnz: int, rank: int, world: int):
    """(first, count): the records of a sorted stream that shard `rank` keeps."""
    lib = A.load()
    first = A.idx_t()
    count = A.idx_t()
    lib.splatt_b200_shard_range(nnz, rank, world, C.byref(first), C.byref(count))
    return int(first.value), int(count.value)


def all_reduce_output(out, group=None):
    """Sum the per-rank partial MTTKRP outputs in place (no-op outside a process group)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def sharded_mttkrp(tensor, mode, mats, out, group=None):
    """One distributed MTTKRP: local partial on this rank's shard, then the all-reduce."""
    tensor.mttkrp(mode, mats, out)
    return all_reduce_output(out, group)
