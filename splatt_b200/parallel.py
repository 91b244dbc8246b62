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


class FusedExchange:
    """MTTKRP with the exchange fused into the kernel (NVLink multicast).

    One symmetric-memory buffer per mode (dims[mode] x ldm fp64) is mapped on every GPU
    and bound to a multicast address; `mttkrp(mode, mats)` launches the root kernel with
    `multimem.red.add.f64` reductions into that address, then one group barrier.  Buffers
    are re-zeroed locally right after they are consumed; with one buffer per mode the
    barrier that ends the next mode's call orders that zeroing before any peer's next
    write, so a single barrier per call suffices (needs >= 2 modes, always true).

    `available()` is False when the process group has no multicast support (then use
    `sharded_mttkrp`: local kernel + NCCL all-reduce)."""

    def __init__(self, tensor, ncolumns, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.t = tensor
        self.R = ncolumns
        self.ldm = ncolumns + (ncolumns & 1)
        self.group = group if group is not None else dist.group.WORLD
        dev = torch.device("cuda", torch.cuda.current_device())
        self.bufs, self.hdls = [], []
        self.ok = True
        self.error = None
        try:
            for m in range(tensor.nmodes):
                b = symm_mem.empty((tensor.dims[m], self.ldm), dtype=torch.float64, device=dev)
                h = symm_mem.rendezvous(b, self.group)
                if not h.multicast_ptr:
                    self.ok = False
                    self.error = "symmetric memory works but the group has no multicast support"
                self.bufs.append(b)
                self.hdls.append(h)
        except Exception as e:      # no symmetric memory in this environment
            self.ok = False
            self.error = f"{type(e).__name__}: {e}"
        if self.ok:
            for b in self.bufs:
                b.zero_()
            torch.cuda.synchronize()
            self.hdls[0].barrier(channel=0)
            torch.cuda.synchronize()

    def available(self) -> bool:
        return self.ok

    def mttkrp(self, mode, mats):
        """Returns the buffer holding the full (summed over ranks) MTTKRP of `mode`."""
        import torch
        lib = self.t.lib
        ptrs = (A.val_p * self.t.nmodes)()
        for m in range(self.t.nmodes):
            if m == mode or mats[m] is None:
                ptrs[m] = A.val_p()
            else:
                ptrs[m] = C.cast(C.c_void_p(mats[m].data_ptr()), A.val_p)
        s = torch.cuda.current_stream().cuda_stream
        rc = lib.splatt_b200_mttkrp_multicast(
            self.t.h, mode, self.R, self.ldm, ptrs,
            C.cast(C.c_void_p(self.hdls[mode].multicast_ptr), A.val_p), C.c_void_p(s))
        if rc != A.SPLATT_SUCCESS:
            raise RuntimeError(f"splatt_b200_mttkrp_multicast failed ({rc})")
        self.hdls[mode].barrier(channel=0)
        return self.bufs[mode]

    def release(self, mode):
        """Call when the result of `mode` has been consumed: re-zero it for its next use."""
        self.bufs[mode].zero_()
