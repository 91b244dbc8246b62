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
    `multimem.red.add.f64` reductions into that address, then one group barrier.  A buffer
    is re-zeroed locally by `release(mode)` once its result is consumed.  A peer may only
    write into it again after that zeroing: some group barrier must lie between the
    release and the buffer's next use.  In the ALS order (release right after each mode)
    the barrier ending the next mode's call provides it for free -- one barrier per call;
    otherwise `mttkrp` inserts the missing barrier itself (every rank takes the same
    decision, so the extra barrier is collective).

    The group barrier after the kernel is by default the kernel's own tail
    (`splatt_b200_mttkrp_multicast_sync`: every CTA fences, the last one adds 1 to a flag on
    every GPU through the multicast address and spins on the local copy) -- no separate
    barrier launch, no launch-skew between the kernel and its barrier.  `kernel_barrier=False`
    (or SPLATT_B200_KERNEL_BARRIER=0) uses the symmetric-memory barrier kernel instead.

    `available()` is False when the process group has no multicast support (then use
    `sharded_mttkrp`: local kernel + NCCL all-reduce)."""

    def __init__(self, tensor, ncolumns, group=None, kernel_barrier=None):
        import os
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.t = tensor
        self.R = ncolumns
        self.ldm = ncolumns + (ncolumns & 1)
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if kernel_barrier is None:
            kernel_barrier = os.environ.get("SPLATT_B200_KERNEL_BARRIER", "1") != "0"
        self.kernel_barrier = bool(kernel_barrier)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.bufs, self.hdls = [], []
        self.ok = True
        self.error = None
        self.flags = None
        try:
            for m in range(tensor.nmodes):
                b = symm_mem.empty((tensor.dims[m], self.ldm), dtype=torch.float64, device=dev)
                h = symm_mem.rendezvous(b, self.group)
                if not h.multicast_ptr:
                    self.ok = False
                    self.error = "symmetric memory works but the group has no multicast support"
                self.bufs.append(b)
                self.hdls.append(h)
            self.flags = symm_mem.empty((64,), dtype=torch.int32, device=dev)
            self.flags_hdl = symm_mem.rendezvous(self.flags, self.group)
            if not self.flags_hdl.multicast_ptr:
                self.ok = False
                self.error = "symmetric memory works but the group has no multicast support"
        except Exception as e:      # no symmetric memory in this environment
            self.ok = False
            self.error = f"{type(e).__name__}: {e}"
        self._barriers = 0                       # group barriers issued so far
        self._epoch = 0                          # in-kernel barriers issued so far
        self._released_at = [-1] * tensor.nmodes  # barrier count when the buffer was last zeroed
        if self.ok:
            for b in self.bufs:
                b.zero_()
            self.flags.zero_()
            torch.cuda.synchronize()
            self._barrier(0)
            torch.cuda.synchronize()

    def _barrier(self, mode):
        self.hdls[mode].barrier(channel=0)
        self._barriers += 1

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
        if self._released_at[mode] == self._barriers:
            self._barrier(mode)      # no barrier since this buffer was zeroed: order it now
        s = torch.cuda.current_stream().cuda_stream
        mc_out = C.cast(C.c_void_p(self.hdls[mode].multicast_ptr), A.val_p)
        if self.kernel_barrier:
            self._epoch += 1
            gs = A.GroupSync(self.flags_hdl.multicast_ptr, self.flags.data_ptr(),
                             self._epoch & 0xffffffff, self.rank, self.world, 0)
            rc = lib.splatt_b200_mttkrp_multicast_sync(self.t.h, mode, self.R, self.ldm, ptrs,
                                                       mc_out, C.byref(gs), C.c_void_p(s))
            if rc != A.SPLATT_SUCCESS:
                raise RuntimeError(f"splatt_b200_mttkrp_multicast_sync failed ({rc})")
            self._barriers += 1          # the kernel's tail IS a group barrier
            return self.bufs[mode]
        rc = lib.splatt_b200_mttkrp_multicast(self.t.h, mode, self.R, self.ldm, ptrs, mc_out,
                                              C.c_void_p(s))
        if rc != A.SPLATT_SUCCESS:
            raise RuntimeError(f"splatt_b200_mttkrp_multicast failed ({rc})")
        self._barrier(mode)
        return self.bufs[mode]

    def release(self, mode):
        """Call when the result of `mode` has been consumed: re-zero it for its next use."""
        self.bufs[mode].zero_()
        self._released_at[mode] = self._barriers


def cpd_als_sharded(tensor, ncolumns, init_factors, ttnormsq, niters=50, tol=1e-5, fused=True,
                    group=None, verbose=False):
    """CPD-ALS over a sharded tensor: one process per GPU.

    Per mode: MTTKRP of this rank's shard, summed over ranks (fused NVLink-multicast exchange
    when available and `fused`, else NCCL all-reduce), then the dense tail ONCE, on rank 0,
    with the same device kernels `splatt_cpd_als` uses (`splatt_b200_als_tail_*`); the new
    factor is broadcast.  Factors are therefore single-valued: replicated tails agree only
    to rounding (atomics and multimem.red arrive in a different order on every GPU) and on
    ill-conditioned problems such replicas drift apart over the iterations.  The
    iteration is the reference's (src/cpd.c:318-373).  Also runs unsharded (no process group).

    init_factors: list of torch CUDA float64 (dims[m] x ncolumns) -- identical on every rank.
    Returns (fit, lambda (numpy), factors (torch, unnormalised as left by the last iteration),
             seconds per iteration list)."""
    import time
    import numpy as np
    import torch
    import torch.distributed as dist
    lib = tensor.lib
    N, R = tensor.nmodes, ncolumns
    ldm = R + (R & 1)
    dev = init_factors[0].device
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    fx = None
    if world > 1 and fused:
        fx = FusedExchange(tensor, R, group)
        if not fx.available():
            fx = None
    mats = []
    for m in range(N):
        a = torch.zeros((tensor.dims[m], ldm), dtype=torch.float64, device=dev)
        a[:, :R] = init_factors[m]
        mats.append(a)
    outs = [torch.empty((tensor.dims[m], ldm), dtype=torch.float64, device=dev) for m in range(N)]
    views = [a[:, :R] for a in mats]
    stream = torch.cuda.current_stream().cuda_stream
    rank = dist.get_rank(group) if world > 1 else 0
    src0 = (dist.get_global_rank(group, 0) if group is not None else 0) if world > 1 else 0
    lead = rank == 0
    h = C.c_void_p()
    rc = lib.splatt_b200_als_tail_create(N, R, ldm, C.c_void_p(stream), C.byref(h))
    if rc != A.SPLATT_SUCCESS:
        raise RuntimeError(f"splatt_b200_als_tail_create failed ({rc})")

    def ptr(t):
        return C.cast(C.c_void_p(t.data_ptr()), A.val_p)
    try:
        if lead:
            for m in range(N):
                lib.splatt_b200_als_tail_gram(h, m, ptr(mats[m]), tensor.dims[m])
        fit = oldfit = 0.0
        lam = np.zeros(R)
        lam_c = (C.c_double * R)()
        times = []
        for it in range(niters):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for m in range(N):
                if fx is not None:
                    m1 = fx.mttkrp(m, mats)[:, :ldm]
                else:
                    tensor.mttkrp(m, mats, outs[m], ncolumns=R)
                    m1 = all_reduce_output(outs[m], group)
                if lead:
                    rc = lib.splatt_b200_als_tail_update(h, m, ptr(m1), ptr(mats[m]),
                                                         tensor.dims[m], 1 if it == 0 else 0)
                    if rc != A.SPLATT_SUCCESS:
                        raise RuntimeError("splatt_b200_als_tail_update failed")
                if world > 1:
                    dist.broadcast(mats[m], src=src0, group=group)    # one factor for everybody
                last_m1 = m1
                if fx is not None and m != N - 1:
                    fx.release(m)
            f = C.c_double()
            if lead:
                rc = lib.splatt_b200_als_tail_fit(h, ptr(mats[N - 1]), ptr(last_m1),
                                                  tensor.dims[N - 1], float(ttnormsq), C.byref(f),
                                                  lam_c)
                if rc != A.SPLATT_SUCCESS:
                    raise RuntimeError("splatt_b200_als_tail_fit failed")
            if fx is not None:
                fx.release(N - 1)
            fit = f.value
            lam = np.array(lam_c[:])
            if world > 1:
                # one fit, one stop decision for the whole group (the reference all-reduces its
                # fit, src/mpi/mpi_cpd.c:760-775)
                ft = torch.zeros(1 + R, dtype=torch.float64, device=dev)
                if lead:
                    ft[0] = fit
                    ft[1:] = torch.from_numpy(lam).to(dev)
                dist.broadcast(ft, src=src0, group=group)
                fh = ft.cpu().numpy()
                fit, lam = float(fh[0]), fh[1:].copy()
            times.append(time.perf_counter() - t0)
            if verbose:
                print(f"  its = {it + 1:3d} ({times[-1]:.4f}s)  fit = {fit:.5f}  "
                      f"delta = {fit - oldfit:+.4e}", flush=True)
            if fit == 1.0 or (it > 0 and abs(fit - oldfit) < tol):
                break
            oldfit = fit
        return fit, lam, [v.clone() for v in views], times
    finally:
        lib.splatt_b200_als_tail_free(h)
