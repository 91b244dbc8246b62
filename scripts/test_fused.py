"""torchrun --nproc-per-node N scripts/test_fused.py : fused MTTKRP+exchange (NVLink multicast)
vs kernel + NCCL all-reduce: same result, timing of both."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import splatt_b200 as S  # noqa: E402
from splatt_b200 import parallel  # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
dim, nnz, R = 10000, 10_000_000 * world, 32
g = torch.Generator(device=dev).manual_seed(1)
ind = [torch.randint(0, dim, (nnz,), device=dev, dtype=torch.int32, generator=g) for _ in range(3)]
vals = torch.rand(nnz, device=dev, dtype=torch.float64, generator=g)
T = S.Tensor.from_coo([dim] * 3, ind, vals, shard_rank=rank, shard_count=world)
mats = [torch.rand(dim, R, device=dev, dtype=torch.float64, generator=g) * 6 - 3 for _ in range(3)]
outs = [torch.empty(dim, R, device=dev, dtype=torch.float64) for _ in range(3)]
fx = parallel.FusedExchange(T, R)
if rank == 0:
    print("multicast available:", fx.available(), fx.error, flush=True)
    try:
        import torch.distributed._symmetric_memory as sm
        b = sm.empty((16,), dtype=torch.float64, device=dev)
        print("symm empty ok", flush=True)
    except Exception as e:
        print("symm empty failed:", e, flush=True)


def sweep_nccl():
    for m in range(3):
        parallel.sharded_mttkrp(T, m, mats, outs[m])


def sweep_fused():
    res = []
    for m in range(3):
        res.append(fx.mttkrp(m, mats))
    return res


def timeit(fn, n=30):
    for _ in range(5):
        fn()
        if fn is sweep_fused:
            for m in range(3):
                fx.release(m)
    dist.barrier(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        if fn is sweep_fused:
            for m in range(3):
                fx.release(m)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = torch.tensor([float(np.median(ts))], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


sweep_nccl()
if fx.available():
    worst = 0.0
    for rep in range(20):        # repeated: the barrier must order the remote reductions every time
        res = sweep_fused()
        for m in range(3):
            err = (torch.linalg.norm(res[m] - outs[m]) / torch.linalg.norm(outs[m])).item()
            worst = max(worst, err)
            assert err < 1e-12, (rep, m, err)
        for m in range(3):
            fx.release(m)
    if rank == 0:
        print(f"fused vs nccl, 20 sweeps x 3 modes: worst rel err {worst:.2e}", flush=True)
t_n = timeit(sweep_nccl)
if rank == 0:
    print(f"N={world} nccl  sweep {t_n*1e3:.1f} us  -> {nnz*R*3/t_n/1e9:.2f} T nnz*R/s", flush=True)
if fx.available():
    t_f = timeit(sweep_fused)
    if rank == 0:
        print(f"N={world} fused sweep {t_f*1e3:.1f} us  -> {nnz*R*3/t_f/1e9:.2f} T nnz*R/s", flush=True)
# ---- distributed CPD-ALS: fused exchange vs NCCL exchange, same trajectory; iteration time
init = [torch.rand(dim, R, device=dev, dtype=torch.float64, generator=torch.Generator(device=dev).manual_seed(5)) * 6 - 3
        for _ in range(3)]
tt = float((vals * vals).sum().item())
fit_f, lam_f, fac_f, t_f = parallel.cpd_als_sharded(T, R, init, tt, niters=6, tol=0.0, fused=True)
fit_n, lam_n, fac_n, t_n = parallel.cpd_als_sharded(T, R, init, tt, niters=6, tol=0.0, fused=False)
if rank == 0:
    print(f"CPD-ALS N={world}: fit fused {fit_f:.10f} nccl {fit_n:.10f}; "
          f"iteration fused {np.median(t_f[1:])*1e3:.3f} ms, nccl {np.median(t_n[1:])*1e3:.3f} ms", flush=True)
assert abs(fit_f - fit_n) < 1e-9
dist.barrier()
dist.destroy_process_group()
