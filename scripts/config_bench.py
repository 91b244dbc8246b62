"""BASELINE.json configs 2-5 as MTTKRP sweeps (strong scaling: the tensor is fixed, N ranks
share it).  python scripts/config_bench.py <config> [--nccl]   (under torchrun for N > 1)

  2: uniform 3-mode 10K^3,        10 M nnz, R = 32
  3: uniform 4-mode 5K^4,         50 M nnz, R = 16
  4: uniform 3-mode 100K^3,      100 M nnz, R = 32
  5: Zipf(1.0) 1M x 1M x 1K,     200 M nnz, R = 64   (modes 0,1 Zipf through a random relabelling)
Prints one JSON line (rank 0): per-mode times, nnz*R/s, exchange used.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import splatt_b200 as S  # noqa: E402
from splatt_b200 import parallel  # noqa: E402

CONFIGS = {
    "2": ([10_000] * 3, 10_000_000, 32, False, 1),
    "3": ([5_000] * 4, 50_000_000, 16, False, 2),
    "4": ([100_000] * 3, 100_000_000, 32, False, 3),
    "5": ([1_000_000, 1_000_000, 1_000], 200_000_000, 64, True, 4),
}
cfg = sys.argv[1] if len(sys.argv) > 1 else "2"
force_nccl = "--nccl" in sys.argv
dims, nnz, R, zipf, seed = CONFIGS[cfg]
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", ""):
    os.environ["NCCL_DEBUG"] = "WARN"
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
g = torch.Generator(device=dev).manual_seed(seed)


def zipf_idx(d):
    w = 1.0 / torch.arange(1, d + 1, device=dev, dtype=torch.float64)
    cdf = torch.cumsum(w / w.sum(), 0)
    u = torch.rand(nnz, device=dev, dtype=torch.float64, generator=g)
    r = torch.searchsorted(cdf, u).clamp_(max=d - 1)
    del u
    return torch.randperm(d, device=dev, generator=g)[r].to(torch.int32)


ind = []
for m, d in enumerate(dims):
    if zipf and m < 2:
        ind.append(zipf_idx(d))
    else:
        ind.append(torch.randint(0, d, (nnz,), device=dev, dtype=torch.int32, generator=g))
vals = torch.rand(nnz, device=dev, dtype=torch.float64, generator=g)
t0 = time.time()
T = S.Tensor.from_coo(dims, ind, vals, shard_rank=rank, shard_count=world)
torch.cuda.synchronize()
build_s = time.time() - t0
del ind, vals
torch.cuda.empty_cache()
mats = [torch.rand(d, R, device=dev, dtype=torch.float64, generator=g) * 6 - 3 for d in dims]
outs = [torch.empty(d, R, device=dev, dtype=torch.float64) for d in dims]
fx = None
if world > 1 and not force_nccl:
    fx = parallel.FusedExchange(T, R)
    if not fx.available():
        fx = None
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
N = len(dims)


def one(m):
    if fx is not None:
        fx.mttkrp(m, mats)
        fx.release(m)
    else:
        T.mttkrp(m, mats, outs[m])
        if world > 1:
            dist.all_reduce(outs[m])


per_mode = []
for m in range(N):
    for _ in range(3):
        one(m)
    ts = []
    for _ in range(10):
        flush.zero_()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); one(m); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = torch.tensor([float(np.median(ts))], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    per_mode.append(t.item())
if rank == 0:
    info = [T.mode_info(m, R) for m in range(N)]
    print(json.dumps({
        "config": cfg, "dims": dims, "nnz_total": nnz, "rank": R, "zipf": zipf, "n_gpus": world,
        "exchange": "none" if world == 1 else ("fused multimem.red" if fx is not None else "NCCL all-reduce"),
        "per_mode_ms": per_mode,
        "per_mode_nnzR_per_s": [nnz * R / (t * 1e-3) for t in per_mode],
        "mean_nnzR_per_s": nnz * R * N / (sum(per_mode) * 1e-3),
        "nfibs_local": [i["nfibs"] for i in info], "alg_bytes_local": [i["alg_bytes"] for i in info],
        "build_seconds": build_s, "device_bytes_local": T.device_bytes}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
