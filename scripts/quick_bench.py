"""Quick device-side timing of the MTTKRP kernels on a synthetic uniform tensor.
usage: python scripts/quick_bench.py [dim] [nnz] [R] [nmodes] [layout]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import splatt_b200 as S  # noqa: E402

dim = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
nnz = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
R = int(sys.argv[3]) if len(sys.argv) > 3 else 32
N = int(sys.argv[4]) if len(sys.argv) > 4 else 3
layout = int(sys.argv[5]) if len(sys.argv) > 5 else 0

g = torch.Generator(device="cuda").manual_seed(1)
dims = [dim] * N
if os.environ.get("DIMS"):                      # e.g. DIMS=10000,10000,64 (long fibers)
    dims = [int(x) for x in os.environ["DIMS"].split(",")]
    N = len(dims)
if len(sys.argv) > 6 and sys.argv[6] == "zipf":
    # config-5 shape family: dim x dim x 1000, Zipf(1.0) on the two long modes
    dims = [dim, dim, 1000]
    N = 3
    def zipf(d):
        w = 1.0 / torch.arange(1, d + 1, device="cuda", dtype=torch.float64)
        cdf = torch.cumsum(w / w.sum(), 0)
        u = torch.rand(nnz, device="cuda", dtype=torch.float64, generator=g)
        r = torch.searchsorted(cdf, u).clamp_(max=d - 1)
        perm = torch.randperm(d, device="cuda", generator=g)
        return perm[r].to(torch.int32)
    ind = [zipf(dim), zipf(dim), torch.randint(0, 1000, (nnz,), device="cuda", dtype=torch.int32, generator=g)]
else:
    ind = [torch.randint(0, d, (nnz,), device="cuda", dtype=torch.int32, generator=g) for d in dims]
vals = torch.rand(nnz, device="cuda", dtype=torch.float64, generator=g)
t0 = time.time()
T = S.Tensor.from_coo(dims, ind, vals, layout=layout, csf_alloc=int(os.environ.get("CSF_ALLOC", "1")), verbosity=3,
                      ncolumns_hint=R, ktile=int(os.environ.get("KTILE", "-1")))
torch.cuda.synchronize()
print(f"build {time.time()-t0:.2f}s  device MB {T.device_bytes/1e6:.0f}")
mats = [torch.rand(d, R, device="cuda", dtype=torch.float64) * 6 - 3 for d in dims]
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda", dtype=torch.float32)
for m in range(N):
    info = T.mode_info(m, R)
    out = torch.empty(dims[m], R, device="cuda", dtype=torch.float64)
    for _ in range(3):
        T.mttkrp(m, mats, out)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        T.mttkrp(m, mats, out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    print(f"mode {m} {info['kind']:8s} nfibs {info['nfibs']}  {ms*1e3:8.1f} us  "
          f"{nnz*R/ms/1e6:8.1f} G nnz*R/s   alg {info['alg_bytes']/1e6:.0f} MB -> "
          f"{info['alg_bytes']/ms/1e6:.0f} GB/s")
