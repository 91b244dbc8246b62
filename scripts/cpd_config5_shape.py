"""CPD-ALS iteration time on a config-5-shaped tensor (1M x 1M x 1K, R = 64): the multi-GPU
engine on 1..n GPUs, row-partitioned tail vs tail on device 0.
usage: python scripts/cpd_config5_shape.py [nnz] [ngpus...]   (SPLATT_B200_MULTI_TIMING=1 for phases)"""
import os
import sys
import time

import torch

sys.path.insert(0, ".")
import splatt_b200 as S  # noqa: E402
from splatt_b200 import _abi as A  # noqa: E402

dims = [1000000, 1000000, 1000]
nnz = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
gpus = [int(x) for x in sys.argv[2:]] or [1, 2]
R = 64
g = torch.Generator(device="cuda").manual_seed(4)
ind = [torch.randint(0, d, (nnz,), device="cuda", dtype=torch.int32, generator=g) for d in dims]
for m, d in enumerate(dims):
    n = min(d, nnz)
    ind[m][:n] = torch.arange(n, device="cuda", dtype=torch.int32)
vals = torch.rand(nnz, device="cuda", dtype=torch.float64, generator=g)
o = S.default_opts()
csf = S.csf_alloc(dims, [i.cpu().numpy() for i in ind], vals.cpu().numpy(), o)
del ind, vals
torch.cuda.empty_cache()


def timed(fn):
    def run(n):
        oo = S.default_opts()
        oo[3], oo[1], oo[4] = n, 0.0, 0
        t0 = time.perf_counter()
        fit = fn(oo)[0]
        return time.perf_counter() - t0, fit
    run(1)
    (ta, f), (tb, _) = run(42), run(2)
    return (ta - tb) / 40 * 1e3, f


for n in gpus:
    for part in ("1", "0"):
        if n == 1 and part == "0":
            continue
        os.environ["SPLATT_B200_PARTITIONED_TAIL"] = part
        mg = S.MultiGpu(csf.ptr, A.CSF_TWOMODE, R, list(range(n)))
        ms, fit = timed(lambda oo: mg.cpd_als(oo, seed=1))
        print(f"{n} GPU(s) partitioned_tail={part} multicast={mg.multicast}: {ms:.2f} ms/iteration "
              f"fit {fit:.3e}", flush=True)
        mg.free()
