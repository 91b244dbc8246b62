"""Fit trajectories of splatt_cpd_als: single GPU vs the multi-GPU engine (debug aid)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import splatt_b200 as S  # noqa: E402

dim, nnz, R, iters = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
devlists = sys.argv[5:] or ["0,0"]
dims = [dim] * 3
g = torch.Generator(device="cuda").manual_seed(1)
ind = [torch.randint(0, dim, (nnz,), device="cuda", dtype=torch.int32, generator=g) for _ in range(3)]
vals = torch.rand(nnz, device="cuda", dtype=torch.float64, generator=g)
o = S.default_opts()
csf = S.csf_alloc(dims, [i.cpu().numpy() for i in ind], vals.cpu().numpy(), o)
o[3], o[1], o[4] = iters, 0.0, 1
print("== single", flush=True)
fit, lam, fac = S.cpd_als(csf.ptr, R, o, seed=1)
print("single fit", fit, flush=True)
for dl in devlists:
    os.environ["SPLATT_B200_DEVICES"] = dl
    print("== devices", dl, flush=True)
    fit2, lam2, fac2 = S.cpd_als(csf.ptr, R, o, seed=1)
    print("multi fit", fit2, "max factor diff", max(float(np.abs(a - b).max()) for a, b in zip(fac, fac2)), flush=True)
    os.environ.pop("SPLATT_B200_DEVICES")
