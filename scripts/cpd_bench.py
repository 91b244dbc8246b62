"""CPD-ALS iteration time: ours (device tail / host tail) vs the reference, config 2."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import splatt_b200 as S  # noqa: E402
from oracle import ref  # noqa: E402

dim, nnz, R, iters = 10000, 10_000_000, 32, 5
g = torch.Generator(device="cuda").manual_seed(1)
ind = [torch.randint(0, dim, (nnz,), device="cuda", dtype=torch.int32, generator=g) for _ in range(3)]
vals = torch.rand(nnz, device="cuda", dtype=torch.float64, generator=g)
ind_h = [i.cpu().numpy() for i in ind]
vals_h = vals.cpu().numpy()
o = S.default_opts()
o[3], o[1], o[4] = iters, 0.0, 1
csf = S.csf_alloc([dim] * 3, ind_h, vals_h, o)
for hs in ("0", "1"):
    os.environ["SPLATT_B200_HOST_SOLVE"] = hs
    t0 = time.time()
    fit, lam, fac = S.cpd_als(csf.ptr, R, o, seed=1)
    print(f"ours host_solve={hs}: total {time.time()-t0:.3f}s for {iters} its, fit {fit:.6f}", flush=True)
if ref.available():
    ro = ref.default_opts()
    ro[3], ro[1], ro[4] = iters, 0.0, 1
    ro[0] = int(os.environ.get("SPLATT_REF_THREADS", "32"))
    tt = ref.RefTensor.from_coo([dim] * 3, [i.astype(np.uint64) for i in ind_h], vals_h)
    rc = ref.RefCsf(tt, ro)
    t0 = time.time()
    fit, lam, fac = rc.cpd_als(R, 1)
    print(f"reference ({int(ro[0])} threads): total {time.time()-t0:.3f}s for {iters} its, fit {fit:.6f}")
