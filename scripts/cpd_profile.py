"""One short CPD-ALS (splatt_cpd_als, device tail) for a launch list / kernel-time breakdown.
usage: python scripts/cpd_profile.py <dim0> <dim1> <dim2> <nnz> <R> [niter]
Run under   ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_gram|k_form_chol|k_solve_rows|k_colnorm|k_finish_lambda|k_scale_cols|k_inner|mttkrp_stream' --csv --log-file ...
or plain (prints wall time per iteration from the library's own report)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import splatt_b200 as S  # noqa: E402

d0, d1, d2, nnz, R = (int(x) for x in sys.argv[1:6])
niter = int(sys.argv[6]) if len(sys.argv) > 6 else 3
dims = [d0, d1, d2]
g = torch.Generator(device="cuda").manual_seed(4)
ind = [torch.randint(0, d, (nnz,), device="cuda", dtype=torch.int32, generator=g) for d in dims]
# cover every slice so that no Gram row is empty
for m, d in enumerate(dims):
    n = min(d, nnz)
    ind[m][:n] = torch.arange(n, device="cuda", dtype=torch.int32)
vals = torch.rand(nnz, device="cuda", dtype=torch.float64, generator=g)
o = S.default_opts()
csf = S.csf_alloc(dims, [i.cpu().numpy() for i in ind], vals.cpu().numpy(), o)
del ind, vals
torch.cuda.empty_cache()
o[3], o[1], o[4] = niter, 0.0, 1
t0 = time.perf_counter()
fit, lam, fac = S.cpd_als(csf.ptr, R, o, seed=1)
print(f"cpd_als {niter} its: {time.perf_counter()-t0:.3f}s  fit {fit:.5f}")
