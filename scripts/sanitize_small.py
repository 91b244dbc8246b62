"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / initcheck): every
kernel family once on tiny tensors -- root / internal / leaf, 3 and 4 modes (ancestor-id side
stream), the device ALS tail (generic and register-tiled kernels), the multi-GPU engine on a
device list that names GPU 0 twice.   compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import splatt_b200 as S  # noqa: E402
from tests.util import cover_all_slices, factor_mats, random_coo  # noqa: E402

for dims, nnz, R in (((60, 40, 50), 3000, 32), ((20, 15, 25, 10), 2500, 16), ((9, 8, 7, 6, 5), 1500, 6)):
    dims, inds, vals = random_coo(dims, nnz, seed=1)
    mats = factor_mats(dims, R)
    dm = [torch.from_numpy(x).cuda() for x in mats]
    for layout in (0, 1):
        T = S.Tensor.from_coo(dims, inds, vals, layout=layout, csf_alloc=0)
        for m in range(len(dims)):
            out = torch.empty((dims[m], R), dtype=torch.float64, device="cuda")
            T.mttkrp(m, dm, out)
        torch.cuda.synchronize()
        T.free()
dims, inds, vals = cover_all_slices(*random_coo((90, 70, 50), 4000, seed=2))
o = S.default_opts()
o[3], o[1], o[4] = 3, 0.0, 0
csf = S.csf_alloc(dims, inds, vals, o)
for R, generic in ((20, "0"), (20, "1"), (40, "0")):
    os.environ["SPLATT_B200_TAIL_GENERIC"] = generic
    fit, lam, fac = S.cpd_als(csf.ptr, R, o, seed=1)
    assert np.isfinite(fit)
mats = factor_mats(dims, 16)
ws = S.MttkrpWorkspace(csf.ptr, 16, o)
for m in range(3):
    ws.mttkrp_csf(mats, m, np.empty((dims[m], 16)))
ws.free()
mg = S.MultiGpu(csf.ptr, int(o[6]), 16, [0, 0])
for m in range(3):
    mg.mttkrp_host(m, mats)
fit, lam, fac = mg.cpd_als(o, seed=1)
mg.free()
print("sanitize_small ok, fit", fit)
