"""Generate the committed golden vectors from the reference's own fixtures.

Run HERE (where /root/reference and oracle/_ref exist):
    python tests/golden/make_golden.py

For every tensor the reference's MTTKRP tests use (tests/splatt_test.h:11-27:
small, med, small4, med4, med5 + small4_zeroidx) this stores, in one .npz:
  dims, ind (uint32, nmodes x nnz, as read by the reference's tt_read, i.e.
  0-based), vals, and for R in RANKS and every mode the output of the
  reference's gold mttkrp_stream (tests/mttkrp_test.c:66) on seeded factors
  (numpy default_rng(1000+R), uniform [-3,3] -- tests/util.py:factor_mats), plus
  the node counts of the reference's default TWOMODE CSFs.
The parity tests recompute the factors from the seed, so only tensors and
expected outputs are stored.  To keep the fixtures small an output of more than
SAMPLE rows is stored as: SAMPLE seeded rows (+ their row numbers), the column
sums and the Frobenius norm -- enough to pin every entry class (float64).
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import ref  # noqa: E402
from tests.util import factor_mats  # noqa: E402

REF_TENSORS = Path("/root/reference/tests/tensors")
NAMES = ["small", "med", "small4", "med4", "med5", "small4_zeroidx"]
RANKS = [3, 16]
SAMPLE = 2048

for name in NAMES:
    tt = ref.RefTensor.read(REF_TENSORS / f"{name}.tns")
    inds, vals = tt.coo()
    out = {"dims": np.array(tt.dims, dtype=np.uint64),
           "ind": np.stack(inds).astype(np.uint32), "vals": vals}
    for R in RANKS:
        mats = factor_mats(tt.dims, R, seed=R)
        for m in range(tt.nmodes):
            g = tt.mttkrp_stream(mats, m, nthreads=1)
            if g.shape[0] <= SAMPLE:
                rows = np.arange(g.shape[0])
            else:
                rows = np.sort(np.random.default_rng(7 + m).choice(g.shape[0], SAMPLE, replace=False))
            out[f"gold_R{R}_m{m}_rows"] = rows.astype(np.uint32)
            out[f"gold_R{R}_m{m}_vals"] = g[rows]
            out[f"gold_R{R}_m{m}_colsum"] = g.sum(axis=0)
            out[f"gold_R{R}_m{m}_fro"] = np.array([np.linalg.norm(g)])
    o = ref.default_opts()
    o[0] = 1
    csf = ref.RefCsf(tt, o)     # sorts tt (after the COO copy above)
    for c in range(csf.count):
        a = csf.arrays(c)
        out[f"csf{c}_perm"] = np.array(a["dim_perm"], dtype=np.uint64)
        out[f"csf{c}_nfibs"] = np.array(a["nfibs"], dtype=np.uint64)
    path = Path(__file__).parent / f"{name}.npz"
    np.savez_compressed(path, **out)
    print(name, tt.dims, tt.nnz, path.stat().st_size // 1024, "KiB")
