"""CPU: pin the oracle.

(1) oracle/restate.c against the committed golden vectors, which were produced by
    the reference's own gold (mttkrp_stream) on the reference's own fixtures
    (tests/golden/make_golden.py).  Tolerance = the reference's own test
    tolerance, abs 1e-10 (tests/mttkrp_test.c:25-30), plus 1e-12 relative.
(2) oracle/restate.c against the compiled reference (oracle/_ref) on seeded
    random tensors: CSF arrays bit-exact, MTTKRP (COO and CSF walk) 1e-12.
"""
from pathlib import Path

import numpy as np
import pytest

from oracle import restate
from tests.util import cover_all_slices, factor_mats, random_coo, rel_fro

GOLD = Path(__file__).parent / "golden"
NAMES = ["small", "med", "small4", "med4", "med5", "small4_zeroidx"]


def load_golden(name):
    z = np.load(GOLD / f"{name}.npz")
    dims = [int(d) for d in z["dims"]]
    inds = [z["ind"][m].astype(np.uint64) for m in range(len(dims))]
    return z, dims, inds, z["vals"]


def check_against_golden(z, R, m, out):
    rows = z[f"gold_R{R}_m{m}_rows"]
    assert np.allclose(out[rows], z[f"gold_R{R}_m{m}_vals"], rtol=1e-12, atol=1e-10)
    assert np.allclose(out.sum(axis=0), z[f"gold_R{R}_m{m}_colsum"], rtol=1e-11, atol=1e-9)
    assert abs(np.linalg.norm(out) - z[f"gold_R{R}_m{m}_fro"][0]) <= 1e-11 * z[f"gold_R{R}_m{m}_fro"][0]


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("R", [3, 16])
def test_restatement_coo_vs_golden(name, R):
    z, dims, inds, vals = load_golden(name)
    mats = factor_mats(dims, R, seed=R)
    for m in range(len(dims)):
        check_against_golden(z, R, m, restate.mttkrp_coo(dims, inds, vals, mats, m))


@pytest.mark.parametrize("name", NAMES)
def test_restatement_csf_vs_golden(name):
    """CSF build + tree-walk MTTKRP (root / internal / leaf) on the default TWOMODE CSFs."""
    z, dims, inds, vals = load_golden(name)
    R = 3
    mats = factor_mats(dims, R, seed=R)
    perms, mode_map = restate.csf_policy(dims, 1)
    csfs = [restate.OracleCsf(dims, inds, vals, p) for p in perms]
    for c, csf in enumerate(csfs):
        a = csf.arrays()
        assert a["dim_perm"] == [int(x) for x in z[f"csf{c}_perm"]]
        assert a["nfibs"] == [int(x) for x in z[f"csf{c}_nfibs"]]
        assert a["nfibs"][-1] == len(vals)
    for m in range(len(dims)):
        # every mode on every CSF: exercises all output depths
        for csf in csfs:
            check_against_golden(z, R, m, csf.mttkrp(mats, m))


def test_zero_index_fixture_equals_one_index_fixture():
    """tests/io_test.c:35-53: 0- and 1-indexed files describe the same tensor."""
    _, d0, i0, v0 = load_golden("small4")
    _, d1, i1, v1 = load_golden("small4_zeroidx")
    assert d0 == d1 and np.array_equal(v0, v1)
    for a, b in zip(i0, i1):
        assert np.array_equal(a, b)


SPECS = [((30, 50), 700), ((13, 7, 11), 150), ((40, 30, 50, 20), 5000), ((12, 15, 10, 20, 9), 3000),
         ((5, 6, 4, 7, 3, 5), 1500)]


@pytest.mark.parametrize("spec", SPECS)
@pytest.mark.parametrize("alloc", [0, 1, 2])
def test_restatement_vs_compiled_reference(refmod, spec, alloc):
    dims, inds, vals = random_coo(spec[0], spec[1], seed=11)
    mats = factor_mats(dims, 5)
    tt = refmod.RefTensor.from_coo(dims, inds, vals)
    gold = [tt.mttkrp_stream(mats, m) for m in range(len(dims))]
    perms, mode_map = restate.csf_policy(dims, alloc)
    o = refmod.default_opts()
    o[0] = 2
    o[6] = alloc
    rc = refmod.RefCsf(tt, o)
    assert rc.count == len(perms)
    for c, perm in enumerate(perms):
        oc = restate.OracleCsf(dims, inds, vals, perm)
        a, b = oc.arrays(), rc.arrays(c)
        assert a["dim_perm"] == b["dim_perm"] and a["nfibs"] == b["nfibs"]
        for l in range(len(dims)):
            assert (a["fids"][l] is None) == (b["fids"][l] is None)
            if a["fids"][l] is not None:
                assert np.array_equal(a["fids"][l], b["fids"][l])
        for l in range(len(dims) - 1):
            assert np.array_equal(a["fptr"][l], b["fptr"][l])
        assert np.array_equal(a["vals"], b["vals"])
        for m in range(len(dims)):
            assert rel_fro(oc.mttkrp(mats, m), gold[m]) < 1e-12
    for m in range(len(dims)):
        assert rel_fro(restate.mttkrp_coo(dims, inds, vals, mats, m), gold[m]) < 1e-13
        if len(dims) > 2:
            # reference production path == reference gold.  (For matrices the reference's CSF
            # leaf-mode kernel does not reproduce its own gold -- src/mttkrp.c:860-943 assumes
            # nmodes >= 3 -- so 2-mode parity is anchored on the COO gold only.)
            out, _ = rc.mttkrp_csf(mats, m)
            assert rel_fro(out, gold[m]) < 1e-12


def test_restatement_cpd_vs_compiled_reference(refmod):
    dims, inds, vals = random_coo((60, 50, 40), 6000, seed=3)
    dims, inds, vals = cover_all_slices(dims, inds, vals)
    o = refmod.default_opts()
    o[0], o[3], o[1], o[4] = 1, 8, 0.0, 0
    tt = refmod.RefTensor.from_coo(dims, inds, vals)
    rc = refmod.RefCsf(tt, o)
    f1, l1, F1 = rc.cpd_als(6, 7)
    f2, l2, F2 = restate.cpd_als(dims, inds, vals, 6, 8, 0.0, 7)
    assert abs(f1 - f2) < 1e-12
    assert np.allclose(l1, l2, rtol=1e-10)
    for a, b in zip(F1, F2):
        assert np.allclose(a, b, rtol=1e-9, atol=1e-12)


def test_compiled_reference_matches_golden(refmod):
    """The compiled reference reproduces the committed vectors (guards the build flags)."""
    z, dims, inds, vals = load_golden("med")
    tt = refmod.RefTensor.from_coo(dims, inds, vals)
    mats = factor_mats(dims, 3, seed=3)
    for m in range(3):
        check_against_golden(z, 3, m, tt.mttkrp_stream(mats, m))
