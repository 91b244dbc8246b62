"""GPU parity: the CUDA path, called through the C ABI, against the reference.

Gold is the reference's own gold (tests/mttkrp_test.c:66): mttkrp_stream on COO,
computed by the unmodified reference in oracle/_ref.  The north-star bar is
1e-6 relative Frobenius; fp64 with a different summation order lands ~1e-15, so
the tests assert 1e-11 (and the reference's own abs 1e-10 scaled by magnitude).
"""
import os

import numpy as np
import pytest

from tests.util import cover_all_slices, factor_mats, random_coo, rel_fro

pytestmark = pytest.mark.gpu

TOL = 1e-11

TENSORS = {
    "t3_small": ((13, 7, 11), 150),
    "t3_mid": ((300, 200, 400), 20000),
    "t3_long_fibers": ((50, 40, 3000), 30000),
    "t3_skew": ((2000, 1500, 60), 40000),
    "t4": ((40, 30, 50, 20), 15000),
    "t5": ((12, 15, 10, 20, 9), 8000),
    "t6": ((6, 7, 5, 8, 9, 4), 4000),
    "t8": ((4, 3, 5, 4, 3, 6, 2, 5), 3000),
}


def _tensor(name):
    dims, nnz = TENSORS[name]
    skew = [1.0, 1.0, 0] if name == "t3_skew" else None
    return random_coo(dims, nnz, seed=hash(name) % 1000, skew=skew)


@pytest.fixture(scope="module")
def S():
    import splatt_b200
    return splatt_b200


def _gold(refmod, dims, inds, vals, mats):
    tt = refmod.RefTensor.from_coo(dims, inds, vals)
    return tt, [tt.mttkrp_stream(mats, m) for m in range(len(dims))]


@pytest.mark.parametrize("name", list(TENSORS))
@pytest.mark.parametrize("R", [3, 16, 32])
@pytest.mark.parametrize("layout", ["allroot", "asgiven"])
def test_dropin_mttkrp_on_reference_csf(S, refmod, name, R, layout, monkeypatch):
    """splatt_mttkrp (ours) on CSFs built by the reference's csf_alloc, every mode,
    every allocation policy -- the drop-in scenario."""
    monkeypatch.setenv("SPLATT_B200_LAYOUT", layout)
    dims, inds, vals = _tensor(name)
    mats = factor_mats(dims, R)
    tt, gold = _gold(refmod, dims, inds, vals, mats)
    for alloc in (0, 1, 2):
        o = refmod.default_opts()
        o[0] = 1
        o[6] = alloc
        csf = refmod.RefCsf(tt, o)
        for m in range(len(dims)):
            out = S.mttkrp(m, R, csf.ptr, mats, o)
            assert rel_fro(out, gold[m]) < TOL, (name, R, layout, alloc, m)
        csf.free()


@pytest.mark.parametrize("R", [1, 2, 5, 10, 48, 64, 70, 130])
def test_ranks(S, refmod, R):
    dims, inds, vals = _tensor("t3_mid")
    mats = factor_mats(dims, R)
    tt, gold = _gold(refmod, dims, inds, vals, mats)
    o = refmod.default_opts()
    o[0] = 1
    csf = refmod.RefCsf(tt, o)
    for m in range(3):
        out = S.mttkrp(m, R, csf.ptr, mats, o)
        assert rel_fro(out, gold[m]) < TOL, (R, m)


@pytest.mark.parametrize("name", ["t3_mid", "t4", "t5"])
@pytest.mark.parametrize("tilelevel", [0, 1, 2])
def test_dropin_on_densetiled_reference_csf(S, refmod, name, tilelevel):
    """Reference CSFs built with SPLATT_DENSETILE (tests/mttkrp_test.c:201-259)."""
    dims, inds, vals = _tensor(name)
    R = 8
    mats = factor_mats(dims, R)
    tt, gold = _gold(refmod, dims, inds, vals, mats)
    for alloc in (0, 1, 2):
        o = refmod.default_opts()
        o[0] = 3
        o[6] = alloc
        o[7] = 1          # SPLATT_DENSETILE
        o[8] = tilelevel
        csf = refmod.RefCsf(tt, o)
        for m in range(len(dims)):
            out = S.mttkrp(m, R, csf.ptr, mats, o)
            assert rel_fro(out, gold[m]) < TOL, (name, tilelevel, alloc, m)
        csf.free()


@pytest.mark.parametrize("name", ["t3_small", "t3_mid", "t4", "t5", "t8"])
@pytest.mark.parametrize("alloc", [0, 1, 2])
def test_csf_alloc_matches_reference(S, refmod, name, alloc):
    """splatt_b200_csf_alloc builds bit-identical CSF arrays to the reference's csf_alloc
    (structure tests of tests/csf_test.c:31-61)."""
    dims, inds, vals = _tensor(name)
    o = refmod.default_opts()
    o[0] = 1
    o[6] = alloc
    tt = refmod.RefTensor.from_coo(dims, inds, vals)
    rc = refmod.RefCsf(tt, o)
    mine = S.csf_alloc(dims, inds, vals, o)
    assert mine.count == rc.count
    for c in range(rc.count):
        a, b = mine.arrays(c), rc.arrays(c)
        assert a["dim_perm"] == b["dim_perm"] and a["dim_iperm"] == b["dim_iperm"]
        assert a["nfibs"] == b["nfibs"] and a["nnz"] == b["nnz"] and a["dims"] == b["dims"]
        for l in range(len(dims)):
            if b["fids"][l] is None:
                assert a["fids"][l] is None
            else:
                assert np.array_equal(a["fids"][l], b["fids"][l]), (c, l)
        for l in range(len(dims) - 1):
            assert np.array_equal(a["fptr"][l], b["fptr"][l]), (c, l)
        assert np.array_equal(a["vals"], b["vals"])


def test_csf_alloc_gaps_keep_root_ids(S, refmod):
    """Empty root slices => fids[0] is materialised (src/csf.c:303-309)."""
    dims, inds, vals = random_coo((500, 40, 30), 300, seed=5)
    o = refmod.default_opts()
    o[6] = 0
    tt = refmod.RefTensor.from_coo(dims, inds, vals)
    rc = refmod.RefCsf(tt, o)
    mine = S.csf_alloc(dims, inds, vals, o)
    a, b = mine.arrays(0), rc.arrays(0)
    assert a["dim_perm"] == b["dim_perm"]
    for l in range(3):
        assert (a["fids"][l] is None) == (b["fids"][l] is None)
        if a["fids"][l] is not None:
            assert np.array_equal(a["fids"][l], b["fids"][l])
    R = 4
    mats = factor_mats(dims, R)
    gold = [tt.mttkrp_stream(mats, m) for m in range(3)]
    for m in range(3):
        assert rel_fro(S.mttkrp(m, R, mine.ptr, mats, o), gold[m]) < TOL
        assert rel_fro(S.mttkrp(m, R, rc.ptr, mats, o), gold[m]) < TOL


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("name", ["t3_mid", "t3_skew", "t4", "t5"])
def test_engine_device_path(S, refmod, name, layout):
    """Device-resident API: COO -> fiber streams -> MTTKRP on torch CUDA tensors."""
    import torch
    dims, inds, vals = _tensor(name)
    R = 16
    mats = factor_mats(dims, R)
    _, gold = _gold(refmod, dims, inds, vals, mats)
    T = S.Tensor.from_coo(dims, inds, vals, layout=layout, csf_alloc=1)
    dmats = [torch.from_numpy(m).cuda() for m in mats]
    for m in range(len(dims)):
        info = T.mode_info(m, R)
        if layout == 0:
            assert info["kind"] == "root"
        out = torch.empty((dims[m], R), dtype=torch.float64, device="cuda")
        T.mttkrp(m, dmats, out)
        torch.cuda.synchronize()
        assert rel_fro(out.cpu().numpy(), gold[m]) < TOL, (name, layout, m, info)
    T.free()


def test_sharded_partials_sum_to_whole(S, refmod):
    """shard_count > 1: per-shard partial outputs add up to the full MTTKRP
    (what the NCCL all-reduce does across ranks)."""
    import torch
    dims, inds, vals = _tensor("t3_skew")
    R = 32
    mats = factor_mats(dims, R)
    _, gold = _gold(refmod, dims, inds, vals, mats)
    dmats = [torch.from_numpy(m).cuda() for m in mats]
    for world in (2, 3, 8):
        for layout in (0, 1):
            shards = [S.Tensor.from_coo(dims, inds, vals, layout=layout, shard_rank=r,
                                        shard_count=world) for r in range(world)]
            assert sum(s.nnz_local for s in shards) == len(vals)
            for m in range(3):
                acc = torch.zeros((dims[m], R), dtype=torch.float64, device="cuda")
                for s in shards:
                    out = torch.empty_like(acc)
                    s.mttkrp(m, dmats, out)
                    acc += out
                assert rel_fro(acc.cpu().numpy(), gold[m]) < TOL, (world, layout, m)


def test_cpd_als_tracks_reference(S, refmod):
    """CPD-ALS: same seed, same iteration count => same fit trajectory end point.
    (The reference has no CPD result test; parity here is against the compiled
    reference itself.)"""
    dims, inds, vals = random_coo((60, 50, 40), 6000, seed=3)
    dims, inds, vals = cover_all_slices(dims, inds, vals)
    o = refmod.default_opts()
    o[0] = 1
    o[3] = 8          # iterations
    o[1] = 0.0        # tolerance: run them all
    o[4] = 0          # quiet
    tt = refmod.RefTensor.from_coo(dims, inds, vals)
    csf = refmod.RefCsf(tt, o)
    fit_ref, lam_ref, fac_ref = csf.cpd_als(6, seed=7)
    fit, lam, fac = S.cpd_als(csf.ptr, 6, o, seed=7)
    assert abs(fit - fit_ref) < 1e-8
    assert np.allclose(lam, lam_ref, rtol=1e-6, atol=1e-9)
    for a, b in zip(fac, fac_ref):
        assert np.allclose(a, b, rtol=1e-5, atol=1e-8)


def test_alias_output_with_own_factor(S, refmod):
    """matrices[mode] may alias matout (matlab/splatt_mttkrp.c:47-68)."""
    dims, inds, vals = _tensor("t3_mid")
    R = 8
    mats = factor_mats(dims, R)
    tt, gold = _gold(refmod, dims, inds, vals, mats)
    o = refmod.default_opts()
    csf = refmod.RefCsf(tt, o)
    ws = S.MttkrpWorkspace(csf.ptr, R, o)
    for m in range(3):
        mm = [x.copy() for x in mats]
        ws.mttkrp_csf(mm, m, mm[m])       # output written over the mode's own factor
        assert rel_fro(mm[m], gold[m]) < TOL
    ws.free()
