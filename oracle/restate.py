"""ctypes wrapper over oracle/restate.c (plain-C restatement of the reference's
hot-path algorithms).  TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SRC = HERE / "restate.c"
OUT = HERE / "_build" / "liboracle.so"

_lib = None
u64p = C.POINTER(C.c_uint64)
dp = C.POINTER(C.c_double)


def build() -> Path:
    OUT.parent.mkdir(exist_ok=True)
    if OUT.exists() and OUT.stat().st_mtime >= SRC.stat().st_mtime:
        return OUT
    subprocess.run(["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-o", str(OUT), str(SRC), "-lm"],
                   check=True)
    return OUT


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        lib = C.CDLL(str(build()))
        lib.oracle_csf_build.restype = C.c_void_p
        lib.oracle_cpd_als.restype = C.c_double
        lib.oracle_csf_policy.restype = C.c_uint64
        _lib = lib
    return _lib


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _pp(arrs, ctype):
    return (C.POINTER(ctype) * len(arrs))(*[a.ctypes.data_as(C.POINTER(ctype)) for a in arrs])


def mttkrp_coo(dims, ind, vals, mats, mode):
    lib = load()
    d = _u64(dims)
    inds = [_u64(i) for i in ind]
    v = np.ascontiguousarray(vals, dtype=np.float64)
    R = [m for i, m in enumerate(mats) if i != mode and m is not None][0].shape[1]
    ms = [np.ascontiguousarray(m, dtype=np.float64) if m is not None else np.zeros((int(d[i]), R))
          for i, m in enumerate(mats)]
    out = np.empty((int(d[mode]), R), dtype=np.float64)
    lib.oracle_mttkrp_coo(C.c_uint64(len(d)), d.ctypes.data_as(u64p), C.c_uint64(len(v)),
                          _pp(inds, C.c_uint64), v.ctypes.data_as(dp), C.c_uint64(R),
                          _pp(ms, C.c_double), C.c_uint64(mode), out.ctypes.data_as(dp))
    return out


def mode_order(dims, which, mode=0):
    lib = load()
    d = _u64(dims)
    perm = np.zeros(8, dtype=np.uint64)
    lib.oracle_mode_order(d.ctypes.data_as(u64p), C.c_uint64(len(d)), C.c_int(which),
                          C.c_uint64(mode), perm.ctypes.data_as(u64p))
    return [int(x) for x in perm[:len(d)]]


def csf_policy(dims, alloc):
    lib = load()
    d = _u64(dims)
    perms = np.zeros(64, dtype=np.uint64)
    mp = np.zeros(8, dtype=np.uint64)
    n = lib.oracle_csf_policy(d.ctypes.data_as(u64p), C.c_uint64(len(d)), C.c_int(alloc),
                              perms.ctypes.data_as(u64p), mp.ctypes.data_as(u64p))
    nm = len(d)
    return ([[int(x) for x in perms[c * 8:c * 8 + nm]] for c in range(int(n))],
            [int(x) for x in mp[:nm]])


class OracleCsf:
    def __init__(self, dims, ind, vals, perm):
        self.lib = load()
        self.dims = [int(x) for x in dims]
        self.nmodes = len(self.dims)
        d = _u64(dims)
        self._keep = ([_u64(i) for i in ind], np.ascontiguousarray(vals, dtype=np.float64))
        p = _u64(perm)
        self.nnz = len(self._keep[1])
        self.h = C.c_void_p(self.lib.oracle_csf_build(
            C.c_uint64(self.nmodes), d.ctypes.data_as(u64p), C.c_uint64(self.nnz),
            _pp(self._keep[0], C.c_uint64), self._keep[1].ctypes.data_as(dp), p.ctypes.data_as(u64p)))

    def arrays(self):
        nf = (C.c_uint64 * 8)()
        perm = (C.c_uint64 * 8)()
        fptr = (u64p * 8)()
        fids = (u64p * 8)()
        vals = dp()
        self.lib.oracle_csf_get(self.h, nf, perm, fptr, fids, C.byref(vals))
        n = self.nmodes
        out = {"nfibs": [int(nf[l]) for l in range(n)], "dim_perm": [int(perm[l]) for l in range(n)],
               "fptr": [], "fids": []}
        for l in range(n):
            out["fids"].append(None if not fids[l] else
                               np.ctypeslib.as_array(fids[l], shape=(int(nf[l]),)).copy())
            if l < n - 1:
                out["fptr"].append(np.ctypeslib.as_array(fptr[l], shape=(int(nf[l]) + 1,)).copy())
        out["vals"] = np.ctypeslib.as_array(vals, shape=(self.nnz,)).copy()
        return out

    def mttkrp(self, mats, mode):
        R = [m for i, m in enumerate(mats) if i != mode and m is not None][0].shape[1]
        ms = [np.ascontiguousarray(m, dtype=np.float64) if m is not None
              else np.zeros((self.dims[i], R)) for i, m in enumerate(mats)]
        out = np.empty((self.dims[mode], R), dtype=np.float64)
        self.lib.oracle_mttkrp_csf(self.h, C.c_uint64(R), _pp(ms, C.c_double), C.c_uint64(mode),
                                   out.ctypes.data_as(dp))
        return out

    def __del__(self):
        try:
            if self.h:
                self.lib.oracle_csf_free(self.h)
                self.h = None
        except Exception:
            pass


def cpd_als(dims, ind, vals, R, niters, tol, seed):
    lib = load()
    d = _u64(dims)
    inds = [_u64(i) for i in ind]
    v = np.ascontiguousarray(vals, dtype=np.float64)
    facs = [np.empty((int(x), R), dtype=np.float64) for x in d]
    lam = np.empty(R, dtype=np.float64)
    fit = lib.oracle_cpd_als(C.c_uint64(len(d)), d.ctypes.data_as(u64p), C.c_uint64(len(v)),
                             _pp(inds, C.c_uint64), v.ctypes.data_as(dp), C.c_uint64(R),
                             C.c_uint64(niters), C.c_double(tol), C.c_uint(seed),
                             _pp(facs, C.c_double), lam.ctypes.data_as(dp))
    return float(fit), lam, facs
