"""ctypes binding of the CPU oracle (oracle/libgkq_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this
module; the product package (gkquad-rs_b200/) never does.  The oracle restates the
reference crate gkquad 0.0.4 line by line (see gkq_oracle.hpp) and is pinned by the
reference's own golden vectors (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libgkq_oracle.so"

ALGO_AUTO, ALGO_QAGS, ALGO_QAGP = 0, 1, 2
TOL_ABSOLUTE, TOL_RELATIVE, TOL_ABS_OR_REL, TOL_ABS_AND_REL = 0, 1, 2, 3
STATUS_NAMES = ["None", "InsufficientIteration", "RoundoffError", "SubrangeTooSmall",
                "Divergent", "NanValueEncountered"]


def build(force: bool = False) -> Path:
    """Compile the oracle with its Makefile (g++ -O3 -ffp-contract=off)."""
    srcs = [_HERE / "oracle_capi.cpp", _HERE / "gkq_oracle.hpp", _HERE / "functors.hpp"]
    stale = (not _LIB_PATH.exists()) or any(s.stat().st_mtime > _LIB_PATH.stat().st_mtime for s in srcs)
    if force or stale:
        env = dict(os.environ)
        subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s"], check=True, env=env)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            build()
        L = C.CDLL(str(_LIB_PATH))
        dp = C.POINTER(C.c_double)
        u32p = C.POINTER(C.c_uint32)
        i64p = C.POINTER(C.c_int64)
        L.gkqo_num_threads.restype = C.c_int
        L.gkqo_functor_info.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int)]
        L.gkqo_integrate_batch.argtypes = [
            C.c_int, C.c_int, C.c_int64, dp, dp, C.c_int64, dp, C.c_int64, i64p, dp, C.c_int64,
            C.c_int, C.c_double, C.c_double, C.c_uint64, C.c_uint64, dp, dp, u32p, u32p, C.c_int]
        L.gkqo_integrate_batch_ex.argtypes = [
            C.c_int, C.c_int, C.c_int64, dp, dp, C.c_int64, dp, C.c_int64, i64p, dp, C.c_int64,
            C.c_int, C.c_double, C.c_double, C.c_uint64, C.c_uint64, dp, dp, C.POINTER(C.c_uint64),
            dp, dp, u32p, u32p, C.c_int]
        L.gkqo_integrate2_batch.argtypes = [
            C.c_int, C.c_int, C.c_int, C.c_int64, dp, dp, C.c_int64, dp, C.c_int64, dp, C.c_int64,
            i64p, dp, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_uint64, dp, dp, u32p, u32p,
            C.c_int, C.c_int, dp, dp, C.POINTER(C.c_uint64)]
        L.gkqo_qk.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, dp, dp]
        L.gkqo_integrate_one.argtypes = [
            C.c_int, C.c_int, C.c_double, C.c_double, dp, dp, C.c_int64, C.c_int, C.c_double,
            C.c_double, C.c_uint64, C.c_uint64, dp, u32p, u32p, i64p, C.c_int64, i64p]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def functor_table(dim: int):
    """[(id, name, nparams)] for dim 1 (1-D), 2 (2-D) or 0 (y-range functors)."""
    out = []
    i = 0
    while True:
        name = C.c_char_p()
        npar = C.c_int()
        if lib().gkqo_functor_info(dim, i, C.byref(name), C.byref(npar)) != 0:
            break
        out.append((i, name.value.decode(), npar.value))
        i += 1
    return out


def functor_id(dim: int, name: str) -> int:
    for i, nm, _ in functor_table(dim):
        if nm == name:
            return i
    raise KeyError(name)


def tol_tuple(tol):
    """('abs', x) | ('rel', y) | ('abs_or_rel', x, y) | ('abs_and_rel', x, y) -> (kind, abs, rel)."""
    k = tol[0]
    if k == "abs":
        return TOL_ABSOLUTE, float(tol[1]), 0.0
    if k == "rel":
        return TOL_RELATIVE, 0.0, float(tol[1])
    if k == "abs_or_rel":
        return TOL_ABS_OR_REL, float(tol[1]), float(tol[2])
    if k == "abs_and_rel":
        return TOL_ABS_AND_REL, float(tol[1]), float(tol[2])
    raise ValueError(tol)


def qk(rule: int, functor: str, a: float, b: float, params=()):
    p = np.asarray(list(params) + [0.0], dtype=np.float64)
    out = np.zeros(4)
    rc = lib().gkqo_qk(rule, functor_id(1, functor), a, b, _dp(p), _dp(out))
    assert rc == 0, rc
    return dict(estimate=out[0], delta=out[1], absvalue=out[2], asc=out[3])


def integrate_one(algo, functor, a, b, params=(), points=(), tol=("abs_or_rel", 1.49e-8, 1.49e-8),
                  max_evals=2000, ws_capacity=0):
    p = np.asarray(list(params) + [0.0], dtype=np.float64)
    pts = np.asarray(list(points) + [0.0], dtype=np.float64)
    out2 = np.zeros(2)
    nev = C.c_uint32()
    st = C.c_uint32()
    order = np.zeros(4096, dtype=np.int64)
    olen = C.c_int64()
    k, ta, tr = tol_tuple(tol)
    rc = lib().gkqo_integrate_one(algo, functor_id(1, functor), a, b, _dp(p), _dp(pts), len(points),
                                  k, ta, tr, max_evals, ws_capacity, _dp(out2), C.byref(nev),
                                  C.byref(st), order.ctypes.data_as(C.POINTER(C.c_int64)), 4096,
                                  C.byref(olen))
    assert rc == 0, rc
    return dict(estimate=out2[0], delta=out2[1], nevals=nev.value, status=st.value,
                order=order[:olen.value].tolist())


def integrate_batch(algo, functor, a, b, params=None, points=None, pts_offsets=None,
                    tol=("abs_or_rel", 1.49e-8, 1.49e-8), max_evals=2000, ws_capacity=0,
                    nthreads=0, n=None, abs_tol_each=None, rel_tol_each=None, max_evals_each=None):
    """a, b: scalars (shared range; n required) or arrays[n]. params: array [np, n] (SoA) or [np]
    (shared). points: shared list, or flat array with pts_offsets[n+1].  *_each: optional per-integral
    tolerance values / budgets (the tolerance kind stays that of `tol`; budgets are clamped to max_evals)."""
    fid = functor_id(1, functor)
    if np.isscalar(a):
        assert n is not None
        aa = np.asarray([a], dtype=np.float64)
        bb = np.asarray([b], dtype=np.float64)
        rs = 0
    else:
        aa = np.ascontiguousarray(a, dtype=np.float64)
        bb = np.ascontiguousarray(b, dtype=np.float64)
        n = aa.shape[0]
        rs = 1
    if params is None:
        pp = np.zeros(1)
        ps = 0
    else:
        pp = np.ascontiguousarray(params, dtype=np.float64)
        ps = n if pp.ndim == 2 else 0
        if pp.ndim == 2:
            assert pp.shape[1] == n
    if pts_offsets is not None:
        po = np.ascontiguousarray(pts_offsets, dtype=np.int64)
        pts = np.ascontiguousarray(points, dtype=np.float64)
        po_p = po.ctypes.data_as(C.POINTER(C.c_int64))
        npts = 0
    else:
        pts = np.asarray(list(points or []) + [0.0], dtype=np.float64)
        npts = len(points or [])
        po_p = None
    est = np.empty(n)
    dlt = np.empty(n)
    nev = np.empty(n, dtype=np.uint32)
    st = np.empty(n, dtype=np.uint32)
    k, ta, tr = tol_tuple(tol)
    ate = None if abs_tol_each is None else np.ascontiguousarray(abs_tol_each, dtype=np.float64)
    rte = None if rel_tol_each is None else np.ascontiguousarray(rel_tol_each, dtype=np.float64)
    mee = None if max_evals_each is None else np.ascontiguousarray(max_evals_each, dtype=np.uint64)
    rc = lib().gkqo_integrate_batch_ex(
        algo, fid, n, _dp(aa), _dp(bb), rs, _dp(pp), ps, po_p, _dp(pts), npts, k, ta, tr,
        max_evals, ws_capacity, None if ate is None else _dp(ate), None if rte is None else _dp(rte),
        None if mee is None else mee.ctypes.data_as(C.POINTER(C.c_uint64)), _dp(est), _dp(dlt),
        nev.ctypes.data_as(C.POINTER(C.c_uint32)), st.ctypes.data_as(C.POINTER(C.c_uint32)), nthreads)
    assert rc == 0, rc
    return dict(estimate=est, delta=dlt, nevals=nev, status=st)


def integrate2_batch(algo, functor2, yrange, x0, x1, fparams=None, yparams=None, points=None,
                     pts_offsets=None, tol=("abs_or_rel", 1.49e-8, 1.49e-8), max_evals=100000,
                     nthreads=0, n=None, swap_xy=False, abs_tol_each=None, rel_tol_each=None,
                     max_evals_each=None):
    """points: list of (x, y) pairs shared by the batch, or flat [m,2] array with pts_offsets.
    swap_xy: the DynamicX form -- the algorithms see g(x, y) = f(y, x); ranges and points are given
    in (outer, inner) order, as the reference passes them on after its own swap."""
    fid = functor_id(2, functor2)
    yid = functor_id(0, yrange)
    if np.isscalar(x0):
        assert n is not None
        aa = np.asarray([x0], dtype=np.float64)
        bb = np.asarray([x1], dtype=np.float64)
        rs = 0
    else:
        aa = np.ascontiguousarray(x0, dtype=np.float64)
        bb = np.ascontiguousarray(x1, dtype=np.float64)
        n = aa.shape[0]
        rs = 1

    def prep(p):
        if p is None:
            return np.zeros(1), 0
        p = np.ascontiguousarray(p, dtype=np.float64)
        return p, (n if p.ndim == 2 else 0)

    fp, fs = prep(fparams)
    yp, ys = prep(yparams)
    if pts_offsets is not None:
        po = np.ascontiguousarray(pts_offsets, dtype=np.int64)
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1)
        po_p = po.ctypes.data_as(C.POINTER(C.c_int64))
        npts = 0
    else:
        flat = [c for xy in (points or []) for c in xy]
        pts = np.asarray(flat + [0.0], dtype=np.float64)
        npts = len(points or [])
        po_p = None
    est = np.empty(n)
    dlt = np.empty(n)
    nev = np.empty(n, dtype=np.uint32)
    st = np.empty(n, dtype=np.uint32)
    k, ta, tr = tol_tuple(tol)
    ate = None if abs_tol_each is None else np.ascontiguousarray(abs_tol_each, dtype=np.float64)
    rte = None if rel_tol_each is None else np.ascontiguousarray(rel_tol_each, dtype=np.float64)
    mee = None if max_evals_each is None else np.ascontiguousarray(max_evals_each, dtype=np.uint64)
    rc = lib().gkqo_integrate2_batch(
        algo, fid, yid, n, _dp(aa), _dp(bb), rs, _dp(fp), fs, _dp(yp), ys, po_p, _dp(pts), npts,
        k, ta, tr, max_evals, _dp(est), _dp(dlt),
        nev.ctypes.data_as(C.POINTER(C.c_uint32)), st.ctypes.data_as(C.POINTER(C.c_uint32)), nthreads,
        1 if swap_xy else 0, None if ate is None else _dp(ate), None if rte is None else _dp(rte),
        None if mee is None else mee.ctypes.data_as(C.POINTER(C.c_uint64)))
    assert rc == 0, rc
    return dict(estimate=est, delta=dlt, nevals=nev, status=st)


def num_threads() -> int:
    return lib().gkqo_num_threads()
